// mpm_substep3d.cu -- the 3D G2P kernel of the default path (single GPU and x-slabs).
//
// Reference statements (cpp_validation/mls-mpm88-explained.cpp, lifted to 27 nodes): :134-179 -- gather v and C,
// advect, F update, plastic projection of snow, Jp.  Same arithmetic as k_g2p_naive<3, FAST> in mpm_kernels.cu (which
// stays as the MPM_FLAG_STRICT / MPM_FLAG_NAIVE kernel); this file is about what bounded that kernel on sm_100a --
// memory latency at 28 % occupancy (86 registers: the whole particle state stayed live across the projection):
//   * C and v are FINAL right after the gather: they are stored at once and their 12 registers are free while the
//     polar + Jacobi projection of F runs (plastic_project3); x and Jp follow with F;
//   * an emigrating particle (x-slab runs, rare) reserves its message slot when its new position is known and packs
//     its record at the end from what the thread itself just stored -- nothing is kept live for that branch;
//   * RESORT = true performs the storage re-sort on the fly like the 2D substep kernel: the new state goes to the
//     slot k_count_rank assigned in the OTHER storage buffer (one 12-byte pass per re-sort instead of an extra
//     216 B / particle reorder pass);
//   * the largest displacement of the substep (CFL, feeds the re-sort interval) comes from dt * v, not from re-reading x.
#include "mpm_gather3.cuh"
#include "mpm_kernels.cuh"

namespace mpm {

namespace {

#ifndef MPM_G2P3_FAST_MINB
#define MPM_G2P3_FAST_MINB 7
#endif
#ifndef MPM_G2P3_WAVES
#define MPM_G2P3_WAVES 0  // > 0: cap the grid at this many CTAs per resident slot, the rest is the grid-stride loop with the
                          // position prefetch.  Measured on c5: 4 waves 1.60 ms, uncapped (one particle per thread) 1.52 ms
#endif

// x-slab runs: the record of an emigrant (layout of emigrate() in mpm_kernels.cu: x v F C Jp mat id pad), read back from
// where the thread stored the particle's new state a moment ago (`out`, slot dst).  Out of line and fed with four
// integers only: inlined, the 28-word record of this rare branch raised the register pressure of EVERY particle (the
// MIG variants of the 3D G2P spilled 164 bytes and ran 28 % slower than the single-GPU variant on the first 2-GPU run).
__device__ __noinline__ void pack_emigrant3(const SoA<3> &out, long long dst, int slot_side, int mat_id, int id,
                                            float *send_lo, float *send_hi) {
  const int sd = slot_side >> 30;
  float4 *r = reinterpret_cast<float4 *>((sd == 0 ? send_lo : send_hi) + (size_t)(slot_side & 0x3fffffff) * MigRec<3>::WORDS);
  const float4 xj = out.xj[dst], vv = out.vm[dst];
  float F[9], C[9];
#pragma unroll
  for (int k = 0; k < 9; k++) {
    F[k] = out.F[k][dst];
    C[k] = out.C[k][dst];
  }
  r[0] = make_float4(xj.x, xj.y, xj.z, vv.x);
  r[1] = make_float4(vv.y, vv.z, F[0], F[1]);
  r[2] = make_float4(F[2], F[3], F[4], F[5]);
  r[3] = make_float4(F[6], F[7], F[8], C[0]);
  r[4] = make_float4(C[1], C[2], C[3], C[4]);
  r[5] = make_float4(C[5], C[6], C[7], C[8]);
  r[6] = make_float4(xj.w, __int_as_float(mat_id), __int_as_float(id), 0.0f);
}

// tile of node velocities a CTA of k_g2p3_tile stages in shared memory: the bin's 4^3 base cells + 1-cell drift margin
// + stencil reach = 8 nodes per axis; z rows padded to 9 float4 (144 B) so that neighbouring y rows do not share banks
constexpr int TN = 8, TZ = 9, TILE_NODES = TN * TN * TZ;

// One particle of the 3D G2P (:134-179).  TILE: gather from the shared-memory tile whose node (0,0,0) is the global
// node (tox, toy, toz) when the stencil lies inside it, else (drifted past the margin: rare) from global memory.
template <bool MIG, bool FLIP, bool RESORT, bool TILE>
__device__ __forceinline__ void g2p3_particle(const G2p3Args &A, long long i, float4 xj, float4 vm, Mat<3> F,
                                              const float4 *tile, const float4 *tile_o, int tox, int toy, int toz,
                                              float &vmax) {
  const Params &P = A.P;
  const int mat_id = __float_as_int(vm.w);
  const SoA<3> &out = RESORT ? A.d : A.s;
  long long dst = i;
  if (RESORT) dst = (long long)A.new_start[A.key[i]] + A.rank[i];
  float x[3] = {xj.x, xj.y, xj.z};
  float Jp = xj.w;
  const Material &mat = P.mat[material_index(P, mat_id)];
  // ---- gather (:136-156), advect (:159), F update (:162) ----
  Stencil<3> st = make_stencil<3>(x, P.inv_dx);
  clamp_base<3>(P, st.base);  // G2P never flags (P2G did)
  float v[3], dv[3] = {0.0f, 0.0f, 0.0f};
  Mat<3> C;
  bool from_tile = false;
  if (TILE) {
    const int l0 = st.base[0] - tox, l1 = st.base[1] - toy, l2 = st.base[2] - toz;
    from_tile = (unsigned)l0 <= (unsigned)(TN - 3) && (unsigned)l1 <= (unsigned)(TN - 3) && (unsigned)l2 <= (unsigned)(TN - 3);
    if (from_tile) {
      const int at = (l0 * TN + l1) * TZ + l2;
      gather3_rows<false>(st, tile + at, tile_o + at, TN * TZ, TZ, FLIP, v, C, dv);
    }
  }
  if (!from_tile) gather3_fast(P, st, A.grid, A.vold, FLIP, v, C, dv);
  const float s4 = 4 * P.inv_dx;  // the constant of :154, applied once
#pragma unroll
  for (int cc = 0; cc < 3; cc++)
#pragma unroll
    for (int r = 0; r < 3; r++) C.d[cc][r] = s4 * C.d[cc][r];
  vmax = fmaxf(vmax, fmaxf(fabsf(v[0]), fmaxf(fabsf(v[1]), fabsf(v[2]))));
#pragma unroll
  for (int k = 0; k < 3; k++) x[k] = x[k] + A.dt * v[k];
  if (FLIP) {
    const float a = P.alpha;
    const float v_in[3] = {vm.x, vm.y, vm.z};
#pragma unroll
    for (int k = 0; k < 3; k++) v[k] = (1.0f - a) * v[k] + a * (v_in[k] + dv[k]);
  }
  F = mat_mul<3>(mat_add<3>(mat_diag<3>(1.0f), mat_scale<3>(A.dt, C)), F);
  // ---- C and v are final: out they go ----
#pragma unroll
  for (int c = 0; c < 3; c++)
#pragma unroll
    for (int r = 0; r < 3; r++) out.C[c * 3 + r][dst] = C.d[c][r];
  // x-slab runs: a particle whose NEW base column left the slab reserves its slot in that side's message
  int side = -1, slot = 0;
  if (MIG) {
    const int nbx = max(0, min(base_coord(x[0], P.inv_dx), P.n_grid - 2));
    if (A.mig.interior) {
      if ((P.slab_lo > 0 && nbx < P.slab_lo + 2) || (P.slab_hi < P.n_grid && nbx + 4 > P.slab_hi))
        atomicOr(A.status, STATUS_CFL);
    } else {
      side = nbx < P.slab_lo ? 0 : (nbx >= P.slab_hi ? 1 : -1);
      if (side >= 0) {
        slot = atomicAdd(&A.mig.count[side], 1);
        if (slot >= A.mig.cap) {
          atomicOr(A.status, STATUS_MIGRATION_OVERFLOW);  // stays here (and will be flagged out of slab)
          side = -1;
        }
      }
    }
  }
  out.vm[dst] = make_float4(v[0], v[1], v[2], __int_as_float(side >= 0 ? DEAD : mat_id));
  if (MIG && side >= 0) slot |= side << 30;  // carried across the projection in one register
  else slot = -1;
  // ---- plasticity (:165-178) ----
  if (mat.kind == KIND_SNOW) {
    const float ratio = plastic_project3(mat.sig_lo, mat.sig_hi, F);  // det(F) / det(F')
    Jp = clampf(Jp * ratio, P.jp_min, P.jp_max);
  } else if (mat.kind != KIND_JELLY) {
    fluid_project(F);
  }
  out.xj[dst] = make_float4(x[0], x[1], x[2], Jp);
#pragma unroll
  for (int c = 0; c < 3; c++)
#pragma unroll
    for (int r = 0; r < 3; r++) out.F[c * 3 + r][dst] = F.d[c][r];
  int id = 0;
  if (RESORT || slot >= 0) id = A.s.id[i];
  if (RESORT) out.id[dst] = id;
  if (MIG && slot >= 0) pack_emigrant3(out, dst, slot, mat_id, id, A.mig.send_lo, A.mig.send_hi);
}

__device__ __forceinline__ void g2p3_report(const G2p3Args &A, float vmax) {
  if (A.stats) {  // largest displacement of this substep in cells: feeds the re-sort interval (see engine)
    const unsigned bits = __reduce_max_sync(0xffffffffu, __float_as_uint(vmax * A.dt * A.P.inv_dx));
    unsigned *slot = reinterpret_cast<unsigned *>(&A.stats[2]);
    if ((threadIdx.x & 31) == 0 && bits > *reinterpret_cast<volatile unsigned *>(slot)) atomicMax(slot, bits);
  }
}

// thread per particle, slots [first, n): the immigrant tail behind the binned range, and everything when
// MPM_G2P3_TILE is off
template <bool MIG, bool FLIP, bool RESORT>
__global__ void __launch_bounds__(128, MPM_G2P3_FAST_MINB) k_g2p3(const __grid_constant__ G2p3Args A) {
  long long n = A.n;
  if (MIG && A.dev_n && n > *A.dev_n) n = *A.dev_n;  // x-slab handles: exact extent on the device
  float vmax = 0.0f;
  long long i = A.first + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (MPM_G2P3_WAVES > 0) {
    // grid-stride loop with ONE prefetch, the position of the thread's next particle (see MPM_G2P3_WAVES)
    const long long stride = (long long)gridDim.x * blockDim.x;
    float4 xj_next = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    if (i < n) xj_next = A.s.xj[i];
    for (; i < n; i += stride) {
      const float4 xj = xj_next;
      if (i + stride < n) xj_next = A.s.xj[i + stride];
      const float4 vm = A.s.vm[i];
      if (MIG && __float_as_int(vm.w) == DEAD) continue;  // slot of a particle that emigrated earlier
      Mat<3> F;
#pragma unroll
      for (int c = 0; c < 3; c++)
#pragma unroll
        for (int r = 0; r < 3; r++) F.d[c][r] = A.s.F[c * 3 + r][i];
      g2p3_particle<MIG, FLIP, RESORT, false>(A, i, xj, vm, F, nullptr, nullptr, 0, 0, 0, vmax);
    }
  } else if (i < n) {
    const float4 xj = A.s.xj[i], vm = A.s.vm[i];
    if (!(MIG && __float_as_int(vm.w) == DEAD)) {  // slot of a particle that emigrated earlier: dropped by the re-sort
      Mat<3> F;
#pragma unroll
      for (int c = 0; c < 3; c++)
#pragma unroll
        for (int r = 0; r < 3; r++) F.d[c][r] = A.s.F[c * 3 + r][i];
      g2p3_particle<MIG, FLIP, RESORT, false>(A, i, xj, vm, F, nullptr, nullptr, 0, 0, 0, vmax);
    }
  }
  g2p3_report(A, vmax);
}

// [opt-in at build time, -DMPM_G2P3_TILE=1: measured SLOWER than k_g2p3 on B200 -- 1.83 ms (5 CTAs/SM) ... 2.23 ms (6)
// against 1.48 ms on the 256^3 scene: the 14 prefetch registers spill at 72 and 27 LDS.128 per particle load the L1
// pipe as much as the read-only-path gather did]
// CTA per chunk of a bin: the 8^3 node velocities the bin's particles can touch are staged in shared memory ONCE
// (coalesced, one memory latency per CTA); a particle's critical chain position -> base cell -> 27 node loads then
// ends in shared memory, and the loads of the thread's next particle (position, material, F) are in flight while
// the current one is computed.  The thread-per-particle kernel above was bound by exactly that chain: long-scoreboard
// stalls 7.8 warps per issue at 39 % occupancy (profiles/r02_ncu_full_c5.md).
template <bool MIG, bool FLIP, bool RESORT>
__global__ void __launch_bounds__(128, MPM_G2P3_FAST_MINB) k_g2p3_tile(const __grid_constant__ G2p3Args A) {
  __shared__ float4 tile[TILE_NODES];
  __shared__ float4 tile_o[FLIP ? TILE_NODES : 1];
  const Params &P = A.P;
  const int tid = threadIdx.x;
  const int4 work = A.chunks[blockIdx.x];
  const int c0 = work.y, m = work.z;
  // global node of tile node (0,0,0): bin origin minus the drift margin
  const int tox = (work.w >> 20) * 4 + P.slab_lo - 1, toy = ((work.w >> 10) & 0x3ff) * 4 - 1, toz = (work.w & 0x3ff) * 4 - 1;
  // first particle of this thread: in flight while the tile loads
  float4 xj_n = make_float4(0.0f, 0.0f, 0.0f, 0.0f), vm_n = xj_n;
  Mat<3> F_n;
  if (tid < m) {
    xj_n = A.s.xj[c0 + tid];
    vm_n = A.s.vm[c0 + tid];
#pragma unroll
    for (int c = 0; c < 3; c++)
#pragma unroll
      for (int r = 0; r < 3; r++) F_n.d[c][r] = A.s.F[c * 3 + r][c0 + tid];
  }
  for (int t = tid; t < TN * TN * TN; t += 128) {
    const int li = t >> 6, lj = (t >> 3) & 7, lk = t & 7;
    const int gi = tox + li, gj = toy + lj, gk = toz + lk;
    float4 g4 = make_float4(0.0f, 0.0f, 0.0f, 0.0f), o4 = g4;
    if (gi >= P.slab_lo && gi < P.slab_lo + P.ncol && gj >= 0 && gj < P.n1 && gk >= 0 && gk < P.n1) {
      const long long node = ((long long)(gi - P.slab_lo) * P.n1 + gj) * P.n1 + gk;
      g4 = __ldg(A.grid + node);
      if (FLIP) o4 = __ldg(A.vold + node);
    }
    tile[(li * TN + lj) * TZ + lk] = g4;
    if (FLIP) tile_o[(li * TN + lj) * TZ + lk] = o4;
  }
  __syncthreads();
  float vmax = 0.0f;
  for (int i = tid; i < m; i += 128) {
    const float4 xj = xj_n, vm = vm_n;
    const Mat<3> F = F_n;
    if (i + 128 < m) {  // software pipeline: the next particle's loads go out before this one is computed
      xj_n = A.s.xj[c0 + i + 128];
      vm_n = A.s.vm[c0 + i + 128];
#pragma unroll
      for (int c = 0; c < 3; c++)
#pragma unroll
        for (int r = 0; r < 3; r++) F_n.d[c][r] = A.s.F[c * 3 + r][c0 + i + 128];
    }
    if (MIG && __float_as_int(vm.w) == DEAD) continue;
    g2p3_particle<MIG, FLIP, RESORT, true>(A, (long long)c0 + i, xj, vm, F, tile, tile_o, tox, toy, toz, vmax);
  }
  g2p3_report(A, vmax);
}

}  // namespace

namespace {

// ------------------------------------------------------------------------------------------------------------------
// The fused 3D substep kernel: G2P of substep n and P2G of substep n+1 in one pass over the particles, one CTA per
// chunk (<= 512 particles) of a non-empty bin of 4x4x4 cells -- the 3D counterpart of k_substep2d (mpm_substep2d.cu).
// Reference statements: :134-179 of the current substep, then :53-102 of the next on the state still in registers.
//   * the G2P half is k_g2p3's body (early stores of C and v, emigrant slot reserved early, on-the-fly re-sort);
//   * the snow projection hands its rotation factor R to the stress of the next P2G (:75-76) -- F' = R S', so the
//     2-iteration Newton polar the stand-alone P2G runs on every snow particle is gone (jelly still runs it);
//   * m C dx waits in the particle's shared-memory record while the projection runs (9 registers less across it);
//     the stress is added to it afterwards: cs_k = (stress + m C) column k * dx, q = m v - sum_k fx_k cs_k;
//   * phase 2 = the cell gather of k_p2g_cells<3>: three threads per work item (one stencil row each), 36 register
//     accumulators, one RED.E.ADD.F32x4 per (item, node); records as four float4 planes.
// ------------------------------------------------------------------------------------------------------------------
#ifndef MPM_SUBSTEP3D_MINB
#define MPM_SUBSTEP3D_MINB 5
#endif
constexpr int B3 = 4, NT3 = 128, CAP3 = 512, M3 = 1, L3 = B3 + 2 * M3, NC3 = L3 * L3 * L3;
constexpr int RM3 = 8, MAXI3 = NC3 + CAP3 / RM3 + 1;

template <bool FLIP, bool MIG, bool RESORT>
__global__ void __launch_bounds__(NT3, MPM_SUBSTEP3D_MINB) k_substep3d(const __grid_constant__ Substep3dArgs A) {
  __shared__ float4 recp[4 * CAP3];     // planes: (fx.xyz, q.x) | (q.y, q.z, cs00, cs01) | (cs02, cs10, cs11, cs12) | (cs20, cs21, cs22, -)
  __shared__ unsigned cr[CAP3];         // (local cell << 16) | rank in cell; 0xffffffff = not binned
  __shared__ unsigned short sorted[CAP3];
  __shared__ int cnt[NC3 + 4];          // per-cell count, then record start
  __shared__ unsigned item[MAXI3];      // cell | first record << 8 | length << 20
  __shared__ int wtot[4], n_items_sh;
  const Params &P = A.P;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int4 work = A.chunks[blockIdx.x];
  const int c0 = work.y, m = work.z;
  // global cell of local cell 0
  const int ox = (work.w >> 20) * B3 + P.slab_lo - M3, oy = ((work.w >> 10) & 0x3ff) * B3 - M3, oz = (work.w & 0x3ff) * B3 - M3;
  const long long n1 = P.n1;
  const float dxs = P.dx;
  unsigned n_fallback = 0;
  float vmax = 0.0f;
  for (int k = tid; k < NC3; k += NT3) cnt[k] = 0;
  __syncthreads();
  // ---------------- phase 1: thread per particle (G2P of this substep, P2G record of the next) ----------------
  float4 xj_next = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  if (tid < m) xj_next = A.s.xj[c0 + tid];
  for (int i = tid; i < m; i += NT3) {
    const long long slot_i = (long long)c0 + i;
    const float4 xj = xj_next;  // issued one iteration ago
    if (i + NT3 < m) xj_next = A.s.xj[slot_i + NT3];
    const float4 vm = A.s.vm[slot_i];
    const int mat_id = __float_as_int(vm.w);
    if (MIG && mat_id == DEAD) {  // slot of a particle that emigrated earlier; dropped by the re-sort
      cr[i] = 0xffffffffu;
      continue;
    }
    Mat<3> F;
#pragma unroll
    for (int c = 0; c < 3; c++)
#pragma unroll
      for (int r = 0; r < 3; r++) F.d[c][r] = A.s.F[c * 3 + r][slot_i];
    const SoA<3> &out = RESORT ? A.d : A.s;
    long long dst = slot_i;
    if (RESORT) dst = (long long)A.new_start[A.key[slot_i]] + A.rank[slot_i];
    float x[3] = {xj.x, xj.y, xj.z};
    float Jp = xj.w;
    const Material &mat = P.mat[material_index(P, mat_id)];
    float mv[3];
    int gone_slot = -1;
    Mat<3> R;
    bool have_R = false;
    {
      // ---- G2P :134-179 ----
      Stencil<3> st = make_stencil<3>(x, P.inv_dx);
      clamp_base<3>(P, st.base);  // G2P never flags (P2G did)
      float v[3], dv[3] = {0.0f, 0.0f, 0.0f};
      Mat<3> C;
      gather3_fast(P, st, A.grid_in, A.vold_in, FLIP, v, C, dv);
      const float s4 = 4 * P.inv_dx;  // the constant of :154, applied once
#pragma unroll
      for (int cc = 0; cc < 3; cc++)
#pragma unroll
        for (int r = 0; r < 3; r++) C.d[cc][r] = s4 * C.d[cc][r];
      vmax = fmaxf(vmax, fmaxf(fabsf(v[0]), fmaxf(fabsf(v[1]), fabsf(v[2]))));
#pragma unroll
      for (int k = 0; k < 3; k++) x[k] = x[k] + A.dt_g2p * v[k];
      if (FLIP) {
        const float a = P.alpha;
        const float v_in[3] = {vm.x, vm.y, vm.z};
#pragma unroll
        for (int k = 0; k < 3; k++) v[k] = (1.0f - a) * v[k] + a * (v_in[k] + dv[k]);
      }
      F = mat_mul<3>(mat_add<3>(mat_diag<3>(1.0f), mat_scale<3>(A.dt_g2p, C)), F);
      // C and v are final: to global memory; m C dx waits in this particle's record for the stress
#pragma unroll
      for (int c = 0; c < 3; c++)
#pragma unroll
        for (int r = 0; r < 3; r++) out.C[c * 3 + r][dst] = C.d[c][r];
      {
        const float md = P.mass_p * dxs;
        recp[CAP3 + i] = make_float4(0.0f, 0.0f, md * C.d[0][0], md * C.d[0][1]);
        recp[2 * CAP3 + i] = make_float4(md * C.d[0][2], md * C.d[1][0], md * C.d[1][1], md * C.d[1][2]);
        recp[3 * CAP3 + i] = make_float4(md * C.d[2][0], md * C.d[2][1], md * C.d[2][2], 0.0f);
      }
#pragma unroll
      for (int k = 0; k < 3; k++) mv[k] = P.mass_p * v[k];
      int side = -1, slot = 0;
      if (MIG) {
        const int nbx = max(0, min(base_coord(x[0], P.inv_dx), P.n_grid - 2));
        if (A.mig.interior) {
          if ((P.slab_lo > 0 && nbx < P.slab_lo + 2) || (P.slab_hi < P.n_grid && nbx + 4 > P.slab_hi))
            atomicOr(A.status, STATUS_CFL);
        } else {
          side = nbx < P.slab_lo ? 0 : (nbx >= P.slab_hi ? 1 : -1);
          if (side >= 0) {
            slot = atomicAdd(&A.mig.count[side], 1);
            if (slot >= A.mig.cap) {
              atomicOr(A.status, STATUS_MIGRATION_OVERFLOW);  // stays here (and will be flagged out of slab)
              side = -1;
            }
          }
        }
      }
      out.vm[dst] = make_float4(v[0], v[1], v[2], __int_as_float(side >= 0 ? DEAD : mat_id));
      if (MIG && side >= 0) gone_slot = slot | (side << 30);
    }
    // ---- plasticity (:165-178); snow hands its rotation factor to the next P2G ----
    if (mat.kind == KIND_SNOW) {
      const float ratio = plastic_project3(mat.sig_lo, mat.sig_hi, F, &R);  // det(F) / det(F')
      Jp = clampf(Jp * ratio, P.jp_min, P.jp_max);
      have_R = true;
    } else if (mat.kind != KIND_JELLY) {
      fluid_project(F);
    }
    out.xj[dst] = make_float4(x[0], x[1], x[2], Jp);
#pragma unroll
    for (int c = 0; c < 3; c++)
#pragma unroll
      for (int r = 0; r < 3; r++) out.F[c * 3 + r][dst] = F.d[c][r];
    int id = 0;
    if (RESORT || gone_slot >= 0) id = A.s.id[slot_i];
    if (RESORT) out.id[dst] = id;
    if (MIG && gone_slot >= 0) {
      pack_emigrant3(out, dst, gone_slot, mat_id, id, A.mig.send_lo, A.mig.send_hi);
      cr[i] = 0xffffffffu;  // no P2G here: the receiving handle scatters it when it arrives
      continue;
    }
    {
      // ---- P2G record of the next substep :53-89 ----
      Stencil<3> st = make_stencil<3>(x, P.inv_dx);
      const int bad = clamp_base<3>(P, st.base);
      if (bad) atomicOr(A.status, bad);
      const Mat<3> stress = p2g_affine<3>(P, mat, A.dt_p2g, F, mat_zero<3>(), Jp, have_R ? &R : nullptr);  // :67-84 (C = 0)
      // cs_k = (stress + m C) column k * dx; q = m v - sum_k fx_k cs_k
      const float4 pb = recp[CAP3 + i], pc = recp[2 * CAP3 + i], pd = recp[3 * CAP3 + i];
      float cs[3][3] = {{pb.z, pb.w, pc.x}, {pc.y, pc.z, pc.w}, {pd.x, pd.y, pd.z}};
#pragma unroll
      for (int k = 0; k < 3; k++)
#pragma unroll
        for (int r = 0; r < 3; r++) cs[k][r] = __fmaf_rn(stress.d[k][r], dxs, cs[k][r]);
#pragma unroll
      for (int k = 0; k < 3; k++)
#pragma unroll
        for (int r = 0; r < 3; r++) mv[r] = __fmaf_rn(-st.fx[k], cs[k][r], mv[r]);
      const int lx = st.base[0] - ox, ly = st.base[1] - oy, lz = st.base[2] - oz;
      if ((unsigned)lx < (unsigned)L3 && (unsigned)ly < (unsigned)L3 && (unsigned)lz < (unsigned)L3) {
        const int cell = (lx * L3 + ly) * L3 + lz;
        const int r = atomicAdd(&cnt[cell], 1);
        cr[i] = ((unsigned)cell << 16) | (unsigned)r;
        recp[i] = make_float4(st.fx[0], st.fx[1], st.fx[2], mv[0]);
        recp[CAP3 + i] = make_float4(mv[1], mv[2], cs[0][0], cs[0][1]);
        recp[2 * CAP3 + i] = make_float4(cs[0][2], cs[1][0], cs[1][1], cs[1][2]);
        recp[3 * CAP3 + i] = make_float4(cs[2][0], cs[2][1], cs[2][2], 0.0f);
      } else {
        // drifted past the 1-cell bin margin since the last re-sort: per-particle REDs, same separable form
        cr[i] = 0xffffffffu;
        n_fallback++;
        float4 *g0 = A.grid_out + ((long long)(st.base[0] - P.slab_lo) * n1 + st.base[1]) * n1 + st.base[2];
#pragma unroll  // (fully unrolled: a run-time index into st.w would move the stencil to local memory for EVERY particle)
        for (int a = 0; a < 3; a++)
#pragma unroll
          for (int b = 0; b < 3; b++)
#pragma unroll
            for (int c = 0; c < 3; c++) {
              const float w = st.w[a][0] * st.w[b][1] * st.w[c][2];
              float val[3];
#pragma unroll
              for (int r = 0; r < 3; r++)
                val[r] = w * __fmaf_rn((float)c, cs[2][r], __fmaf_rn((float)b, cs[1][r], __fmaf_rn((float)a, cs[0][r], mv[r])));
              atomicAdd(g0 + ((long long)a * n1 + b) * n1 + c, make_float4(val[0], val[1], val[2], w * P.mass_p));
            }
      }
    }
  }
  __syncthreads();
  // ---------------- scans on all four warps (two cells per thread): record starts and work items per cell -----------
  {
    const int k0 = 2 * tid, k1 = 2 * tid + 1;
    const int v0 = k0 < NC3 ? cnt[k0] : 0, v1 = k1 < NC3 ? cnt[k1] : 0;
    const int parts0 = (v0 + RM3 - 1) / RM3, parts1 = (v1 + RM3 - 1) / RM3;
    const int p0 = v0 | (parts0 << 16), p1 = v1 | (parts1 << 16);  // both sums stay below 2^16
    int inc = p0 + p1;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += t;
    }
    if (lane == 31) wtot[wid] = inc;
    __syncthreads();  // also: every count has been read before any is overwritten
    int pre = 0;
#pragma unroll
    for (int w = 0; w < 3; w++)
      if (w < wid) pre += wtot[w];
    int ex = pre + inc - (p0 + p1);
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const int k = h == 0 ? k0 : k1, v = h == 0 ? v0 : v1, parts = h == 0 ? parts0 : parts1;
      if (k < NC3) {
        const int start = ex & 0xffff, ifirst = ex >> 16;
        cnt[k] = start;
        if (parts > 0) {
          const int per = parts == 1 ? v : (v + parts - 1) / parts;  // even split
          for (int sub = 0; sub < parts; sub++) {
            const int n0 = start + sub * per, len = min(per, v - sub * per);
            item[ifirst + sub] = (unsigned)k | ((unsigned)n0 << 8) | ((unsigned)len << 20);
          }
        }
      }
      ex += h == 0 ? p0 : 0;
    }
    if (tid == NT3 - 1) n_items_sh = (pre + inc) >> 16;
  }
  __syncthreads();
  const int n_items = n_items_sh;
  for (int i = tid; i < m; i += NT3) {
    const unsigned c = cr[i];
    if (c != 0xffffffffu) sorted[cnt[c >> 16] + (c & 0xffffu)] = (unsigned short)i;
  }
  __syncthreads();
  // ---------------- phase 2: three threads per work item (one stencil row each), 36 register accumulators ----------
  for (int t3 = tid; t3 < n_items * 3; t3 += NT3) {
    const int it = t3 / 3, a = t3 - it * 3;  // the three rows of an item sit in adjacent lanes: their record reads broadcast
    const unsigned ds = item[it];
    const int cell = ds & 0xff, n0 = (ds >> 8) & 0xfff, len = ds >> 20;
    f2 acc_xy[3][3], acc_zm[3][3];
#pragma unroll
    for (int b = 0; b < 3; b++)
#pragma unroll
      for (int c = 0; c < 3; c++) acc_xy[b][c] = acc_zm[b][c] = sp2(0.0f);
    for (int jj = n0; jj < n0 + len; jj++) {
      const int i = sorted[jj];
      const float4 ra = recp[i], rb = recp[CAP3 + i], rc = recp[2 * CAP3 + i], rd = recp[3 * CAP3 + i];
      const float fx[3] = {ra.x, ra.y, ra.z};
      float w[3][3];
#pragma unroll
      for (int k = 0; k < 3; k++) {  // the three expressions of :61-63
        w[0][k] = 0.5f * ((1.5f - fx[k]) * (1.5f - fx[k]));
        w[1][k] = 0.75f - ((fx[k] - 1.0f) * (fx[k] - 1.0f));
        w[2][k] = 0.5f * ((fx[k] - 0.5f) * (fx[k] - 0.5f));
      }
      // node value = w_abc * (Q + a CS_0 + b CS_1 + c CS_2) with Q = (q, mass_p), CS_k = (cs_k, 0): two FFMA2 per step
      const f2 cs_xy[3] = {mk2(rb.z, rb.w), mk2(rc.y, rc.z), mk2(rd.x, rd.y)};
      const f2 cs_zm[3] = {mk2(rc.x, 0.0f), mk2(rc.w, 0.0f), mk2(rd.z, 0.0f)};
      const f2 q_xy = mk2(ra.w, rb.x), q_zm = mk2(rb.y, P.mass_p);
      const float wa = a == 0 ? w[0][0] : (a == 1 ? w[1][0] : w[2][0]);
      const f2 xa_xy = fma2(sp2((float)a), cs_xy[0], q_xy), xa_zm = fma2(sp2((float)a), cs_zm[0], q_zm);
#pragma unroll
      for (int b = 0; b < 3; b++) {
        const float wab = wa * w[b][1];
        const f2 yb_xy = b == 0 ? xa_xy : (b == 1 ? add2(xa_xy, cs_xy[1]) : fma2(sp2(2.0f), cs_xy[1], xa_xy));
        const f2 yb_zm = b == 0 ? xa_zm : (b == 1 ? add2(xa_zm, cs_zm[1]) : fma2(sp2(2.0f), cs_zm[1], xa_zm));
#pragma unroll
        for (int c = 0; c < 3; c++) {
          const float wabc = wab * w[c][2];
          const f2 z_xy = c == 0 ? yb_xy : (c == 1 ? add2(yb_xy, cs_xy[2]) : fma2(sp2(2.0f), cs_xy[2], yb_xy));
          const f2 z_zm = c == 0 ? yb_zm : (c == 1 ? add2(yb_zm, cs_zm[2]) : fma2(sp2(2.0f), cs_zm[2], yb_zm));
          acc_xy[b][c] = fma2(sp2(wabc), z_xy, acc_xy[b][c]);
          acc_zm[b][c] = fma2(sp2(wabc), z_zm, acc_zm[b][c]);
        }
      }
    }
    const int lx = cell / (L3 * L3), ly = (cell / L3) % L3, lz = cell % L3;
    float4 *gp = A.grid_out + ((long long)(ox + lx + a - P.slab_lo) * n1 + (oy + ly)) * n1 + (oz + lz);
#pragma unroll
    for (int b = 0; b < 3; b++)
#pragma unroll
      for (int c = 0; c < 3; c++)
        atomicAdd(gp + b * n1 + c, make_float4(acc_xy[b][c].x, acc_xy[b][c].y, acc_zm[b][c].x, acc_zm[b][c].y));  // RED.E.ADD.F32x4
  }
  if (A.stats) {
    if (n_fallback) {
      atomicAdd(&A.stats[0], (unsigned long long)n_fallback);
      atomicAdd(&A.stats[1], (unsigned long long)n_fallback);
    }
    const unsigned bits = __reduce_max_sync(0xffffffffu, __float_as_uint(vmax * A.dt_g2p * P.inv_dx));
    unsigned *slot = reinterpret_cast<unsigned *>(&A.stats[2]);
    if (lane == 0 && bits > *reinterpret_cast<volatile unsigned *>(slot)) atomicMax(slot, bits);
  }
}

}  // namespace

void launch_g2p3(const G2p3Args &a, bool flip, bool mig, bool resort, cudaStream_t st) {
  if (a.n - a.first <= 0) return;
  unsigned blocks = (unsigned)((a.n - a.first + 127) / 128);
  if (MPM_G2P3_WAVES > 0) {
    static int sms[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64) {
      if (!sms[dev]) cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev);
      const unsigned cap = (unsigned)(sms[dev] > 0 ? sms[dev] : 148) * MPM_G2P3_FAST_MINB * MPM_G2P3_WAVES;
      if (blocks > cap) blocks = cap;
    }
  }
#define MPM_G3(F_, M_, R_) k_g2p3<F_, M_, R_><<<blocks, 128, 0, st>>>(a)
  // template order: MIG, FLIP, RESORT
  if (mig) {
    if (flip) { if (resort) MPM_G3(true, true, true); else MPM_G3(true, true, false); }
    else      { if (resort) MPM_G3(true, false, true); else MPM_G3(true, false, false); }
  } else {
    if (flip) { if (resort) MPM_G3(false, true, true); else MPM_G3(false, true, false); }
    else      { if (resort) MPM_G3(false, false, true); else MPM_G3(false, false, false); }
  }
#undef MPM_G3
}

// the binned range through the work list (a.chunks / a.n_chunks), one CTA per chunk
void launch_g2p3_tile(const G2p3Args &a, bool flip, bool mig, bool resort, cudaStream_t st) {
  if (a.n_chunks <= 0) return;
#define MPM_G3T(F_, M_, R_) k_g2p3_tile<F_, M_, R_><<<a.n_chunks, 128, 0, st>>>(a)
  if (mig) {
    if (flip) { if (resort) MPM_G3T(true, true, true); else MPM_G3T(true, true, false); }
    else      { if (resort) MPM_G3T(true, false, true); else MPM_G3T(true, false, false); }
  } else {
    if (flip) { if (resort) MPM_G3T(false, true, true); else MPM_G3T(false, true, false); }
    else      { if (resort) MPM_G3T(false, false, true); else MPM_G3T(false, false, false); }
  }
#undef MPM_G3T
}

}  // namespace mpm

namespace mpm {
int substep3d_chunk_capacity() { return 512; }
void launch_substep3d(const Substep3dArgs &a, bool flip, bool mig, bool resort, cudaStream_t st) {
  const int grid = a.n_chunks;
  if (grid <= 0) return;
#define MPM_S3D(F_, M_, R_) k_substep3d<F_, M_, R_><<<grid, 128, 0, st>>>(a)
  if (flip) {
    if (mig) { if (resort) MPM_S3D(true, true, true); else MPM_S3D(true, true, false); }
    else     { if (resort) MPM_S3D(true, false, true); else MPM_S3D(true, false, false); }
  } else {
    if (mig) { if (resort) MPM_S3D(false, true, true); else MPM_S3D(false, true, false); }
    else     { if (resort) MPM_S3D(false, false, true); else MPM_S3D(false, false, false); }
  }
#undef MPM_S3D
}
}  // namespace mpm
