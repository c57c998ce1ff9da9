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

template <bool MIG, bool FLIP, bool RESORT>
__global__ void __launch_bounds__(128, MPM_G2P3_FAST_MINB) k_g2p3(const __grid_constant__ G2p3Args A) {
  const Params &P = A.P;
  const long long i = A.first + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long n = A.n;
  if (MIG && A.dev_n && n > *A.dev_n) n = *A.dev_n;  // x-slab handles: exact extent on the device
  float vmax = 0.0f;
  if (i < n) {
    const float4 xj = A.s.xj[i], vm = A.s.vm[i];
    const int mat_id = __float_as_int(vm.w);
    if (!(MIG && mat_id == DEAD)) {  // slot of a particle that emigrated earlier: dropped by the re-sort
      Mat<3> F;
#pragma unroll
      for (int c = 0; c < 3; c++)
#pragma unroll
        for (int r = 0; r < 3; r++) F.d[c][r] = A.s.F[c * 3 + r][i];
      const SoA<3> &out = RESORT ? A.d : A.s;
      long long dst = i;
      if (RESORT) dst = (long long)A.new_start[A.key[i]] + A.rank[i];
      float x[3] = {xj.x, xj.y, xj.z};
      float Jp = xj.w;
      const Material &mat = P.mat[material_index(P, mat_id)];
      {
        // ---- gather (:136-156), advect (:159), F update (:162) ----
        Stencil<3> st = make_stencil<3>(x, P.inv_dx);
        clamp_base<3>(P, st.base);  // G2P never flags (P2G did)
        float v[3], dv[3] = {0.0f, 0.0f, 0.0f};
        Mat<3> C;
        gather3_fast(P, st, A.grid, A.vold, FLIP, v, C, dv);
        const float s4 = 4 * P.inv_dx;  // the constant of :154, applied once
#pragma unroll
        for (int cc = 0; cc < 3; cc++)
#pragma unroll
          for (int r = 0; r < 3; r++) C.d[cc][r] = s4 * C.d[cc][r];
        vmax = fmaxf(fabsf(v[0]), fmaxf(fabsf(v[1]), fabsf(v[2])));
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
        if (MIG && slot >= 0) {
          // the record of an emigrant (layout of emigrate() in mpm_kernels.cu): x v F C Jp mat id pad; v and C are
          // read back from where this thread stored them
          const int sd = slot >> 30;
          float *r = (sd == 0 ? A.mig.send_lo : A.mig.send_hi) + (size_t)(slot & 0x3fffffff) * MigRec<3>::WORDS;
          const float4 vv = out.vm[dst];
          float rec[MigRec<3>::WORDS];
          rec[0] = x[0]; rec[1] = x[1]; rec[2] = x[2];
          rec[3] = vv.x; rec[4] = vv.y; rec[5] = vv.z;
#pragma unroll
          for (int c = 0; c < 3; c++)
#pragma unroll
            for (int k = 0; k < 3; k++) {
              rec[6 + c * 3 + k] = F.d[c][k];
              rec[15 + c * 3 + k] = out.C[c * 3 + k][dst];
            }
          rec[24] = Jp;
          rec[25] = __int_as_float(mat_id);
          rec[26] = __int_as_float(id);
          rec[27] = 0.0f;
#pragma unroll
          for (int k = 0; k < MigRec<3>::WORDS / 4; k++)
            reinterpret_cast<float4 *>(r)[k] = make_float4(rec[4 * k], rec[4 * k + 1], rec[4 * k + 2], rec[4 * k + 3]);
        }
      }
    }
  }
  if (A.stats) {  // largest displacement of this substep in cells: feeds the re-sort interval (see engine)
    const unsigned bits = __reduce_max_sync(0xffffffffu, __float_as_uint(vmax * A.dt * P.inv_dx));
    unsigned *slot = reinterpret_cast<unsigned *>(&A.stats[2]);
    if ((threadIdx.x & 31) == 0 && bits > *reinterpret_cast<volatile unsigned *>(slot)) atomicMax(slot, bits);
  }
}

}  // namespace

void launch_g2p3(const G2p3Args &a, bool flip, bool mig, bool resort, cudaStream_t st) {
  if (a.n - a.first <= 0) return;
  const unsigned blocks = (unsigned)((a.n - a.first + 127) / 128);
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

}  // namespace mpm
