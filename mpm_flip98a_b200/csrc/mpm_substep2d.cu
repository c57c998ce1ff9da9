// mpm_substep2d.cu -- the 2D substep kernel (default path, single GPU and x-slabs): G2P of substep n and
// P2G of substep n+1 in one pass over the particles, one CTA per chunk (<= 768 particles) of a non-empty bin of
// 8x8 cells.
//
// Reference statements (cpp_validation/mls-mpm88-explained.cpp): :134-179 (G2P: gather, advect, F update,
// SVD clamp, Jp) of the current substep, then :53-102 (P2G) of the next one on the state still in registers.
// Same algorithm as k_p2g_cells<2,...,FUSED=1> in mpm_kernels.cu (which stays as the MPM_FLAG_STRICT /
// 3D / stand-alone-P2G kernel); this file is its instruction diet for sm_100a -- the round-1 kernel was
// issue-bound (75 % of issue slots, 715 thread-instructions per particle, DRAM 37 % busy):
//   * packed fp32 pairs (FMUL2 / FADD2 / FFMA2, mpm_math2.cuh) for every 2-vector: positions, velocities,
//     matrix columns, per-axis weights, node accumulators;
//   * separable G2P gather: 27 packed FMAs instead of 81 scalar operations, node velocities loaded as 8 bytes;
//   * FLIP is a template parameter (a runtime alpha cost 50 predicated-off instructions per particle);
//   * 32-bit slab-local indices, one address computation per stencil row;
//   * the per-chunk scans run on all four warps (packed count|items scan), item descriptors are built once
//     per cell (no integer division per work item);
//   * records split into two float4 planes (half the shared-memory bank conflicts of 32-byte records);
//   * a work list entry per chunk, read with ONE load (the round-1 chain active bin -> bin range -> particles cost
//     three dependent memory latencies at every CTA start), dense bins spread over several CTAs.
// (Tried and removed: persistent CTAs walking the work list with a prefetch of the next chunk's first particles --
// 6.45 vs 6.15 ms on c4 at equal clocks; even compiled out, the loop structure cost 4 % more instructions.)
// RESORT = true additionally performs the storage re-sort on the fly: each particle's new state is written
// to its slot in the OTHER storage buffer (slot = new bin start + rank, both computed from the positions
// before this substep by k_count_rank), so a re-sort costs one 12-byte pass instead of a radix sort plus a
// 120-byte reorder.
#include "mpm_kernels.cuh"
#include "mpm_math2.cuh"

namespace mpm {

namespace {

#ifndef MPM_SUBSTEP2D_CAP
#define MPM_SUBSTEP2D_CAP 768
#endif
constexpr int B = 8, NT = 128, CAP = MPM_SUBSTEP2D_CAP, M = 1, L = B + 2 * M, NC = L * L;
constexpr int RM = 8;                       // a cell with more records is split evenly into ceil(n/RM) items
constexpr int MAXI = NC + CAP / RM + 1;     // work items per chunk, upper bound

#ifndef MPM_SUBSTEP2D_MINB
#define MPM_SUBSTEP2D_MINB 7
#endif

struct PS {  // particle state in registers, matrices as columns
  f2 x, v;
  M2c C, F;
  float Jp;
  int mat;
};

__device__ __forceinline__ void load_g2p2(const SoA<2> &s, int i, PS &p, bool need_v) {
  p.x = s.x[i];
  const float4 F = s.F[i];
  p.F.c0 = mk2(F.x, F.y);
  p.F.c1 = mk2(F.z, F.w);
  p.Jp = s.Jp[i];
  p.mat = s.mat[i];
  if (need_v) p.v = s.v[i];
}
__device__ __forceinline__ void store_state2(const SoA<2> &s, int i, const PS &p) {
  s.x[i] = p.x;
  s.v[i] = p.v;
  s.C[i] = make_float4(p.C.c0.x, p.C.c0.y, p.C.c1.x, p.C.c1.y);
  s.F[i] = make_float4(p.F.c0.x, p.F.c0.y, p.F.c1.x, p.F.c1.y);
  s.Jp[i] = p.Jp;
}

// x-slab runs: pack a particle whose new base column left the slab for the neighbour (record + id), see
// emigrate() in mpm_kernels.cu.  Rare: kept out of line.  The record travels BY VALUE (four float4 in registers): a
// reference to the caller's particle state would pin that state in local memory for the whole hot loop (measured:
// 14 STL per particle, +24 % kernel time on every x-slab handle).
__device__ __noinline__ bool emigrate2(const MigPtrs &mig, int side, float4 r0, float4 r1, float4 r2, float4 r3,
                                       int *__restrict__ status) {
  const int slot = atomicAdd(&mig.count[side], 1);
  if (slot >= mig.cap) {
    atomicOr(status, STATUS_MIGRATION_OVERFLOW);
    return false;
  }
  float4 *r = reinterpret_cast<float4 *>((side == 0 ? mig.send_lo : mig.send_hi) + (size_t)slot * MigRec<2>::WORDS);
  r[0] = r0;
  r[1] = r1;
  r[2] = r2;
  r[3] = r3;
  return true;
}

// a particle that drifted past the 1-cell bin margin since the last re-sort: plain per-particle scatter
// with the reference's exact association (:92-100)
__device__ __noinline__ void scatter_fallback(const Params &P, float4 *__restrict__ grid, int bx, int by, f2 fx, f2 mv,
                                              M2c A, unsigned char *__restrict__ touched, int tiles_y) {
  {  // the tiles of the 3x3 nodes (at most four)
    const int t0 = (bx - P.slab_lo) >> 3, t1 = (bx - P.slab_lo + 2) >> 3, u0 = by >> 3, u1 = (by + 2) >> 3;
    touched[t0 * tiles_y + u0] = 1;
    touched[t0 * tiles_y + u1] = 1;
    touched[t1 * tiles_y + u0] = 1;
    touched[t1 * tiles_y + u1] = 1;
  }
  Stencil<2> st;
  st.base[0] = bx; st.base[1] = by;
  st.fx[0] = fx.x; st.fx[1] = fx.y;
  f2 w[3];
  weights2(fx, w);
#pragma unroll
  for (int k = 0; k < 3; k++) { st.w[k][0] = w[k].x; st.w[k][1] = w[k].y; }
  const Mat<2> affine = to_mat(A);
  const float mvv[2] = {mv.x, mv.y};
#pragma unroll
  for (int a = 0; a < 3; a++)
#pragma unroll
    for (int b = 0; b < 3; b++) {
      float nv[3];
      p2g_node_value<2>(P, st, affine, mvv, a, b, 0, nv);
      atomicAdd(&grid[(bx - P.slab_lo + a) * P.n1 + by + b], make_float4(nv[0], nv[1], nv[2], 0.0f));
    }
}

}  // namespace


template <bool FLIP, bool MIG, bool RESORT>
__global__ void __launch_bounds__(NT, MPM_SUBSTEP2D_MINB) k_substep2d(const __grid_constant__ Substep2dArgs A) {
  __shared__ float4 recA[CAP];          // fx.x fx.y  m*v.x m*v.y
  __shared__ float4 recB[CAP];          // affine column 0 | column 1
  __shared__ unsigned cr[CAP];          // (local cell << 16) | rank in cell; 0xffffffff = not binned
  __shared__ unsigned short sorted[CAP];
  __shared__ int cnt[NC + 4];           // per-cell count, then record start
  __shared__ unsigned item[MAXI];       // cell | first record << 8 | length << 20
  __shared__ int wtot[4], n_items_sh;
  const Params &P = A.P;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int n1 = P.n1;
  const float s4 = 4 * P.inv_dx;
  const int x_lo = P.slab_lo, x_hi = min(P.slab_hi, P.n_grid - 1) - 1;  // clamp range of base x (clamp_base)
  unsigned n_fallback = 0;
  float vmax = 0.0f;  // fastest particle of this thread (max norm): feeds the re-sort interval (CFL), see engine
  const int4 work = A.chunks[blockIdx.x];  // one load: no bin -> range -> particle chain at CTA start
  const int c0 = work.y, m = work.z;
  const int ox = (work.w >> 16) * B + P.slab_lo - M, oy = (work.w & 0xffff) * B - M;  // global cell of local cell 0
  const int bin_xy = work.w;
  {
    if (tid < NC) cnt[tid] = 0;
    if (tid < 9) {
      // this CTA's REDs land on the nodes of cells [ox, ox+L) x [oy, oy+L): the 3x3 tiles around the bin's own
      const int ttx = (bin_xy >> 16) - 1 + tid / 3, tty = (bin_xy & 0xffff) - 1 + tid % 3;
      if (ttx >= 0 && ttx < A.tiles_x && tty >= 0 && tty < A.tiles_y) A.touched_out[ttx * A.tiles_y + tty] = 1;
    }
    __syncthreads();
    // ---------------- phase 1: thread per particle (G2P of this substep, P2G record of the next) -------------
    PS nxt;
    if (tid < m) load_g2p2(A.s, c0 + tid, nxt, FLIP);
    for (int i = tid; i < m; i += NT) {
      PS p = nxt;  // loads issued one iteration ago; the next particle's go out now (software pipeline)
      if (i + NT < m) load_g2p2(A.s, c0 + i + NT, nxt, FLIP);
      const int slot = c0 + i;
      int dst = slot;
      if (MIG && p.mat == DEAD) {  // slot of a particle that emigrated earlier; dropped by the re-sort
        cr[i] = 0xffffffffu;
        continue;
      }
      {
        // ---- G2P :134-179 ----
        Sten2 st = stencil2(p.x, P.inv_dx);
        int bx = max(x_lo, min(st.bx, x_hi)), by = max(0, min(st.by, P.n_grid - 2));  // G2P never flags (P2G did)
        if (RESORT) dst = A.new_start[A.key[slot]] + (int)A.rank[slot];  // slot in the new order (k_count_rank)
        const Material &mat = P.mat[material_index(P, p.mat)];
        f2 wd[3];
        wd[0] = mul2(st.w[0], sub2(sp2(0.0f), st.fx));
        wd[1] = mul2(st.w[1], sub2(sp2(1.0f), st.fx));
        wd[2] = mul2(st.w[2], sub2(sp2(2.0f), st.fx));
        Gather2 g;
        g.v = g.c0 = g.c1 = g.vo = sp2(0.0f);
        const int node = (bx - P.slab_lo) * n1 + by;
        const float4 *gp = A.grid_in + node;
#pragma unroll
        for (int a = 0; a < 3; a++) {
          const float2 *row = reinterpret_cast<const float2 *>(gp + a * n1);  // node (a, b) = row[2*b]
          const f2 g0 = __ldg(row), g1 = __ldg(row + 2), g2 = __ldg(row + 4);
          gather2_row(g, st, wd, a, g0, g1, g2);
          if (FLIP) {
            const float2 *ro = A.vold_in + node + a * n1;
            gather2_row_old(g, st, a, __ldg(ro), __ldg(ro + 1), __ldg(ro + 2));
          }
        }
        const f2 v_in = FLIP ? p.v : sp2(0.0f);
        p.v = g.v;
        p.C.c0 = mul2(sp2(s4), g.c0);  // the 4*inv_dx of :154, applied once
        p.C.c1 = mul2(sp2(s4), g.c1);
        g2p_finish2(P, mat, A.dt_g2p, p.x, p.v, p.C, p.F, p.Jp, v_in, sub2(g.v, g.vo));
        vmax = fmaxf(vmax, fmaxf(fabsf(g.v.x), fabsf(g.v.y)));  // advection uses the gathered velocity (:159)
      }
      // ---- where does it go: x-slab emigration, storage slot ----
      bool gone = false;
      if (MIG) {
        const int nbx = max(0, min(base_coord(p.x.x, P.inv_dx), P.n_grid - 2));
        if (A.mig.interior) {
          // overlapped schedule: this launch covers bins >= 2 bin columns from the cuts -- verified, not assumed
          if ((P.slab_lo > 0 && nbx < P.slab_lo + 2) || (P.slab_hi < P.n_grid && nbx + 4 > P.slab_hi))
            atomicOr(A.status, STATUS_CFL);
        } else {
          const int side = nbx < P.slab_lo ? 0 : (nbx >= P.slab_hi ? 1 : -1);
          if (side >= 0)
            gone = emigrate2(A.mig, side, make_float4(p.x.x, p.x.y, p.v.x, p.v.y),
                             make_float4(p.F.c0.x, p.F.c0.y, p.F.c1.x, p.F.c1.y),
                             make_float4(p.C.c0.x, p.C.c0.y, p.C.c1.x, p.C.c1.y),
                             make_float4(p.Jp, __int_as_float(p.mat), __int_as_float(A.s.id[slot]), 0.0f), A.status);
        }
      }
      const SoA<2> &out = RESORT ? A.d : A.s;
      if (gone) p.mat = DEAD;
      store_state2(out, dst, p);
      if (RESORT) {
        out.mat[dst] = p.mat;
        out.id[dst] = A.s.id[slot];
      } else if (gone) {
        out.mat[dst] = DEAD;
      }
      if (gone) {
        cr[i] = 0xffffffffu;
        continue;
      }
      {
        // ---- P2G record of the next substep :53-89 ----
        Sten2 st = stencil2(p.x, P.inv_dx);
        int bx = st.bx, by = st.by, bad = 0;
        if (bx < x_lo) { bx = x_lo; bad = STATUS_DOMAIN; }
        if (bx > x_hi) { bx = x_hi; bad = STATUS_DOMAIN; }
        if (by < 0) { by = 0; bad = STATUS_DOMAIN; }
        if (by > P.n_grid - 2) { by = P.n_grid - 2; bad = STATUS_DOMAIN; }
        if (bad) atomicOr(A.status, bad);
        const Material &mat = P.mat[material_index(P, p.mat)];
        const M2c aff = affine2(P, mat, A.dt_p2g, p.F, p.C, p.Jp);
        const f2 mv = mul2(sp2(P.mass_p), p.v);
        const int lx = bx - ox, ly = by - oy;
        if ((unsigned)lx < (unsigned)L && (unsigned)ly < (unsigned)L) {
          const int cell = lx * L + ly;
          const int r = atomicAdd(&cnt[cell], 1);
          cr[i] = ((unsigned)cell << 16) | (unsigned)r;
          recA[i] = make_float4(st.fx.x, st.fx.y, mv.x, mv.y);
          recB[i] = make_float4(aff.c0.x, aff.c0.y, aff.c1.x, aff.c1.y);
        } else {
          cr[i] = 0xffffffffu;
          n_fallback++;
          scatter_fallback(P, A.grid_out, bx, by, st.fx, mv, aff, A.touched_out, A.tiles_y);
        }
      }
    }
    __syncthreads();
    // ---------------- scans on all four warps: record starts and work items per cell ----------------
    {
      const int v = tid < NC ? cnt[tid] : 0;
      const int parts = (v + RM - 1) / RM;
      int inc = v | (parts << 16);  // both sums stay below 2^16 (<= CAP records, <= MAXI items)
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
      }
      if (lane == 31) wtot[wid] = inc;
      __syncthreads();
      int pre = 0;
#pragma unroll
      for (int w = 0; w < 3; w++)
        if (w < wid) pre += wtot[w];
      const int ex = pre + inc - (v | (parts << 16));
      const int start = ex & 0xffff, ifirst = ex >> 16;
      if (tid < NC) {
        cnt[tid] = start;
        if (parts > 0) {
          // even split; the common cases without an integer division
          const int per = parts == 1 ? v : (parts == 2 ? (v + 1) >> 1 : (v + parts - 1) / parts);
          for (int sub = 0; sub < parts; sub++) {
            const int n0 = start + sub * per, len = min(per, v - sub * per);
            item[ifirst + sub] = (unsigned)tid | ((unsigned)n0 << 8) | ((unsigned)len << 20);
          }
        }
      }
      if (tid == NT - 1) n_items_sh = (pre + inc) >> 16;  // total number of items
    }
    __syncthreads();
    const int n_items = n_items_sh;
    for (int i = tid; i < m; i += NT) {
      const unsigned c = cr[i];
      if (c != 0xffffffffu) sorted[cnt[c >> 16] + (c & 0xffffu)] = (unsigned short)i;
    }
    __syncthreads();
    // ---------------- phase 2: thread per work item, 3x3 node sums in registers, one vector RED per node -------
    for (int it = tid; it < n_items; it += NT) {
      const unsigned ds = item[it];
      const int cell = ds & 0xff, n0 = (ds >> 8) & 0xfff, len = ds >> 20;
      f2 acc[3][3];
      f2 m01[3];    // mass sums of nodes (a,0), (a,1)
      float m2[3];  // ... and (a,2)
#pragma unroll
      for (int a = 0; a < 3; a++) {
        m01[a] = sp2(0.0f);
        m2[a] = 0.0f;
#pragma unroll
        for (int b = 0; b < 3; b++) acc[a][b] = sp2(0.0f);
      }
      for (int jj = n0; jj < n0 + len; jj++) {
        const int i = sorted[jj];
        const float4 ra = recA[i], rb = recB[i];
        const f2 fx = mk2(ra.x, ra.y), mv = mk2(ra.z, ra.w);
        f2 w[3];
        weights2(fx, w);
        // separable form of :92-100: w_ab * (q + a*cs0 + b*cs1), cs_k = affine column k * dx, q = m*v - cs0*fx - cs1*fy
        const f2 cs0 = mul2(mk2(rb.x, rb.y), sp2(P.dx)), cs1 = mul2(mk2(rb.z, rb.w), sp2(P.dx));
        f2 q = fma2(sp2(-fx.x), cs0, mv);
        q = fma2(sp2(-fx.y), cs1, q);
        const f2 wy01 = mk2(w[0].y, w[1].y);
#pragma unroll
        for (int a = 0; a < 3; a++) {
          const f2 xa = a == 0 ? q : (a == 1 ? add2(q, cs0) : fma2(sp2(2.0f), cs0, q));
          const f2 w01 = mul2(sp2(w[a].x), wy01);
          const float w2 = w[a].x * w[2].y;
          acc[a][0] = fma2(sp2(w01.x), xa, acc[a][0]);
          acc[a][1] = fma2(sp2(w01.y), add2(xa, cs1), acc[a][1]);
          acc[a][2] = fma2(sp2(w2), fma2(sp2(2.0f), cs1, xa), acc[a][2]);
          m01[a] = fma2(w01, sp2(P.mass_p), m01[a]);
          m2[a] = __fmaf_rn(w2, P.mass_p, m2[a]);
        }
      }
      const int lx = cell / L, ly = cell - lx * L;
      float4 *gp = A.grid_out + (ox + lx - P.slab_lo) * n1 + (oy + ly);
#pragma unroll
      for (int a = 0; a < 3; a++) {
        atomicAdd(gp + a * n1 + 0, make_float4(acc[a][0].x, acc[a][0].y, m01[a].x, 0.0f));  // RED.E.ADD.F32x4
        atomicAdd(gp + a * n1 + 1, make_float4(acc[a][1].x, acc[a][1].y, m01[a].y, 0.0f));
        atomicAdd(gp + a * n1 + 2, make_float4(acc[a][2].x, acc[a][2].y, m2[a], 0.0f));
      }
    }
  }
  if (A.stats) {
    if (n_fallback) {
      atomicAdd(&A.stats[0], (unsigned long long)n_fallback);
      atomicAdd(&A.stats[1], (unsigned long long)n_fallback);
    }
    // largest displacement of this substep in cells (non-negative floats order like their bit patterns); the
    // atomic is only issued by a warp that raises the maximum
    const unsigned bits = __reduce_max_sync(0xffffffffu, __float_as_uint(vmax * A.dt_g2p * P.inv_dx));
    unsigned *slot = reinterpret_cast<unsigned *>(&A.stats[2]);
    if (lane == 0 && bits > *reinterpret_cast<volatile unsigned *>(slot)) atomicMax(slot, bits);
  }
}

int substep2d_chunk_capacity() { return CAP; }

void launch_substep2d(const Substep2dArgs &a, bool flip, bool mig, bool resort, cudaStream_t st) {
  const int grid = a.n_chunks;
  if (grid <= 0) return;
#define MPM_S2D(F_, M_, R_) k_substep2d<F_, M_, R_><<<grid, NT, 0, st>>>(a)
  if (flip) {
    if (mig) { if (resort) MPM_S2D(true, true, true); else MPM_S2D(true, true, false); }
    else     { if (resort) MPM_S2D(true, false, true); else MPM_S2D(true, false, false); }
  } else {
    if (mig) { if (resort) MPM_S2D(false, true, true); else MPM_S2D(false, true, false); }
    else     { if (resort) MPM_S2D(false, false, true); else MPM_S2D(false, false, false); }
  }
#undef MPM_S2D
}

}  // namespace mpm
