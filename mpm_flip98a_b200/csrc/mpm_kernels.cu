// mpm_kernels.cu -- substep kernels, sm_100a.
//
// Reference statements each kernel restates (cpp_validation/mls-mpm88-explained.cpp):
//   k_p2g_naive, k_p2g_cells<FUSED=0>   :53-102   particle -> grid scatter (fused APIC momentum + MLS-MPM stress, :86-89)
//   k_grid_update                       :105-131  normalise by mass, gravity, sticky / separating boundaries
//   k_g2p_naive, k_g2p_bins             :134-179  gather v and C, advect, F update, SVD plasticity clamp, Jp
//   k_p2g_cells<FUSED=1>                :134-179 of substep n, then :53-102 of substep n+1 on the state still in
//                                       registers -- the default 2D substep kernel
// The grid reset (:50) is a cudaMemsetAsync issued by the engine.  All arithmetic lives in mpm_math.cuh and is
// shared with the CPU-side bitwise check (tests/host_check.cpp); this file is about data movement:
// SoA streams, the per-bin shared-memory cell sort, register accumulation, vector REDs, emigrant packing.
#include "mpm_kernels.cuh"
#include "mpm_math2.cuh"
#include "mpm_gather3.cuh"

namespace mpm {

// ------------------------------------------------------------------------------------------------
// grid update, one thread per node.  float4 in, float4 out (in place).
// ------------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(256) k_grid_update(Params P, float dt, float4 *__restrict__ grid,
                                                     void *__restrict__ vold_, long long nodes) {
  long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= nodes) return;
  // node coordinates (global): n = ((i - lo)*n1 + j)[*n1 + k]
  int k = 0, j, i;
  long long r = n;
  if (D == 3) {
    k = (int)(r % P.n1);
    r /= P.n1;
  }
  j = (int)(r % P.n1);
  i = (int)(r / P.n1) + P.slab_lo;
  float4 g4 = grid[n];
  float g[4] = {g4.x, g4.y, g4.z, g4.w};
  const bool flip = P.alpha != 0.0f;
  float vo[3];
  if (grid_node_update<D>(P, dt, i, j, k, g, vo)) grid[n] = make_float4(g[0], g[1], g[2], g[3]);
  // nodes with m == 0 keep their memset zeros (:50)
  if (flip) {
    if (D == 2) ((float2 *)vold_)[n] = make_float2(vo[0], vo[1]);
    else ((float4 *)vold_)[n] = make_float4(vo[0], vo[1], vo[2], 0.0f);
  }
}

// ------------------------------------------------------------------------------------------------
// 2D default path: grid update of one buffer AND reset of the other, restricted to the 8x8-node tiles that were
// actually written.  Every kernel that scatters into a grid buffer marks the tiles it can touch in that buffer's
// byte map (`touched`); a tile nobody marked holds zeros and needs neither the update (:109 skips empty nodes
// anyway) nor the reset.  One launch replaces the memset of the whole grid (:50) plus the update pass over all
// nodes: on the 8192^2 pool scene 58 % of the tiles are never touched.
//   upd / t_upd : buffer that holds this substep's P2G sums -> (v, 1|0) in place; its marks stay (it is the buffer
//                 that will be reset in the next substep)
//   clr / t_clr : buffer the NEXT P2G will scatter into -> zeros, marks cleared
// CTA = 8 node columns x 128 nodes (16 tiles), a warp per column: 512-byte coalesced segments.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_grid_tiles(Params P, float dt, float4 *__restrict__ upd, float2 *__restrict__ vold,
                                                    const unsigned char *__restrict__ t_upd, float4 *__restrict__ clr,
                                                    unsigned char *__restrict__ t_clr, int tiles_y,
                                                    const unsigned long long *__restrict__ stats, int guard_steps,
                                                    int *__restrict__ status) {
  const int tx = blockIdx.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (guard_steps > 0 && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
    // overlapped slab schedule: the interior launch runs WITHOUT the per-particle migration code; what makes that safe
    // is checked here, once per substep: no particle can have travelled the >= 14 cells between an interior bin and the
    // slab cut while (substeps since the re-sort + 2) x (largest per-substep displacement measured since) <= 8 cells
    const float d = __uint_as_float((unsigned)(stats[2] & 0xffffffffull));
    if (d * (float)(guard_steps + 2) > 8.0f) atomicOr(status, STATUS_CFL);
  }
  const int i = tx * 8 + w;  // local node column
  const int j0 = blockIdx.y * 128;
  const bool flip = P.alpha != 0.0f;
  unsigned char fu[4], fc[4];
#pragma unroll
  for (int r = 0; r < 4; r++) {
    const int ty = (j0 + r * 32) / 8 + lane / 8;
    const bool in = ty < tiles_y;
    fu[r] = in ? t_upd[tx * tiles_y + ty] : 0;
    fc[r] = in && clr ? t_clr[tx * tiles_y + ty] : 0;
  }
  __syncthreads();  // every flag of this CTA's tiles has been read before any is cleared
  if (i < P.ncol) {
#pragma unroll
    for (int r = 0; r < 4; r++) {
      const int j = j0 + r * 32 + lane;
      if (j >= P.n1) continue;
      const long long nd = (long long)i * P.n1 + j;
      if (fu[r]) {
        const float4 g4 = upd[nd];
        float g[4] = {g4.x, g4.y, g4.z, g4.w}, vo[3];
        if (grid_node_update<2>(P, dt, i + P.slab_lo, j, 0, g, vo)) upd[nd] = make_float4(g[0], g[1], g[2], g[3]);
        if (flip) vold[nd] = make_float2(vo[0], vo[1]);
      }
      if (fc[r]) clr[nd] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
  }
  if (clr && w == 0 && lane < 16) {
    const int ty = j0 / 8 + lane;
    if (ty < tiles_y) t_clr[tx * tiles_y + ty] = 0;
  }
}
void launch_grid_tiles(const Params &P, float dt, float4 *upd, void *vold, const unsigned char *t_upd, float4 *clr,
                       unsigned char *t_clr, int tiles_x, int tiles_y, cudaStream_t st, const unsigned long long *stats,
                       int guard_steps, int *status) {
  dim3 grid((unsigned)tiles_x, (unsigned)((P.n1 + 127) / 128));
  k_grid_tiles<<<grid, 256, 0, st>>>(P, dt, upd, (float2 *)vold, t_upd, clr, t_clr, tiles_y, stats, guard_steps, status);
}

template <int D>
void launch_grid_update(const Params &P, float dt, GridPtrs<D> g, cudaStream_t st) {
  unsigned blocks = (unsigned)((g.nodes + 255) / 256);
  k_grid_update<D><<<blocks, 256, 0, st>>>(P, dt, g.g, g.vold, g.nodes);
}
template void launch_grid_update<2>(const Params &, float, GridPtrs<2>, cudaStream_t);
template void launch_grid_update<3>(const Params &, float, GridPtrs<3>, cudaStream_t);

// ------------------------------------------------------------------------------------------------
// naive P2G: one thread per particle, 3^D vector REDs (RED.E.ADD.F32x4) into L2.
// ------------------------------------------------------------------------------------------------
// dev_n (x-slab handles): the exact storage extent lives on the device (immigrants are appended without the host
// learning their number); the host launches over an upper bound and the kernels stop at *dev_n
template <int D, bool MIG>
__global__ void __launch_bounds__(128) k_p2g_naive(Params P, float dt, SoA<D> s, long long first, long long n,
                                                   float4 *__restrict__ grid, int *__restrict__ status,
                                                   const int *__restrict__ dev_n) {
  long long i = first + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (MIG && dev_n && n > *dev_n) n = *dev_n;
  if (i >= n) return;
  PState<D> p;
  load_full(s, i, p);
  if (MIG && p.mat == DEAD) return;  // x-slab runs only: slot of an emigrated particle
  Stencil<D> st = make_stencil<D>(p.x, P.inv_dx);
  int bad = clamp_base<D>(P, st.base);
  if (bad) atomicOr(status, bad);
  const Material &mat = P.mat[material_index(P, p.mat)];
  Mat<D> affine = p2g_affine<D>(P, mat, dt, p.F, p.C, p.Jp);
  float mv[D];
#pragma unroll
  for (int c = 0; c < D; c++) mv[c] = P.mass_p * p.v[c];
#pragma unroll
  for (int a = 0; a < 3; a++)
#pragma unroll
    for (int b = 0; b < 3; b++)
#pragma unroll
      for (int c = 0; c < (D == 3 ? 3 : 1); c++) {
        float nv[D + 1];
        p2g_node_value<D>(P, st, affine, mv, a, b, c, nv);
        long long node = node_index<D>(P, st.base[0] + a, st.base[1] + b, D == 3 ? st.base[D - 1] + c : 0);
        float4 val = make_float4(nv[0], nv[1], nv[2], D == 3 ? nv[D] : 0.0f);
        atomicAdd(&grid[node], val);  // :97-100
      }
}

template <int D>
void launch_p2g_naive(const Params &P, float dt, const SoA<D> &s, long long first, long long n, GridPtrs<D> g,
                      int *status, cudaStream_t st, const int *dev_n) {
  if (n - first <= 0) return;
  unsigned blocks = (unsigned)((n - first + 127) / 128);
  if (P.multi) k_p2g_naive<D, true><<<blocks, 128, 0, st>>>(P, dt, s, first, n, g.g, status, dev_n);
  else k_p2g_naive<D, false><<<blocks, 128, 0, st>>>(P, dt, s, first, n, g.g, status, nullptr);
}
template void launch_p2g_naive<2>(const Params &, float, const SoA<2> &, long long, long long, GridPtrs<2>, int *,
                                  cudaStream_t, const int *);
template void launch_p2g_naive<3>(const Params &, float, const SoA<3> &, long long, long long, GridPtrs<3>, int *,
                                  cudaStream_t, const int *);

// ------------------------------------------------------------------------------------------------
// G2P building blocks (used by the G2P kernels below and by the fused G2P->P2G kernel)
// ------------------------------------------------------------------------------------------------
// One particle of G2P (:134-179).  `fetch(a, b, c, base, gv, vo)` delivers node velocity (and, for FLIP,
// its pre-gravity value) of stencil offset (a,b,c): straight from L2/L1 (naive kernel) or from the
// shared-memory tile of the particle's bin (binned kernel).
// x-slab runs: a particle whose NEW base cell (what the next P2G will use, :55) left [slab_lo, slab_hi) is
// packed (record + id) into the send buffer of that side and its slot marked dead.  Returns true if it left.
template <int D>
__device__ __forceinline__ bool emigrate(const Params &P, const SoA<D> &s, long long i, PState<D> &p, const MigPtrs &mig,
                                         int *__restrict__ status) {
  int bx = base_coord(p.x[0], P.inv_dx);
  bx = max(0, min(bx, P.n_grid - 2));
  const int side = bx < P.slab_lo ? 0 : (bx >= P.slab_hi ? 1 : -1);
  if (side < 0) return false;
  int slot = atomicAdd(&mig.count[side], 1);
  if (slot >= mig.cap) {
    atomicOr(status, STATUS_MIGRATION_OVERFLOW);  // stays here (and will be flagged out of slab)
    return false;
  }
  constexpr int W = MigRec<D>::WORDS;
  float *r = (side == 0 ? mig.send_lo : mig.send_hi) + (size_t)slot * W;
  float rec[W];
#pragma unroll
  for (int c = 0; c < D; c++) {
    rec[c] = p.x[c];
    rec[D + c] = p.v[c];
  }
#pragma unroll
  for (int c = 0; c < D; c++)
#pragma unroll
    for (int k = 0; k < D; k++) {
      rec[2 * D + c * D + k] = p.F.d[c][k];
      rec[2 * D + D * D + c * D + k] = p.C.d[c][k];
    }
  rec[2 * D + 2 * D * D] = p.Jp;
  rec[2 * D + 2 * D * D + 1] = __int_as_float(p.mat);
  rec[W - 2] = __int_as_float(s.id[i]);
  rec[W - 1] = 0.0f;
#pragma unroll
  for (int k = 0; k < W / 4; k++)
    reinterpret_cast<float4 *>(r)[k] = make_float4(rec[4 * k], rec[4 * k + 1], rec[4 * k + 2], rec[4 * k + 3]);
  p.mat = DEAD;
  mark_dead(s, i);
  store_state(s, i, p);  // 3D keeps the material id (now DEAD) in vm.w
  return true;
}

// loads particle i, gathers from the grid and returns the NEW state in p (nothing stored yet)
// FLIPC: -1 = alpha is tested at run time; 0 / 1 = compiled for alpha == 0 / != 0 (a run-time flag leaves the
// whole blend in the instruction stream, predicated off: ~6 issue slots per stencil node)
template <int D, bool FAST, typename Fetch, int FLIPC = -1>
__device__ __forceinline__ void g2p_update(const Params &P, float dt, const SoA<D> &s, long long i, const Fetch &fetch,
                                           PState<D> &p, bool loaded = false) {
  const bool flip = FLIPC < 0 ? (P.alpha != 0.0f) : (FLIPC != 0);
  if (!loaded) load_g2p(s, i, p, flip);
  Stencil<D> st = make_stencil<D>(p.x, P.inv_dx);
  clamp_base<D>(P, st.base);
  const Material &mat = P.mat[material_index(P, p.mat)];
  float v_in[D], dv[D], v[D];
#pragma unroll
  for (int c = 0; c < D; c++) {
    v_in[c] = flip ? p.v[c] : 0.0f;
    dv[c] = 0.0f;
    v[c] = 0.0f;  // :145
  }
  Mat<D> C = mat_zero<D>();  // :144
  constexpr bool FG = FAST;  // fast forms: C comes back without the constant 4*inv_dx of :154
  fetch.template gather<FG>(P, st, flip, v, C, dv);
  if (FG) {  // the 4*inv_dx of :154, applied once
    const float s4 = 4 * P.inv_dx;
#pragma unroll
    for (int cc = 0; cc < D; cc++)
#pragma unroll
      for (int r = 0; r < D; r++) C.d[cc][r] = s4 * C.d[cc][r];
  }
#pragma unroll
  for (int c = 0; c < D; c++) p.v[c] = v[c];
  p.C = C;
  g2p_finish<D>(P, mat, dt, p.x, p.v, p.C, p.F, p.Jp, v_in, dv);
}

// returns true when the particle was advanced and stored in place (false: dead slot or emigrated)
template <int D, bool MIG, bool FAST, typename Fetch, int FLIPC = -1>
__device__ __forceinline__ bool g2p_one(const Params &P, float dt, const SoA<D> &s, long long i, const Fetch &fetch,
                                        const MigPtrs &mig, int *__restrict__ status) {
  PState<D> p;
  g2p_update<D, FAST, Fetch, FLIPC>(P, dt, s, i, fetch, p);
  const bool dead = MIG && p.mat == DEAD;  // predicate after the loads were issued, not an early exit
  if (dead) return false;
  if (MIG && emigrate<D>(P, s, i, p, mig, status)) return false;
  store_state(s, i, p);
  return true;
}

// nodes straight from global memory through the read-only path
template <int D>
struct GlobalFetch {
  const float4 *__restrict__ grid;
  const void *__restrict__ vold_;
  template <bool FAST>
  __device__ __forceinline__ void gather(const Params &P, const Stencil<D> &st, bool flip, float *v, Mat<D> &C,
                                         float *dv) const {
    if constexpr (FAST && D == 3) {
      gather3_fast(P, st, grid, (const float4 *)vold_, flip, v, C, dv);
    } else {
#pragma unroll
      for (int a = 0; a < 3; a++)
#pragma unroll
        for (int b = 0; b < 3; b++)
#pragma unroll
          for (int c = 0; c < (D == 3 ? 3 : 1); c++) {
            long long node = node_index<D>(P, st.base[0] + a, st.base[1] + b, D == 3 ? st.base[D - 1] + c : 0);
            float4 g4 = __ldg(&grid[node]);
            float gv[3] = {g4.x, g4.y, g4.z};
            float vo[3] = {0.0f, 0.0f, 0.0f};
            if (flip) {
              if (D == 2) {
                float2 o = __ldg(&((const float2 *)vold_)[node]);
                vo[0] = o.x; vo[1] = o.y;
              } else {
                float4 o = __ldg(&((const float4 *)vold_)[node]);
                vo[0] = o.x; vo[1] = o.y; vo[2] = o.z;
              }
            }
            if (FAST) g2p_accumulate_fast<D>(st, a, b, c, gv, vo, flip, v, C, dv);
            else g2p_accumulate<D>(P, st, a, b, c, gv, vo, flip, v, C, dv);
          }
    }
  }
};

// ------------------------------------------------------------------------------------------------
// binned P2G ("cell gather"): one CTA per bin of B^D cells.  No floating-point atomics in shared
// memory (on sm_100a atomicAdd(float) on shared is a CAS loop, ATOMS.CAST.SPIN) and no shuffles:
//   phase 1  thread per particle (coalesced SoA loads): stencil, stress, affine (:55-89) -> a compact
//            record (fx, mass*v, affine) in shared memory; an integer counting sort by local cell
//            (native ATOMS.ADD) orders the records by the cell of their base node;
//   phase 2  thread per (cell[, stencil row]): walks its cell's records, forms the very same node
//            contributions as :92-100 (p2g_node_value) and sums them in REGISTERS -- every particle of
//            a cell hits the same 3^D nodes -- then issues one vector RED per (cell, node).
// Global atomics drop from 3^D per particle to 3^D per occupied cell.  Bin membership may be stale
// (storage is re-sorted every few substeps): the local cell grid carries a 1-cell margin and a
// particle that drifted further falls back to per-particle REDs, so the result never depends on it.
// ------------------------------------------------------------------------------------------------
template <int D>
struct CellRec;
template <>
struct CellRec<2> {
  float4 a, b;  // fx.x fx.y mv.x mv.y | A00 A10 A01 A11 (column-major d[c][r])
};
template <>
struct CellRec<3> {
  float4 a, b, c, d;  // fx.xyz mv.x | mv.y mv.z A[0][0] A[0][1] | A[0][2] A[1][0..2] | A[2][0..2] pad
};

template <int D>
__device__ __forceinline__ void pack_rec(CellRec<D> &r, const float *fx, const float *mv, const Mat<D> &A);
template <>
__device__ __forceinline__ void pack_rec<2>(CellRec<2> &r, const float *fx, const float *mv, const Mat<2> &A) {
  r.a = make_float4(fx[0], fx[1], mv[0], mv[1]);
  r.b = make_float4(A.d[0][0], A.d[0][1], A.d[1][0], A.d[1][1]);
}
template <>
__device__ __forceinline__ void pack_rec<3>(CellRec<3> &r, const float *fx, const float *mv, const Mat<3> &A) {
  r.a = make_float4(fx[0], fx[1], fx[2], mv[0]);
  r.b = make_float4(mv[1], mv[2], A.d[0][0], A.d[0][1]);
  r.c = make_float4(A.d[0][2], A.d[1][0], A.d[1][1], A.d[1][2]);
  r.d = make_float4(A.d[2][0], A.d[2][1], A.d[2][2], 0.0f);
}
template <int D>
__device__ __forceinline__ void unpack_rec(const CellRec<D> &r, float *fx, float *mv, Mat<D> &A);
template <>
__device__ __forceinline__ void unpack_rec<2>(const CellRec<2> &r, float *fx, float *mv, Mat<2> &A) {
  fx[0] = r.a.x; fx[1] = r.a.y; mv[0] = r.a.z; mv[1] = r.a.w;
  A.d[0][0] = r.b.x; A.d[0][1] = r.b.y; A.d[1][0] = r.b.z; A.d[1][1] = r.b.w;
}
template <>
__device__ __forceinline__ void unpack_rec<3>(const CellRec<3> &r, float *fx, float *mv, Mat<3> &A) {
  fx[0] = r.a.x; fx[1] = r.a.y; fx[2] = r.a.z; mv[0] = r.a.w;
  mv[1] = r.b.x; mv[2] = r.b.y; A.d[0][0] = r.b.z; A.d[0][1] = r.b.w;
  A.d[0][2] = r.c.x; A.d[1][0] = r.c.y; A.d[1][1] = r.c.z; A.d[1][2] = r.c.w;
  A.d[2][0] = r.d.x; A.d[2][1] = r.d.y; A.d[2][2] = r.d.z;
}

template <int D>
__device__ __forceinline__ void red_node(const Params &P, float4 *grid, int i, int j, int k, const float *nv) {
  long long node = node_index<D>(P, i, j, k);
  float4 val = make_float4(nv[0], nv[1], nv[2], D == 3 ? nv[D] : 0.0f);
  atomicAdd(&grid[node], val);  // RED.E.ADD.F32x4
}

// B = bin edge in cells, NT = threads, CAP = records per shared-memory chunk,
// TPC = threads per cell in phase 2 (1: all 3^D nodes; 3: one stencil row `a` each)
#ifndef MPM_P2G_MINB
#define MPM_P2G_MINB 8
#endif
#ifndef MPM_P2G_MINB3
#define MPM_P2G_MINB3 6
#endif
#ifndef MPM_P2G3_PREFETCH
#define MPM_P2G3_PREFETCH 0  // measured on c5: 1.83 ms with the prefetch pipeline (68 bytes of spills at 80 registers) vs 1.73 ms
#endif
#ifndef MPM_FUSED_MINB
#define MPM_FUSED_MINB 7
#endif
// FUSED: phase 1 first runs G2P of the CURRENT substep on the particle (gather from `grid_in`, the
// updated grid; new state stored in place) and then forms the P2G record of the NEXT substep from the
// state it still holds in registers: each particle is read once and written once per substep
// (84 B instead of 140 B of HBM traffic) and P2G never waits on particle loads.
template <int D, int B, int NT, int CAP, int TPC, bool FAST, bool MIG, bool FUSED>
__global__ void __launch_bounds__(NT, FUSED ? MPM_FUSED_MINB : (D == 3 ? MPM_P2G_MINB3 : MPM_P2G_MINB))
    k_p2g_cells(Params P, BinGeom G, float dt, SoA<D> s, const int *__restrict__ bin_start, float4 *__restrict__ grid,
                int *__restrict__ status, unsigned long long *__restrict__ stats, const float4 *__restrict__ grid_in,
                const void *__restrict__ vold_in, float dt_g2p, MigPtrs mig) {
  constexpr int M = 1, L = B + 2 * M;
  constexpr int NC = D == 2 ? L * L : L * L * L;
  // records as float4 PLANES (a | b [| c | d]): lane i of a warp writes 16 bytes at 16*i -- conflict-free -- and the
  // phase-2 reads of a dozen scattered records spread over 8 bank groups instead of 2 (the 64-byte AoS records cost
  // 16 wavefronts per LDS.128: the 3D kernel ran at 77 % of the L1 data-pipe peak, profiles/r02_ncu_full_c5.md)
  constexpr int NPL = D == 2 ? 2 : 4;
  __shared__ float4 recp[NPL * CAP];
  __shared__ unsigned short cell_of[CAP], rank_of[CAP], sorted[CAP];
  __shared__ int cnt[NC + 1];
  // phase-2 work items: (cell, first record, record count); a cell with more than RM records is
  // split evenly so that the lanes of a warp walk runs of similar length
  constexpr int RM = 8, MAXI = NC + CAP / RM + 1;
  __shared__ unsigned short item_first[NC + 1];   // per cell: index of its first item (exclusive scan); 16 bits keep
                                                  // the 3D kernel at 37.7 KB = SIX resident CTAs per SM instead of five
  __shared__ unsigned short item_cell[MAXI];
  __shared__ int wsum[4];
  const int tid = threadIdx.x;
  const int bin = G.active ? G.active[blockIdx.x] : (int)blockIdx.x;
  const int s0 = bin_start[bin], s1 = bin_start[bin + 1];
  if (s0 >= s1) return;
  // bin coordinates -> origin (global cell index of local cell 0, i.e. bin origin minus the margin)
  int o[3] = {0, 0, 0};
  {
    int r = bin;
    if (D == 3) { o[2] = (r % G.nb[2]) * B - M; r /= G.nb[2]; }
    o[1] = (r % G.nb[1]) * B - M;
    o[0] = (r / G.nb[1]) * B + P.slab_lo - M;
  }
  unsigned n_fallback = 0;
  for (int c0 = s0; c0 < s1; c0 += CAP) {
    const int m = min(CAP, s1 - c0);
    for (int k = tid; k <= NC; k += NT) cnt[k] = 0;
    __syncthreads();
    // ---- phase 1: thread per particle ----
    // FUSED: software pipeline -- the loads of a thread's NEXT particle are issued before it computes the
    // current one (measured on c4: 8.65 -> 7.73 ms; memory latency was the top stall reason)
    PState<D> nxt;
    if (FUSED && tid < m) load_g2p(s, (long long)c0 + tid, nxt, P.alpha != 0.0f);
    // 3D stand-alone P2G: the same software pipeline -- position, velocity / material and F of the thread's next
    // particle are in flight while the current one runs its Newton polar; C is only needed at the very end of the
    // record (:89) and is loaded per particle without anyone waiting for it
    constexpr bool PF3 = MPM_P2G3_PREFETCH && !FUSED && D == 3;
    if (PF3 && tid < m) load_g2p(s, (long long)c0 + tid, nxt, true);
    for (int i = tid; i < m; i += NT) {
      PState<D> p;
      if (FUSED) {
        GlobalFetch<D> fetch{grid_in, vold_in};
        p = nxt;  // this particle's loads were issued one iteration ago; the next one's go out now
        if (i + NT < m) load_g2p(s, (long long)c0 + i + NT, nxt, P.alpha != 0.0f);
        g2p_update<D, FAST>(P, dt_g2p, s, (long long)c0 + i, fetch, p, true);  // :134-179 of this substep
        if (MIG) {
          // dead slot: nothing to store; leaving the slab: packed for the neighbour, no P2G here (the
          // receiving handle scatters it when it arrives)
          if (p.mat != DEAD) {
            if (mig.interior) {
              // overlapped schedule: this launch runs while the slab exchanges its boundary; its bins lie
              // >= 2 bin columns from the cuts, so nothing here can leave the slab or reach the shared node
              // columns -- verified, not assumed
              const int bx = base_coord(p.x[0], P.inv_dx);
              if ((P.slab_lo > 0 && bx < P.slab_lo + 2) || (P.slab_hi < P.n_grid && bx + 4 > P.slab_hi))
                atomicOr(status, STATUS_CFL);
              store_state(s, (long long)c0 + i, p);
            } else if (!emigrate<D>(P, s, (long long)c0 + i, p, mig, status)) {
              store_state(s, (long long)c0 + i, p);
            }
          }
        } else {
          store_state(s, (long long)c0 + i, p);
        }
      } else if constexpr (PF3) {
        p = nxt;
        if (i + NT < m) load_g2p(s, (long long)c0 + i + NT, nxt, true);
#pragma unroll
        for (int c = 0; c < D; c++)
#pragma unroll
          for (int r = 0; r < D; r++) p.C.d[c][r] = s.C[c * D + r][(long long)c0 + i];
      } else {
        load_full(s, (long long)c0 + i, p);
      }
      if (MIG && p.mat == DEAD) {  // x-slab runs only: emigrated (now or since the last re-sort)
        cell_of[i] = 0xffffu;
        continue;
      }
      Stencil<D> st = make_stencil<D>(p.x, P.inv_dx);
      int bad = clamp_base<D>(P, st.base);
      if (bad) atomicOr(status, bad);
      const Material &mat = P.mat[material_index(P, p.mat)];
      Mat<D> affine = p2g_affine<D>(P, mat, dt, p.F, p.C, p.Jp);
      float mv[D];
#pragma unroll
      for (int c = 0; c < D; c++) mv[c] = P.mass_p * p.v[c];
      int l[3] = {st.base[0] - o[0], st.base[1] - o[1], D == 3 ? st.base[D - 1] - o[2] : 0};
      bool inside = (unsigned)l[0] < (unsigned)L && (unsigned)l[1] < (unsigned)L && (D == 2 || (unsigned)l[2] < (unsigned)L);
      if (inside) {
        int cell = l[0] * L + l[1];
        if (D == 3) cell = cell * L + l[2];
        int r = atomicAdd(&cnt[cell], 1);
        cell_of[i] = (unsigned short)cell;
        rank_of[i] = (unsigned short)r;
        if constexpr (FAST && D == 3) {
          // phase 2 works on the separable form w_abc * (q + a cs_0 + b cs_1 + c cs_2) (see there): form cs_k = affine
          // column k * dx and q = m v - sum_k fx_k cs_k ONCE here instead of in each of the item's three threads
#pragma unroll
          for (int k = 0; k < D; k++)
#pragma unroll
            for (int r = 0; r < D; r++) affine.d[k][r] = affine.d[k][r] * P.dx;
#pragma unroll
          for (int k = 0; k < D; k++)
#pragma unroll
            for (int r = 0; r < D; r++) mv[r] = __fmaf_rn(-st.fx[k], affine.d[k][r], mv[r]);
        }
        CellRec<D> rr;
        pack_rec<D>(rr, st.fx, mv, affine);
        recp[i] = rr.a;
        recp[CAP + i] = rr.b;
        if constexpr (D == 3) {
          recp[2 * CAP + i] = rr.c;
          recp[3 * CAP + i] = rr.d;
        }
      } else {  // drifted past the margin since the last re-sort: plain per-particle scatter
        cell_of[i] = 0xffffu;
        n_fallback++;
#pragma unroll
        for (int a = 0; a < 3; a++)
#pragma unroll
          for (int b = 0; b < 3; b++)
#pragma unroll
            for (int c = 0; c < (D == 3 ? 3 : 1); c++) {
              float nv[D + 1];
              p2g_node_value<D>(P, st, affine, mv, a, b, c, nv);
              red_node<D>(P, grid, st.base[0] + a, st.base[1] + b, D == 3 ? st.base[D - 1] + c : 0, nv);
            }
      }
    }
    __syncthreads();
    // ---- exclusive scans over the NC <= 2*NT cells on ALL warps (two cells per thread, count | items packed in one
    // word: both totals stay below 2^16): record starts and item starts.  (Round 1 scanned on warp 0 in 7 rounds
    // while the other warps waited at the barrier.) ----
    {
      static_assert(NC <= 2 * NT && NT == 128, "two cells per thread, four warps");
      const int lane = tid & 31, wid = tid >> 5;
      const int k0 = 2 * tid, k1 = 2 * tid + 1;
      const int v0 = k0 < NC ? cnt[k0] : 0, v1 = k1 < NC ? cnt[k1] : 0;
      const int p0 = v0 | (((v0 + RM - 1) / RM) << 16), p1 = v1 | (((v1 + RM - 1) / RM) << 16);
      int inc = p0 + p1;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
      }
      if (lane == 31) wsum[wid] = inc;
      __syncthreads();  // also: every cnt[] has been read before any is overwritten
      int pre = 0;
#pragma unroll
      for (int w = 0; w < 3; w++)
        if (w < wid) pre += wsum[w];
      const int ex = pre + inc - (p0 + p1);
      if (k0 < NC) {
        cnt[k0] = ex & 0xffff;
        item_first[k0] = (unsigned short)(ex >> 16);
      }
      if (k1 < NC) {
        cnt[k1] = (ex + p0) & 0xffff;
        item_first[k1] = (unsigned short)((ex + p0) >> 16);
      }
      if (tid == NT - 1) {
        cnt[NC] = (pre + inc) & 0xffff;
        item_first[NC] = (unsigned short)((pre + inc) >> 16);
      }
    }
    __syncthreads();
    for (int k = tid; k < NC; k += NT)
      for (int it = item_first[k]; it < item_first[k + 1]; it++) item_cell[it] = (unsigned short)k;
    for (int i = tid; i < m; i += NT) {
      unsigned c = cell_of[i];
      if (c != 0xffffu) sorted[cnt[c] + rank_of[i]] = (unsigned short)i;
    }
    __syncthreads();
    // ---- phase 2: thread per (cell, stencil row) ----
    const int n_items = item_first[NC];
    for (int item = tid; item < n_items * TPC; item += NT) {
      // TPC == 3: the three stencil rows of an item sit in ADJACENT lanes, so their reads of the item's records are
      // one shared-memory broadcast instead of three wavefronts from three warps (the 3D kernel was bound by the
      // L1 data pipe: 77 % of its wavefront peak, 260 M bank conflicts); `a` is a per-lane value, selected below
      const int it = TPC == 1 ? item : item / TPC;
      const int a_lo = TPC == 1 ? 0 : item % TPC, a_n = TPC == 1 ? 3 : 1;
      const int cell = item_cell[it];
      int n0, n1;
      {
        const int c0r = cnt[cell], k = cnt[cell + 1] - c0r;
        const int parts = item_first[cell + 1] - item_first[cell], sub = it - item_first[cell];
        const int per = (k + parts - 1) / parts;  // even split
        n0 = c0r + sub * per;
        n1 = min(n0 + per, c0r + k);
      }
      constexpr int NB = D == 3 ? 3 : 1;
      __align__(16) float acc[3][3][NB][D + 1];
#pragma unroll
      for (int a = 0; a < 3; a++)
#pragma unroll
        for (int b = 0; b < 3; b++)
#pragma unroll
          for (int c = 0; c < NB; c++)
#pragma unroll
            for (int q = 0; q <= D; q++) acc[a][b][c][q] = 0.0f;
      for (int jj = n0; jj < n1; jj++) {
        const int i = sorted[jj];
        Stencil<D> st;
        float mv[D];
        Mat<D> affine;
        {
          CellRec<D> rr;
          rr.a = recp[i];
          rr.b = recp[CAP + i];
          if constexpr (D == 3) {
            rr.c = recp[2 * CAP + i];
            rr.d = recp[3 * CAP + i];
          }
          unpack_rec<D>(rr, st.fx, mv, affine);
        }
#pragma unroll
        for (int k = 0; k < D; k++) {  // the same three expressions as make_stencil (:61-63)
          st.w[0][k] = 0.5f * ((1.5f - st.fx[k]) * (1.5f - st.fx[k]));
          st.w[1][k] = 0.75f - ((st.fx[k] - 1.0f) * (st.fx[k] - 1.0f));
          st.w[2][k] = 0.5f * ((st.fx[k] - 0.5f) * (st.fx[k] - 0.5f));
        }
        if constexpr (FAST && D == 3) {
          // 3D: the separable form below on packed pairs.  A node value is the 4-vector (m v + A dpos, m); with
          // Q = (q, mass_p) and CS_k = (affine column k * dx, 0) it is w_abc * (Q + a CS_0 + b CS_1 + c CS_2): every
          // step is two FFMA2 (xy | z,m) instead of four scalar operations.
          // (the record already holds cs_k in the affine slots and q in the m*v slots: phase 1)
          f2 cs_xy[3], cs_zm[3];
#pragma unroll
          for (int k = 0; k < 3; k++) {
            cs_xy[k] = mk2(affine.d[k][0], affine.d[k][1]);
            cs_zm[k] = mk2(affine.d[k][2], 0.0f);
          }
          const f2 q_xy = mk2(mv[0], mv[1]), q_zm = mk2(mv[2], P.mass_p);
          const int a = a_lo;  // TPC == 3: this thread's stencil row
          const float wa = a == 0 ? st.w[0][0] : (a == 1 ? st.w[1][0] : st.w[2][0]);
          const f2 xa_xy = fma2(sp2((float)a), cs_xy[0], q_xy), xa_zm = fma2(sp2((float)a), cs_zm[0], q_zm);
#pragma unroll
          for (int b = 0; b < 3; b++) {
            const float wab = wa * st.w[b][1];
            const f2 yb_xy = b == 0 ? xa_xy : (b == 1 ? add2(xa_xy, cs_xy[1]) : fma2(sp2(2.0f), cs_xy[1], xa_xy));
            const f2 yb_zm = b == 0 ? xa_zm : (b == 1 ? add2(xa_zm, cs_zm[1]) : fma2(sp2(2.0f), cs_zm[1], xa_zm));
#pragma unroll
            for (int c = 0; c < 3; c++) {
              const float wabc = wab * st.w[c][2];
              const f2 z_xy = c == 0 ? yb_xy : (c == 1 ? add2(yb_xy, cs_xy[2]) : fma2(sp2(2.0f), cs_xy[2], yb_xy));
              const f2 z_zm = c == 0 ? yb_zm : (c == 1 ? add2(yb_zm, cs_zm[2]) : fma2(sp2(2.0f), cs_zm[2], yb_zm));
              f2 *acc2 = reinterpret_cast<f2 *>(acc[0][b][c]);  // (x, y), (z, m): 16-byte aligned quadruples
              acc2[0] = fma2(sp2(wabc), z_xy, acc2[0]);
              acc2[1] = fma2(sp2(wabc), z_zm, acc2[1]);
            }
          }
        } else if (FAST) {
          // Separable form of :92-100 with explicit FMAs: w*(mv + A*((o - fx)*dx)) == w*(q + sum_k o_k*cs_k),
          // cs_k = A.col(k)*dx, q = mv - sum_k cs_k*fx_k.  Algebraically identical, ~2.5x fewer
          // instructions; rounding differs from the reference's association at the 1e-7 level
          // (the same size as its own summation-order noise).  MPM_FLAG_STRICT keeps the exact form.
          float cs[D][D], q[D];
#pragma unroll
          for (int k = 0; k < D; k++)
#pragma unroll
            for (int r = 0; r < D; r++) cs[k][r] = affine.d[k][r] * P.dx;
#pragma unroll
          for (int r = 0; r < D; r++) {
            q[r] = mv[r];
#pragma unroll
            for (int k = 0; k < D; k++) q[r] = __fmaf_rn(-cs[k][r], st.fx[k], q[r]);
          }
#pragma unroll
          for (int aa = 0; aa < 3; aa++) {
            if (aa >= a_n) break;
            const int a = TPC == 1 ? aa : a_lo;
            const float wa = TPC == 1 ? st.w[aa][0] : (a == 0 ? st.w[0][0] : (a == 1 ? st.w[1][0] : st.w[2][0]));
            float xa[D];
#pragma unroll
            for (int r = 0; r < D; r++) xa[r] = __fmaf_rn((float)a, cs[0][r], q[r]);
#pragma unroll
            for (int b = 0; b < 3; b++) {
              const float wab = wa * st.w[b][1];
              float yb[D];
#pragma unroll
              for (int r = 0; r < D; r++) yb[r] = __fmaf_rn((float)b, cs[1][r], xa[r]);
#pragma unroll
              for (int c = 0; c < NB; c++) {
                float wabc = wab, zc[D];
#pragma unroll
                for (int r = 0; r < D; r++) zc[r] = yb[r];
                if (D == 3) {
                  wabc = wab * st.w[c][D - 1];
#pragma unroll
                  for (int r = 0; r < D; r++) zc[r] = __fmaf_rn((float)c, cs[D - 1][r], yb[r]);
                }
#pragma unroll
                for (int r = 0; r < D; r++) acc[aa][b][c][r] = __fmaf_rn(wabc, zc[r], acc[aa][b][c][r]);
                acc[aa][b][c][D] = __fmaf_rn(wabc, P.mass_p, acc[aa][b][c][D]);
              }
            }
          }
        } else {
  #pragma unroll
          for (int aa = 0; aa < 3; aa++) {
            if (aa >= a_n) break;
            const int a = TPC == 1 ? aa : a_lo;
  #pragma unroll
            for (int b = 0; b < 3; b++)
  #pragma unroll
              for (int c = 0; c < NB; c++) {
                float nv[D + 1];
                if (TPC == 1) {
                  p2g_node_value<D>(P, st, affine, mv, aa, b, c, nv);
                } else {
                  // `a` is a per-lane runtime value: select its weight and shift fx instead of indexing
                  // registers dynamically.  Bit-exact: ((float)a - fx) == (0.0f - (fx - (float)a)) because
                  // round-to-nearest is symmetric, so dpos (:94) is the identical float.
                  Stencil<D> sa = st;
                  sa.w[0][0] = a == 0 ? st.w[0][0] : (a == 1 ? st.w[1][0] : st.w[2][0]);
                  sa.fx[0] = st.fx[0] - (float)a;
                  p2g_node_value<D>(P, sa, affine, mv, 0, b, c, nv);
                }
  #pragma unroll
                for (int q = 0; q <= D; q++) acc[aa][b][c][q] = acc[aa][b][c][q] + nv[q];
              }
          }
        }
      }
      // cell coordinates -> global node of stencil offset (0,0,0)
      int lc[3];
      {
        int r = cell;
        if (D == 3) { lc[2] = r % L; r /= L; } else lc[2] = 0;
        lc[1] = r % L;
        lc[0] = r / L;
      }
#pragma unroll
      for (int aa = 0; aa < 3; aa++) {
        if (aa >= a_n) break;
        const int a = TPC == 1 ? aa : a_lo;
#pragma unroll
        for (int b = 0; b < 3; b++)
#pragma unroll
          for (int c = 0; c < NB; c++)
            red_node<D>(P, grid, o[0] + lc[0] + a, o[1] + lc[1] + b, D == 3 ? o[2] + lc[2] + c : 0, acc[aa][b][c]);
      }
    }
    __syncthreads();
  }
  if (stats && n_fallback) {
    atomicAdd(&stats[0], (unsigned long long)n_fallback);
    atomicAdd(&stats[1], (unsigned long long)n_fallback);
  }
}

template <int D>
bool p2g_cells_supported(const BinGeom &G) {
  return G.edge == (D == 2 ? 8 : 4);
}
template bool p2g_cells_supported<2>(const BinGeom &);
template bool p2g_cells_supported<3>(const BinGeom &);

template <int D, bool FAST, bool MIG, bool FUSED>
static void launch_cells_variant(const Params &P, const BinGeom &G, float dt, const SoA<D> &s, const int *bin_start,
                                 float4 *grid, int *status, unsigned long long *stats, const float4 *grid_in,
                                 const void *vold_in, float dt_g2p, MigPtrs mig, cudaStream_t st) {
  if constexpr (FUSED && D == 3) {
    // never launched: the fused 3D kernel is k_substep3d (mpm_substep3d.cu); this generic one spills ~0.5 KB when
    // fused in 3D (measured 5.1 vs 3.7 ms), so it is not instantiated
    (void)P; (void)G; (void)dt; (void)s; (void)bin_start; (void)grid; (void)status; (void)stats; (void)grid_in;
    (void)vold_in; (void)dt_g2p; (void)mig; (void)st;
  } else if constexpr (D == 2)
    k_p2g_cells<2, 8, 128, 768, 1, FAST, MIG, FUSED><<<(G.active ? G.n_active : G.n_bins), 128, 0, st>>>(P, G, dt, s, bin_start, grid, status, stats,
                                                                            grid_in, vold_in, dt_g2p, mig);
  else
    k_p2g_cells<3, 4, 128, 512, 3, FAST, MIG, FUSED><<<(G.active ? G.n_active : G.n_bins), 128, 0, st>>>(P, G, dt, s, bin_start, grid, status, stats,
                                                                            grid_in, vold_in, dt_g2p, mig);
}

template <int D>
void launch_p2g_cells(const Params &P, const BinGeom &G, float dt, const SoA<D> &s, long long n, const int *bin_start,
                      GridPtrs<D> g, int *status, unsigned long long *stats, bool strict, cudaStream_t st) {
  if (n <= 0) return;
  if (strict) {
    if (P.multi) launch_cells_variant<D, false, true, false>(P, G, dt, s, bin_start, g.g, status, stats, nullptr, nullptr, 0.0f, MigPtrs{nullptr, nullptr, nullptr, 0, 0, 0}, st);
    else launch_cells_variant<D, false, false, false>(P, G, dt, s, bin_start, g.g, status, stats, nullptr, nullptr, 0.0f, MigPtrs{nullptr, nullptr, nullptr, 0, 0, 0}, st);
  } else {
    if (P.multi) launch_cells_variant<D, true, true, false>(P, G, dt, s, bin_start, g.g, status, stats, nullptr, nullptr, 0.0f, MigPtrs{nullptr, nullptr, nullptr, 0, 0, 0}, st);
    else launch_cells_variant<D, true, false, false>(P, G, dt, s, bin_start, g.g, status, stats, nullptr, nullptr, 0.0f, MigPtrs{nullptr, nullptr, nullptr, 0, 0, 0}, st);
  }
}
template void launch_p2g_cells<2>(const Params &, const BinGeom &, float, const SoA<2> &, long long, const int *,
                                  GridPtrs<2>, int *, unsigned long long *, bool, cudaStream_t);
template void launch_p2g_cells<3>(const Params &, const BinGeom &, float, const SoA<3> &, long long, const int *,
                                  GridPtrs<3>, int *, unsigned long long *, bool, cudaStream_t);

// fused G2P(dt_g2p, reading g_in) + P2G(dt_p2g, scattering into grid_out)
template <int D>
void launch_g2p2g(const Params &P, const BinGeom &G, float dt_g2p, float dt_p2g, const SoA<D> &s, long long n,
                  const int *bin_start, GridPtrs<D> g_in, float4 *grid_out, int *status, unsigned long long *stats,
                  MigPtrs mig, bool strict, cudaStream_t st) {
  if (n <= 0) return;
#define MPM_FUSED(FAST_, MIG_) \
  launch_cells_variant<D, FAST_, MIG_, true>(P, G, dt_p2g, s, bin_start, grid_out, status, stats, g_in.g, g_in.vold, dt_g2p, mig, st)
  if (strict) {
    if (mig.enabled) MPM_FUSED(false, true);
    else MPM_FUSED(false, false);
  } else {
    if (mig.enabled) MPM_FUSED(true, true);
    else MPM_FUSED(true, false);
  }
#undef MPM_FUSED
}
template void launch_g2p2g<2>(const Params &, const BinGeom &, float, float, const SoA<2> &, long long, const int *,
                              GridPtrs<2>, float4 *, int *, unsigned long long *, MigPtrs, bool, cudaStream_t);
template void launch_g2p2g<3>(const Params &, const BinGeom &, float, float, const SoA<3> &, long long, const int *,
                              GridPtrs<3>, float4 *, int *, unsigned long long *, MigPtrs, bool, cudaStream_t);

// ------------------------------------------------------------------------------------------------
// naive G2P: one thread per particle, 3^D node reads through the read-only path, in-place update.
// ------------------------------------------------------------------------------------------------
// NOTE: no min-blocks hint here on purpose: with one, ptxas front-loads all 3^D node loads (56 regs),
// which measured 14% slower on B200 than the interleaved schedule it picks without (48 regs).
#ifndef MPM_G2P3_MINB
#define MPM_G2P3_MINB 6
#endif
template <int D, bool MIG, bool FAST, bool FLIP>
__global__ void __launch_bounds__(128, (D == 3 && FAST) ? MPM_G2P3_MINB : 1) k_g2p_naive(Params P, float dt, SoA<D> s, long long first, long long n,
                                                   const float4 *__restrict__ grid, const void *__restrict__ vold_,
                                                   MigPtrs mig, int *__restrict__ status,
                                                   unsigned long long *__restrict__ stats, const int *__restrict__ dev_n) {
  long long i = first + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (MIG && dev_n && n > *dev_n) n = *dev_n;
  float disp = 0.0f;
  if (i < n) {
    GlobalFetch<D> fetch{grid, vold_};
    float x0[D];
    load_pos(s, i, x0);
    const bool moved = g2p_one<D, MIG, FAST, GlobalFetch<D>, FLIP ? 1 : 0>(P, dt, s, i, fetch, mig, status);
    if (moved) {
      float x1[D];
      load_pos(s, i, x1);  // just written by this thread
#pragma unroll
      for (int k = 0; k < D; k++) disp = fmaxf(disp, fabsf(x1[k] - x0[k]) * P.inv_dx);
    }
  }
  if (stats) {  // largest displacement of this substep in cells: feeds the re-sort interval (see engine)
    const unsigned bits = __reduce_max_sync(0xffffffffu, __float_as_uint(disp));
    unsigned *slot = reinterpret_cast<unsigned *>(&stats[2]);
    if ((threadIdx.x & 31) == 0 && bits > *reinterpret_cast<volatile unsigned *>(slot)) atomicMax(slot, bits);
  }
}

// ------------------------------------------------------------------------------------------------
// binned G2P: one CTA per bin stages the (B+4)^D node velocities its particles can touch (bin + 1-cell
// drift margin + stencil reach) in shared memory once, instead of 3^D global loads per particle.
// A particle that drifted past the margin since the last re-sort gathers from global memory.
// ------------------------------------------------------------------------------------------------
template <int D, int B, bool FLIP>
struct TileFetch {
  static constexpr int T = B + 4;
  const float4 *tile;   // (vx, vy, vold_x, vold_y) in 2D; (vx, vy, vz, -) in 3D
  const float4 *tile_o; // 3D FLIP only: (vold_x, vold_y, vold_z, -)
  int o[3];             // global coordinate of tile node (0,0,0)
  GlobalFetch<D> global;
  template <bool FAST>
  __device__ __forceinline__ void gather(const Params &P, const Stencil<D> &st, bool flip, float *v, Mat<D> &C,
                                         float *dv) const {
    const int l0 = st.base[0] - o[0], l1 = st.base[1] - o[1], l2 = D == 3 ? st.base[D - 1] - o[2] : 0;
    const bool inside = (unsigned)l0 <= (unsigned)(T - 3) && (unsigned)l1 <= (unsigned)(T - 3) &&
                        (D == 2 || (unsigned)l2 <= (unsigned)(T - 3));
    if (!inside) {
      global.template gather<FAST>(P, st, flip, v, C, dv);
      return;
    }
    const int at = D == 2 ? l0 * T + l1 : (l0 * T + l1) * T + l2;
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
      for (int b = 0; b < 3; b++)
#pragma unroll
        for (int c = 0; c < (D == 3 ? 3 : 1); c++) {
          const int k = D == 2 ? at + a * T + b : at + (a * T + b) * T + c;
          float4 g4 = tile[k];
          float gv[3] = {g4.x, g4.y, g4.z};
          float vo[3] = {0.0f, 0.0f, 0.0f};
          if (FLIP) {
            if (D == 2) {
              vo[0] = g4.z; vo[1] = g4.w;
            } else {
              float4 o4 = tile_o[k];
              vo[0] = o4.x; vo[1] = o4.y; vo[2] = o4.z;
            }
          }
          if (FAST) g2p_accumulate_fast<D>(st, a, b, c, gv, vo, FLIP, v, C, dv);
          else g2p_accumulate<D>(P, st, a, b, c, gv, vo, FLIP, v, C, dv);
        }
  }
};

template <int D, int B, int NT, bool FLIP, bool MIG, bool FAST>
__global__ void __launch_bounds__(NT) k_g2p_bins(Params P, BinGeom G, float dt, SoA<D> s, const int *__restrict__ bin_start,
                                                 const float4 *__restrict__ grid, const void *__restrict__ vold_,
                                                 MigPtrs mig, int *__restrict__ status) {
  constexpr int T = B + 4, NN = D == 2 ? T * T : T * T * T;
  __shared__ float4 tile[NN];
  __shared__ float4 tile_o[(D == 3 && FLIP) ? NN : 1];
  const int bin = G.active ? G.active[blockIdx.x] : (int)blockIdx.x;
  const int s0 = bin_start[bin], s1 = bin_start[bin + 1];
  if (s0 >= s1) return;
  TileFetch<D, B, FLIP> fetch;
  {
    int r = bin;
    fetch.o[2] = 0;
    if (D == 3) { fetch.o[2] = (r % G.nb[2]) * B - 1; r /= G.nb[2]; }
    fetch.o[1] = (r % G.nb[1]) * B - 1;
    fetch.o[0] = (r / G.nb[1]) * B + P.slab_lo - 1;
  }
  fetch.tile = tile;
  fetch.tile_o = tile_o;
  fetch.global = GlobalFetch<D>{grid, vold_};
  for (int k = threadIdx.x; k < NN; k += NT) {
    int l[3];
    {
      int r = k;
      if (D == 3) { l[2] = r % T; r /= T; } else l[2] = 0;
      l[1] = r % T;
      l[0] = r / T;
    }
    const int gi = fetch.o[0] + l[0], gj = fetch.o[1] + l[1], gk = D == 3 ? fetch.o[2] + l[2] : 0;
    float4 g4 = make_float4(0.0f, 0.0f, 0.0f, 0.0f), o4 = g4;
    const bool ok = gi >= P.slab_lo && gi < P.slab_lo + P.ncol && gj >= 0 && gj < P.n1 && gk >= 0 && gk < P.n1;
    if (ok) {
      const long long node = node_index<D>(P, gi, gj, gk);
      g4 = __ldg(&grid[node]);
      if (FLIP) {
        if (D == 2) {
          float2 vo = __ldg(&((const float2 *)vold_)[node]);
          g4.z = vo.x; g4.w = vo.y;
        } else {
          o4 = __ldg(&((const float4 *)vold_)[node]);
        }
      }
    }
    tile[k] = g4;
    if (D == 3 && FLIP) tile_o[k] = o4;
  }
  __syncthreads();
  for (long long i = (long long)s0 + threadIdx.x; i < s1; i += NT) g2p_one<D, MIG, FAST>(P, dt, s, i, fetch, mig, status);
}

template <int D>
void launch_g2p_bins(const Params &P, const BinGeom &G, float dt, const SoA<D> &s, long long n, const int *bin_start,
                     GridPtrs<D> g, MigPtrs mig, int *status, bool strict, cudaStream_t st) {
  if (n <= 0) return;
  const bool flip = P.alpha != 0.0f;
  constexpr int B = D == 2 ? 8 : 4;
#define MPM_G2P_BINS(FLIP_, MIG_)                                                                                      \
  {                                                                                                                    \
    if (strict) k_g2p_bins<D, B, 128, FLIP_, MIG_, false><<<(G.active ? G.n_active : G.n_bins), 128, 0, st>>>(P, G, dt, s, bin_start, g.g, g.vold, mig, status); \
    else k_g2p_bins<D, B, 128, FLIP_, MIG_, true><<<(G.active ? G.n_active : G.n_bins), 128, 0, st>>>(P, G, dt, s, bin_start, g.g, g.vold, mig, status);         \
  }
  if (flip) {
    if (mig.enabled) MPM_G2P_BINS(true, true)
    else MPM_G2P_BINS(true, false)
  } else {
    if (mig.enabled) MPM_G2P_BINS(false, true)
    else MPM_G2P_BINS(false, false)
  }
#undef MPM_G2P_BINS
}
template void launch_g2p_bins<2>(const Params &, const BinGeom &, float, const SoA<2> &, long long, const int *,
                                 GridPtrs<2>, MigPtrs, int *, bool, cudaStream_t);
template void launch_g2p_bins<3>(const Params &, const BinGeom &, float, const SoA<3> &, long long, const int *,
                                 GridPtrs<3>, MigPtrs, int *, bool, cudaStream_t);

template <int D>
void launch_g2p_naive(const Params &P, float dt, const SoA<D> &s, long long first, long long n, GridPtrs<D> g,
                      MigPtrs mig, int *status, bool strict, cudaStream_t st, unsigned long long *stats, const int *dev_n) {
  if (n - first <= 0) return;
  unsigned blocks = (unsigned)((n - first + 127) / 128);
  const bool flip = P.alpha != 0.0f;
#define MPM_G2P_NAIVE(MIG_, FAST_)                                                                               \
  {                                                                                                              \
    if (flip) k_g2p_naive<D, MIG_, FAST_, true><<<blocks, 128, 0, st>>>(P, dt, s, first, n, g.g, g.vold, mig, status, stats, dev_n); \
    else k_g2p_naive<D, MIG_, FAST_, false><<<blocks, 128, 0, st>>>(P, dt, s, first, n, g.g, g.vold, mig, status, stats, dev_n);     \
  }
  if (mig.enabled) {
    if (strict) MPM_G2P_NAIVE(true, false)
    else MPM_G2P_NAIVE(true, true)
  } else {
    if (strict) MPM_G2P_NAIVE(false, false)
    else MPM_G2P_NAIVE(false, true)
  }
#undef MPM_G2P_NAIVE
}
template void launch_g2p_naive<2>(const Params &, float, const SoA<2> &, long long, long long, GridPtrs<2>, MigPtrs,
                                  int *, bool, cudaStream_t, unsigned long long *, const int *);
template void launch_g2p_naive<3>(const Params &, float, const SoA<3> &, long long, long long, GridPtrs<3>, MigPtrs,
                                  int *, bool, cudaStream_t, unsigned long long *, const int *);

// ------------------------------------------------------------------------------------------------
// x-slab exchange helpers: ghost-column sum, immigrant unpack
// ------------------------------------------------------------------------------------------------
__global__ void k_halo_add(float4 *__restrict__ dst, const float4 *__restrict__ src, long long count) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  float4 a = dst[i], b = src[i];
  // own partial + neighbour's partial: the neighbour computes b + a, the same float (commutative)
  dst[i] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}
void launch_halo_add(float4 *dst, const float4 *src, long long count, cudaStream_t st) {
  if (count <= 0) return;
  k_halo_add<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(dst, src, count);
}

template <int D>
__device__ __forceinline__ void unpack_mig_record(const float *__restrict__ r, PState<D> &p, int &id) {
  constexpr int W = MigRec<D>::WORDS;
#pragma unroll
  for (int c = 0; c < D; c++) {
    p.x[c] = r[c];
    p.v[c] = r[D + c];
  }
#pragma unroll
  for (int c = 0; c < D; c++)
#pragma unroll
    for (int k = 0; k < D; k++) {
      p.F.d[c][k] = r[2 * D + c * D + k];
      p.C.d[c][k] = r[2 * D + D * D + c * D + k];
    }
  p.Jp = r[2 * D + 2 * D * D];
  p.mat = __float_as_int(r[2 * D + 2 * D * D + 1]);
  id = __float_as_int(r[W - 2]);
}

// Arrivals of one exchange: the records of both neighbours' messages are appended behind the current storage extent.
// Their number is only known on the device (message headers), so is the extent (*ext); threads [0,K) serve the
// lower neighbour's message, [K,2K) the upper one's.
template <int D>
__global__ void k_immigrate(const float *__restrict__ recv_lo, const int *__restrict__ cnt_lo, const float *__restrict__ recv_hi,
                            const int *__restrict__ cnt_hi, int K, SoA<D> s, const int *__restrict__ ext, long long cap,
                            int *__restrict__ status) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int side = t >= K ? 1 : 0, idx = t - side * K;
  const int n_lo = cnt_lo ? min(*cnt_lo, K) : 0, n_hi = cnt_hi ? min(*cnt_hi, K) : 0;
  if (idx >= (side ? n_hi : n_lo)) return;
  const long long slot = (long long)ext[0] + (side ? n_lo : 0) + idx;
  if (slot >= cap) {
    atomicOr(status, STATUS_MIGRATION_OVERFLOW);  // no room: the particle is lost and the run is flagged
    return;
  }
  PState<D> p;
  int id;
  unpack_mig_record<D>((side ? recv_hi : recv_lo) + (size_t)idx * MigRec<D>::WORDS, p, id);
  store_state(s, slot, p);
  store_tags(s, slot, p.mat, id);
}

// P2G share of the particles in a migration message, restricted to the node columns [col_lo, col_hi): a particle
// that changes slab contributes to the shared columns on the SENDER's side (before its partial sums leave) and to
// the remaining columns on the RECEIVER's side, so one exchange per substep carries ghost sums and emigrants together.
// Exact association (:92-100); dt = the substep the scatter belongs to.
template <int D>
__global__ void k_scatter_records(Params P, float dt, const float *__restrict__ recs, const int *__restrict__ cnt, int K,
                                  float4 *__restrict__ grid, int col_lo, int col_hi, int *__restrict__ status) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= min(*cnt, K)) return;
  PState<D> p;
  int id;
  unpack_mig_record<D>(recs + (size_t)t * MigRec<D>::WORDS, p, id);
  Stencil<D> st = make_stencil<D>(p.x, P.inv_dx);
  bool bad = false;
#pragma unroll
  for (int k = 0; k < D; k++) {  // global clamp only: the base column of such a particle lies outside one of the two slabs
    if (st.base[k] < 0) { st.base[k] = 0; bad = true; }
    if (st.base[k] > P.n_grid - 2) { st.base[k] = P.n_grid - 2; bad = true; }
  }
  if (bad) atomicOr(status, STATUS_DOMAIN);
  const Material &mat = P.mat[material_index(P, p.mat)];
  const Mat<D> affine = p2g_affine<D>(P, mat, dt, p.F, p.C, p.Jp);
  float mv[D];
#pragma unroll
  for (int c = 0; c < D; c++) mv[c] = P.mass_p * p.v[c];
#pragma unroll
  for (int a = 0; a < 3; a++) {
    const int col = st.base[0] + a;
    if (col < col_lo || col >= col_hi) continue;
#pragma unroll
    for (int b = 0; b < 3; b++)
#pragma unroll
      for (int c = 0; c < (D == 3 ? 3 : 1); c++) {
        float nv[D + 1];
        p2g_node_value<D>(P, st, affine, mv, a, b, c, nv);
        red_node<D>(P, grid, col, st.base[1] + b, D == 3 ? st.base[D - 1] + c : 0, nv);
      }
  }
}

// extent / live-count bookkeeping of an x-slab handle, on the device: ext[0] = storage extent, ext[1] = live particles
__global__ void k_slab_counters(int *ext, const int *add_a, const int *add_b, const int *sub_a, const int *sub_b, int K,
                                long long cap, const int *set_extent) {
  if (threadIdx.x || blockIdx.x) return;
  int add = (add_a ? min(*add_a, K) : 0) + (add_b ? min(*add_b, K) : 0);
  int sub = (sub_a ? min(*sub_a, K) : 0) + (sub_b ? min(*sub_b, K) : 0);
  if (set_extent) ext[0] = *set_extent;
  long long e = (long long)ext[0] + add;
  ext[0] = (int)(e < cap ? e : cap);
  ext[1] += add - sub;
}

template <int D>
void launch_immigrate(const float *recv_lo, const int *cnt_lo, const float *recv_hi, const int *cnt_hi, int K,
                      const SoA<D> &s, const int *ext, long long cap, int *status, cudaStream_t st) {
  if (K <= 0 || (!cnt_lo && !cnt_hi)) return;
  k_immigrate<D><<<(unsigned)((2 * K + 255) / 256), 256, 0, st>>>(recv_lo, cnt_lo, recv_hi, cnt_hi, K, s, ext, cap, status);
}
template void launch_immigrate<2>(const float *, const int *, const float *, const int *, int, const SoA<2> &, const int *,
                                  long long, int *, cudaStream_t);
template void launch_immigrate<3>(const float *, const int *, const float *, const int *, int, const SoA<3> &, const int *,
                                  long long, int *, cudaStream_t);
template <int D>
void launch_scatter_records(const Params &P, float dt, const float *recs, const int *cnt, int K, float4 *grid, int col_lo,
                            int col_hi, int *status, cudaStream_t st) {
  if (K <= 0 || !cnt || col_lo >= col_hi) return;
  k_scatter_records<D><<<(unsigned)((K + 127) / 128), 128, 0, st>>>(P, dt, recs, cnt, K, grid, col_lo, col_hi, status);
}
template void launch_scatter_records<2>(const Params &, float, const float *, const int *, int, float4 *, int, int, int *,
                                        cudaStream_t);
template void launch_scatter_records<3>(const Params &, float, const float *, const int *, int, float4 *, int, int, int *,
                                        cudaStream_t);
void launch_slab_counters(int *ext, const int *add_a, const int *add_b, const int *sub_a, const int *sub_b, int K,
                          long long cap, const int *set_extent, cudaStream_t st) {
  k_slab_counters<<<1, 32, 0, st>>>(ext, add_a, add_b, sub_a, sub_b, K, cap, set_extent);
}

// ------------------------------------------------------------------------------------------------
// AoS (the reference's Particle record, :28-42: x v F C Jp c) <-> SoA
// ------------------------------------------------------------------------------------------------
template <int D>
__global__ void k_aos_to_soa(const float *__restrict__ aos, long long first, long long count, SoA<D> s,
                             const int *__restrict__ ids) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count) return;
  constexpr int W = 2 * D + 2 * D * D + 2;
  const float *r = aos + t * W;
  PState<D> p;
#pragma unroll
  for (int c = 0; c < D; c++) {
    p.x[c] = r[c];
    p.v[c] = r[D + c];
  }
#pragma unroll
  for (int c = 0; c < D; c++)
#pragma unroll
    for (int k = 0; k < D; k++) {
      p.F.d[c][k] = r[2 * D + c * D + k];
      p.C.d[c][k] = r[2 * D + D * D + c * D + k];
    }
  p.Jp = r[2 * D + 2 * D * D];
  p.mat = __float_as_int(r[2 * D + 2 * D * D + 1]);
  long long i = first + t;
  store_state(s, i, p);
  store_tags(s, i, p.mat, ids ? ids[t] : (int)i);
}
template <int D>
void launch_aos_to_soa(const float *aos, long long first, long long count, const SoA<D> &s, const int *ids,
                       cudaStream_t st) {
  if (count <= 0) return;
  k_aos_to_soa<D><<<(unsigned)((count + 255) / 256), 256, 0, st>>>(aos, first, count, s, ids);
}
template void launch_aos_to_soa<2>(const float *, long long, long long, const SoA<2> &, const int *, cudaStream_t);
template void launch_aos_to_soa<3>(const float *, long long, long long, const SoA<3> &, const int *, cudaStream_t);

template <int D>
__global__ void k_soa_to_aos(SoA<D> s, long long n, long long id0, long long count, float *__restrict__ aos,
                             int *__restrict__ ids_out) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  constexpr int W = 2 * D + 2 * D * D + 2;
  long long id = s.id[i];
  // the filters first: a record that is not wanted costs 8 bytes, not the whole state
  if (ids_out) {  // storage order: records [id0, id0+count) of the storage with their ids (-1 = dead slot)
    if (i < id0 || i >= id0 + count) return;
  } else if (id < id0 || id >= id0 + count || load_mat(s, i) == DEAD) {
    return;
  }
  PState<D> p;
  load_full(s, i, p);
  if (ids_out) {
    ids_out[i - id0] = p.mat == DEAD ? -1 : (int)id;
    id = i;
  }
  float *r = aos + (id - id0) * W;
#pragma unroll
  for (int c = 0; c < D; c++) {
    r[c] = p.x[c];
    r[D + c] = p.v[c];
  }
#pragma unroll
  for (int c = 0; c < D; c++)
#pragma unroll
    for (int k = 0; k < D; k++) {
      r[2 * D + c * D + k] = p.F.d[c][k];
      r[2 * D + D * D + c * D + k] = p.C.d[c][k];
    }
  r[2 * D + 2 * D * D] = p.Jp;
  r[2 * D + 2 * D * D + 1] = __int_as_float(p.mat);
}
template <int D>
void launch_soa_to_aos(const SoA<D> &s, long long n, long long id0, long long count, float *aos, int *ids_out,
                       cudaStream_t st) {
  if (n <= 0) return;
  k_soa_to_aos<D><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(s, n, id0, count, aos, ids_out);
}
template void launch_soa_to_aos<2>(const SoA<2> &, long long, long long, long long, float *, int *, cudaStream_t);
template void launch_soa_to_aos<3>(const SoA<3> &, long long, long long, long long, float *, int *, cudaStream_t);

template <int D>
__global__ void k_reorder(SoA<D> src, SoA<D> dst, const int *__restrict__ order, long long n) {
  long long slot = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= n) return;
  long long i = order[slot];
  PState<D> p;
  load_full(src, i, p);
  store_state(dst, slot, p);
  store_tags(dst, slot, p.mat, src.id[i]);
}
template <int D>
void launch_reorder(const SoA<D> &src, const SoA<D> &dst, const int *order, long long n, cudaStream_t st) {
  if (n <= 0) return;
  k_reorder<D><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(src, dst, order, n);
}
template void launch_reorder<2>(const SoA<2> &, const SoA<2> &, const int *, long long, cudaStream_t);
template void launch_reorder<3>(const SoA<3> &, const SoA<3> &, const int *, long long, cudaStream_t);

}  // namespace mpm
