// mpm_deterministic.cu -- MPM_FLAG_DETERMINISTIC: a P2G without atomics whose sums have a FIXED order.
//
// The reference's P2G (cpp_validation/mls-mpm88-explained.cpp:53-102) is a serial loop: every grid node receives
// its contributions in particle order, so the program is bit-reproducible.  The default GPU path sums with atomics
// in whatever order the hardware delivers.  This mode restores a definite order:
//   1. the storage is re-sorted by CELL with a stable sort before every substep (so the order of the particles is a
//      function of the simulation history only, not of thread scheduling);
//   2. k_det_records forms each particle's P2G record (:55-89: fx, m v, affine) once, with the reference's exact
//      association;
//   3. k_det_p2g runs one thread per grid NODE: it walks the 3^d cells whose particles can touch the node in x-major
//      cell order and, inside a cell, the particles in storage order, adding :97-100 sequentially from zero.
// That is exactly the order in which the reference's loop would meet these particles if it were handed the particle
// array in storage order -- so the grid after P2G is BITWISE what the CPU oracle computes from that array
// (tests/test_gpu_deterministic.py), total grid mass included, and two runs of the same input are bit-identical.
// Grid update and G2P have no reductions; they use the exact-association kernels.  Single-handle only (an x-slab cut
// would split a node's sum between two handles); a validation mode: ~10x slower than the default path.
#include "mpm_kernels.cuh"

namespace mpm {

// cell key (x-major over the global cell grid of base cells [0, n_grid-2]) of every slot
template <int D>
__global__ void k_det_cell_keys(Params P, SoA<D> s, long long n, unsigned *__restrict__ key, int *__restrict__ status) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float x[D];
  load_pos(s, i, x);
  int base[D];
#pragma unroll
  for (int k = 0; k < D; k++) base[k] = base_coord(x[k], P.inv_dx);
  const int bad = clamp_base<D>(P, base);
  if (bad) atomicOr(status, bad);
  const unsigned nc = (unsigned)(P.n_grid - 1);  // cells per axis that can be a base cell
  unsigned kk = (unsigned)base[0];
#pragma unroll
  for (int k = 1; k < D; k++) kk = kk * nc + (unsigned)base[k];
  key[i] = kk;
}

// per-slot P2G record, exact association: fx | m*v | affine (column-major)   -- 8 floats (2D) / 16 floats (3D)
template <int D>
__global__ void k_det_records(Params P, float dt, SoA<D> s, long long n, float *__restrict__ rec) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  constexpr int W = D == 2 ? 8 : 16;
  PState<D> p;
  load_full(s, i, p);
  Stencil<D> st = make_stencil<D>(p.x, P.inv_dx);
  const Material &mat = P.mat[material_index(P, p.mat)];
  const Mat<D> affine = p2g_affine<D>(P, mat, dt, p.F, p.C, p.Jp);
  float *r = rec + i * W;
#pragma unroll
  for (int k = 0; k < D; k++) {
    r[k] = st.fx[k];
    r[D + k] = P.mass_p * p.v[k];
  }
#pragma unroll
  for (int c = 0; c < D; c++)
#pragma unroll
    for (int k = 0; k < D; k++) r[2 * D + c * D + k] = affine.d[c][k];
}

// one thread per node: sequential sum, fixed order (see the header of this file)
template <int D>
__global__ void __launch_bounds__(128) k_det_p2g(Params P, const float *__restrict__ rec, const int *__restrict__ cell_start,
                                                 float4 *__restrict__ grid, long long nodes) {
  const long long nd = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (nd >= nodes) return;
  constexpr int W = D == 2 ? 8 : 16;
  int k = 0, j, i;
  long long r = nd;
  if (D == 3) {
    k = (int)(r % P.n1);
    r /= P.n1;
  }
  j = (int)(r % P.n1);
  i = (int)(r / P.n1);
  const int nc = P.n_grid - 1;
  float acc[D + 1];
#pragma unroll
  for (int q = 0; q <= D; q++) acc[q] = 0.0f;
  for (int cx = i - 2; cx <= i; cx++) {
    if (cx < 0 || cx >= nc) continue;
    for (int cy = j - 2; cy <= j; cy++) {
      if (cy < 0 || cy >= nc) continue;
      for (int cz = (D == 3 ? k - 2 : 0); cz <= (D == 3 ? k : 0); cz++) {
        if (D == 3 && (cz < 0 || cz >= nc)) continue;
        const long long cell = D == 2 ? (long long)cx * nc + cy : ((long long)cx * nc + cy) * nc + cz;
        const int p0 = cell_start[cell], p1 = cell_start[cell + 1];
        for (int p = p0; p < p1; p++) {
          const float *rr = rec + (long long)p * W;
          Stencil<D> st;
          float mv[D];
          Mat<D> affine;
#pragma unroll
          for (int q = 0; q < D; q++) {
            st.fx[q] = rr[q];
            mv[q] = rr[D + q];
            st.w[0][q] = 0.5f * ((1.5f - st.fx[q]) * (1.5f - st.fx[q]));    // :61-63
            st.w[1][q] = 0.75f - ((st.fx[q] - 1.0f) * (st.fx[q] - 1.0f));
            st.w[2][q] = 0.5f * ((st.fx[q] - 0.5f) * (st.fx[q] - 0.5f));
          }
#pragma unroll
          for (int c = 0; c < D; c++)
#pragma unroll
            for (int q = 0; q < D; q++) affine.d[c][q] = rr[2 * D + c * D + q];
          float nv[D + 1];
          // runtime stencil offsets: select the weight and shift fx instead of indexing registers dynamically;
          // ((float)a - fx) == (0.0f - (fx - (float)a)) exactly, so dpos (:94) is the identical float
          const int off[3] = {i - cx, j - cy, D == 3 ? k - cz : 0};
          Stencil<D> sa = st;
#pragma unroll
          for (int q = 0; q < D; q++) {
            const int a = off[q];
            sa.w[0][q] = a == 0 ? st.w[0][q] : (a == 1 ? st.w[1][q] : st.w[2][q]);
            sa.fx[q] = st.fx[q] - (float)a;
          }
          p2g_node_value<D>(P, sa, affine, mv, 0, 0, 0, nv);
#pragma unroll
          for (int q = 0; q <= D; q++) acc[q] = acc[q] + nv[q];  // :97-100, in order
        }
      }
    }
  }
  grid[nd] = make_float4(acc[0], acc[1], acc[2], D == 3 ? acc[D] : 0.0f);
}

template <int D>
void launch_det_cell_keys(const Params &P, const SoA<D> &s, long long n, unsigned *key, int *status, cudaStream_t st) {
  if (n <= 0) return;
  k_det_cell_keys<D><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(P, s, n, key, status);
}
template void launch_det_cell_keys<2>(const Params &, const SoA<2> &, long long, unsigned *, int *, cudaStream_t);
template void launch_det_cell_keys<3>(const Params &, const SoA<3> &, long long, unsigned *, int *, cudaStream_t);

template <int D>
void launch_det_p2g(const Params &P, float dt, const SoA<D> &s, long long n, float *rec, const int *cell_start,
                    float4 *grid, long long nodes, cudaStream_t st) {
  if (n > 0) k_det_records<D><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(P, dt, s, n, rec);
  k_det_p2g<D><<<(unsigned)((nodes + 127) / 128), 128, 0, st>>>(P, rec, cell_start, grid, nodes);
}
template void launch_det_p2g<2>(const Params &, float, const SoA<2> &, long long, float *, const int *, float4 *, long long,
                                cudaStream_t);
template void launch_det_p2g<3>(const Params &, float, const SoA<3> &, long long, float *, const int *, float4 *, long long,
                                cudaStream_t);

}  // namespace mpm
