// mpm_kernels.cu -- substep kernels, sm_100a.
//
// Reference statements each kernel restates (cpp_validation/mls-mpm88-explained.cpp):
//   k_p2g_*        :53-102   particle -> grid scatter (fused APIC momentum + MLS-MPM stress, :86-89)
//   k_grid_update  :105-131  normalise by mass, gravity, sticky / separating boundaries
//   k_g2p_*        :134-179  gather v and C, advect, F update, SVD plasticity clamp, Jp
// The grid reset (:50) is a cudaMemsetAsync issued by the engine.
#include "mpm_kernels.cuh"

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
template <int D>
__global__ void __launch_bounds__(128) k_p2g_naive(Params P, float dt, SoA<D> s, long long n, float4 *__restrict__ grid,
                                                   int *__restrict__ status) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  PState<D> p;
  load_full(s, i, p);
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
void launch_p2g_naive(const Params &P, float dt, const SoA<D> &s, long long n, GridPtrs<D> g, int *status,
                      cudaStream_t st) {
  if (n <= 0) return;
  unsigned blocks = (unsigned)((n + 127) / 128);
  k_p2g_naive<D><<<blocks, 128, 0, st>>>(P, dt, s, n, g.g, status);
}
template void launch_p2g_naive<2>(const Params &, float, const SoA<2> &, long long, GridPtrs<2>, int *, cudaStream_t);
template void launch_p2g_naive<3>(const Params &, float, const SoA<3> &, long long, GridPtrs<3>, int *, cudaStream_t);

// ------------------------------------------------------------------------------------------------
// naive G2P: one thread per particle, 3^D node reads through the read-only path, in-place update.
// ------------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(128) k_g2p_naive(Params P, float dt, SoA<D> s, long long n,
                                                   const float4 *__restrict__ grid, const void *__restrict__ vold_) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const bool flip = P.alpha != 0.0f;
  PState<D> p;
  load_g2p(s, i, p, flip);
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
        g2p_accumulate<D>(P, st, a, b, c, gv, vo, flip, v, C, dv);
      }
#pragma unroll
  for (int c = 0; c < D; c++) p.v[c] = v[c];
  p.C = C;
  g2p_finish<D>(P, mat, dt, p.x, p.v, p.C, p.F, p.Jp, v_in, dv);
  store_state(s, i, p);
}

template <int D>
void launch_g2p_naive(const Params &P, float dt, const SoA<D> &s, long long n, GridPtrs<D> g, cudaStream_t st) {
  if (n <= 0) return;
  unsigned blocks = (unsigned)((n + 127) / 128);
  k_g2p_naive<D><<<blocks, 128, 0, st>>>(P, dt, s, n, g.g, g.vold);
}
template void launch_g2p_naive<2>(const Params &, float, const SoA<2> &, long long, GridPtrs<2>, cudaStream_t);
template void launch_g2p_naive<3>(const Params &, float, const SoA<3> &, long long, GridPtrs<3>, cudaStream_t);

// ------------------------------------------------------------------------------------------------
// AoS (the reference's Particle record, :28-42: x v F C Jp c) <-> SoA
// ------------------------------------------------------------------------------------------------
template <int D>
__global__ void k_aos_to_soa(const float *__restrict__ aos, long long first, long long count, SoA<D> s) {
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
  store_tags(s, i, p.mat, (int)i);
}
template <int D>
void launch_aos_to_soa(const float *aos, long long first, long long count, const SoA<D> &s, cudaStream_t st) {
  if (count <= 0) return;
  k_aos_to_soa<D><<<(unsigned)((count + 255) / 256), 256, 0, st>>>(aos, first, count, s);
}
template void launch_aos_to_soa<2>(const float *, long long, long long, const SoA<2> &, cudaStream_t);
template void launch_aos_to_soa<3>(const float *, long long, long long, const SoA<3> &, cudaStream_t);

template <int D>
__global__ void k_soa_to_aos(SoA<D> s, long long n, long long id0, long long count, float *__restrict__ aos) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  long long id = s.id[i];
  if (id < id0 || id >= id0 + count) return;
  constexpr int W = 2 * D + 2 * D * D + 2;
  PState<D> p;
  load_full(s, i, p);
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
void launch_soa_to_aos(const SoA<D> &s, long long n, long long id0, long long count, float *aos, cudaStream_t st) {
  if (n <= 0) return;
  k_soa_to_aos<D><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(s, n, id0, count, aos);
}
template void launch_soa_to_aos<2>(const SoA<2> &, long long, long long, long long, float *, cudaStream_t);
template void launch_soa_to_aos<3>(const SoA<3> &, long long, long long, long long, float *, cudaStream_t);

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
