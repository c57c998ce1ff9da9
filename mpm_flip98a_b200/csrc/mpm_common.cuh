// mpm_common.cuh -- device-side data layout shared by the kernels and the engine.
//
// Particle state is SoA in HBM (the reference keeps a 56-byte AoS `Particle`,
// cpp_validation/mls-mpm88-explained.cpp:28-42; it only crosses the C-ABI in that form):
//   2D: x float2 | v float2 | C float4 | F float4 | Jp float | mat int | id int      (60 B)
//   3D: xj float4 (x,y,z,Jp) | vm float4 (vx,vy,vz,mat) | C 9 float planes | F 9 float planes | id int  (108 B)
// C and F are column-major (d[col][row]) like taichi.h:7575, so a 2D float4 is (m00,m10,m01,m11)
// in (row,col) notation == the reference's in-memory order.
// Grid nodes are float4: 2D (m*vx, m*vy, m, 0) -> (vx, vy, 1|0, 0) after the grid update;
// 3D (m*vx, m*vy, m*vz, m) -> (vx, vy, vz, 1|0).  Node index = ((i - slab_lo)*n1 + j)[*n1 + k],
// x-major like the reference's grid[i][j] (:47), so an x-slab (and its ghost columns) is contiguous.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "mpm_math.cuh"

namespace mpm {

enum { STATUS_DOMAIN = 1, STATUS_CFL = 2, STATUS_MIGRATION_OVERFLOW = 4 };

// Material-id value of a storage slot whose particle emigrated to a neighbouring slab; every kernel
// skips it and the next re-sort drops it.  (c == INT_MIN is therefore reserved at the C-ABI.)
constexpr int DEAD = (int)0x80000000;

// Emigrant staging of one handle (x-slab runs only): packed records = the AoS record + id + pad.
struct MigPtrs {
  float *send_lo, *send_hi;  // device, `cap` records each
  int *count;                // device: [0] = lo, [1] = hi
  int cap;
  int enabled;
  int interior;  // 1: this launch covers bins far from the slab cuts and must not emigrate (checked, flagged)
};
template <int D>
struct MigRec {
  static constexpr int AOS = 2 * D + 2 * D * D + 2;  // 14 | 26 words
  static constexpr int WORDS = AOS + 2;              // + id + pad: 16 | 28 words (64 | 112 B)
};

template <int D>
struct SoA;

template <>
struct SoA<2> {
  float2 *x, *v;
  float4 *C, *F;
  float *Jp;
  int *mat;
  int *id;
};

template <>
struct SoA<3> {
  float4 *xj;  // x, y, z, Jp
  float4 *vm;  // vx, vy, vz, material id (bit pattern)
  float *C[9];
  float *F[9];
  int *id;
};

template <int D>
struct PState {
  float x[D], v[D];
  Mat<D> C, F;
  float Jp;
  int mat;
};

// ---- loads / stores ----------------------------------------------------------------------------
__device__ __forceinline__ void load_full(const SoA<2> &s, long long i, PState<2> &p) {
  float2 x = s.x[i], v = s.v[i];
  float4 C = s.C[i], F = s.F[i];
  p.x[0] = x.x; p.x[1] = x.y; p.v[0] = v.x; p.v[1] = v.y;
  p.C.d[0][0] = C.x; p.C.d[0][1] = C.y; p.C.d[1][0] = C.z; p.C.d[1][1] = C.w;
  p.F.d[0][0] = F.x; p.F.d[0][1] = F.y; p.F.d[1][0] = F.z; p.F.d[1][1] = F.w;
  p.Jp = s.Jp[i];
  p.mat = s.mat[i];
}
__device__ __forceinline__ void load_full(const SoA<3> &s, long long i, PState<3> &p) {
  float4 x = s.xj[i], v = s.vm[i];
  p.x[0] = x.x; p.x[1] = x.y; p.x[2] = x.z; p.Jp = x.w;
  p.v[0] = v.x; p.v[1] = v.y; p.v[2] = v.z; p.mat = __float_as_int(v.w);
#pragma unroll
  for (int c = 0; c < 3; c++)
#pragma unroll
    for (int r = 0; r < 3; r++) {
      p.C.d[c][r] = s.C[c * 3 + r][i];
      p.F.d[c][r] = s.F[c * 3 + r][i];
    }
}
// what G2P needs: x, F, Jp, mat (+ v when blending)
__device__ __forceinline__ void load_g2p(const SoA<2> &s, long long i, PState<2> &p, bool need_v) {
  float2 x = s.x[i];
  float4 F = s.F[i];
  p.x[0] = x.x; p.x[1] = x.y;
  p.F.d[0][0] = F.x; p.F.d[0][1] = F.y; p.F.d[1][0] = F.z; p.F.d[1][1] = F.w;
  p.Jp = s.Jp[i];
  p.mat = s.mat[i];
  if (need_v) {
    float2 v = s.v[i];
    p.v[0] = v.x; p.v[1] = v.y;
  }
}
__device__ __forceinline__ void load_g2p(const SoA<3> &s, long long i, PState<3> &p, bool need_v) {
  float4 x = s.xj[i], v = s.vm[i];  // vm also carries the material id
  p.x[0] = x.x; p.x[1] = x.y; p.x[2] = x.z; p.Jp = x.w;
  p.v[0] = v.x; p.v[1] = v.y; p.v[2] = v.z; p.mat = __float_as_int(v.w);
#pragma unroll
  for (int c = 0; c < 3; c++)
#pragma unroll
    for (int r = 0; r < 3; r++) p.F.d[c][r] = s.F[c * 3 + r][i];
}
// everything G2P produces (x, v, C, F, Jp); mat/id are written only when moving between buffers
__device__ __forceinline__ void store_state(const SoA<2> &s, long long i, const PState<2> &p) {
  s.x[i] = make_float2(p.x[0], p.x[1]);
  s.v[i] = make_float2(p.v[0], p.v[1]);
  s.C[i] = make_float4(p.C.d[0][0], p.C.d[0][1], p.C.d[1][0], p.C.d[1][1]);
  s.F[i] = make_float4(p.F.d[0][0], p.F.d[0][1], p.F.d[1][0], p.F.d[1][1]);
  s.Jp[i] = p.Jp;
}
__device__ __forceinline__ void store_state(const SoA<3> &s, long long i, const PState<3> &p) {
  s.xj[i] = make_float4(p.x[0], p.x[1], p.x[2], p.Jp);
  s.vm[i] = make_float4(p.v[0], p.v[1], p.v[2], __int_as_float(p.mat));
#pragma unroll
  for (int c = 0; c < 3; c++)
#pragma unroll
    for (int r = 0; r < 3; r++) {
      s.C[c * 3 + r][i] = p.C.d[c][r];
      s.F[c * 3 + r][i] = p.F.d[c][r];
    }
}
__device__ __forceinline__ void store_tags(const SoA<2> &s, long long i, int mat, int id) {
  s.mat[i] = mat;
  s.id[i] = id;
}
__device__ __forceinline__ void store_tags(const SoA<3> &s, long long i, int /*mat: lives in vm.w*/, int id) {
  s.id[i] = id;
}
// store_state() that follows writes vm.w = DEAD in 3D (p.mat); 2D keeps mat in its own array
__device__ __forceinline__ void mark_dead(const SoA<2> &s, long long i) { s.mat[i] = DEAD; }
__device__ __forceinline__ void mark_dead(const SoA<3> &, long long) {}
__device__ __forceinline__ int load_mat(const SoA<2> &s, long long i) { return s.mat[i]; }
__device__ __forceinline__ int load_mat(const SoA<3> &s, long long i) { return __float_as_int(s.vm[i].w); }
__device__ __forceinline__ void load_pos(const SoA<2> &s, long long i, float *x) {
  float2 v = s.x[i];
  x[0] = v.x; x[1] = v.y;
}
__device__ __forceinline__ void load_pos(const SoA<3> &s, long long i, float *x) {
  float4 v = s.xj[i];
  x[0] = v.x; x[1] = v.y; x[2] = v.z;
}

// Base cell clamped into the addressable range; flags STATUS_DOMAIN when it had to clamp.
// x-range is the owned slab [slab_lo, slab_hi) (whole grid on a single GPU: [0, n_grid-1)).
template <int D>
__device__ __forceinline__ int clamp_base(const Params &P, int *base) {
  int bad = 0;
  int lo = P.slab_lo, hi = min(P.slab_hi, P.n_grid - 1) - 1;
  if (base[0] < lo) { base[0] = lo; bad = STATUS_DOMAIN; }
  if (base[0] > hi) { base[0] = hi; bad = STATUS_DOMAIN; }
#pragma unroll
  for (int k = 1; k < D; k++) {
    if (base[k] < 0) { base[k] = 0; bad = STATUS_DOMAIN; }
    if (base[k] > P.n_grid - 2) { base[k] = P.n_grid - 2; bad = STATUS_DOMAIN; }
  }
  return bad;
}

template <int D>
__device__ __forceinline__ long long node_index(const Params &P, int i, int j, int k) {
  long long n = (long long)(i - P.slab_lo) * P.n1 + j;
  if (D == 3) n = n * P.n1 + k;
  return n;
}

}  // namespace mpm
