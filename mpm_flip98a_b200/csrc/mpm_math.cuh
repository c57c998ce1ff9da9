// mpm_math.cuh -- per-particle arithmetic of the MLS-MPM substep, in registers.
//
// Semantics follow the reference statement by statement (all paths relative to /root/reference/):
//   weights / base cell        cpp_validation/mls-mpm88-explained.cpp:55-64, 136-142
//   hardening, stress, affine  cpp_validation/mls-mpm88-explained.cpp:67-89
//   G2P tail (advect, F, SVD clamp, Jp)   cpp_validation/mls-mpm88-explained.cpp:159-178
//   polar_decomp / svd (2x2)   cpp_validation/taichi.h:8375-8420
//   determinant                cpp_validation/taichi.h:7850-7859
//   column-major products      cpp_validation/taichi.h:7591-7597, 7639-7645
// Every expression keeps the reference's association order; the translation unit is compiled with
// -fmad=false so no multiply-add is contracted and the result of every function here is bitwise
// what the reference's x86-64 build computes (expf/cbrtf excepted: within 2 ulp of libm).
// The 3x3 SVD (fixed-sweep one-sided Jacobi) has no counterpart in the reference.
//
// Everything is MPM_HD (host + device) so the same code is exercised on the CPU by
// tests/host_check.cu -- the product only ever calls it from kernels.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define MPM_HD __host__ __device__ __forceinline__
#else
#define MPM_HD inline
#endif

namespace mpm {

enum { KIND_FLUID = 0, KIND_JELLY = 1, KIND_SNOW = 2 };

struct Material {
  int kind;
  float mu_0, lambda_0;  // Lame parameters, computed on the host in fp32 (:25-26)
  float hardening;
  float sig_lo, sig_hi;
};

// Kernel-side constants of one handle (passed by value to every kernel).
struct Params {
  int n_grid;        // global cells per axis
  int n1;            // nodes per axis = n_grid + 1
  float dx, inv_dx;  // :12-13 (inv_dx = 1.0f / dx, NOT float(n_grid))
  float mass_p, vol_p;
  float gravity[3];
  float boundary;
  float jp_min, jp_max;
  float alpha;
  int n_materials;
  Material mat[4];
  int slab_lo, slab_hi;  // owned base-cell columns [lo, hi)
  int ncol;              // local node columns = slab_hi - slab_lo + 2
  int multi;             // 1 when the slab is a strict part of the grid (dead slots, migration)
};

template <int D>
struct Vec {
  float d[D];
};
// column-major like taichi.h:7575: d[col][row], operator()(i,j) == d[j][i]
template <int D>
struct Mat {
  float d[D][D];
};

template <int D>
MPM_HD Mat<D> mat_zero() {
  Mat<D> r;
#pragma unroll
  for (int c = 0; c < D; c++)
#pragma unroll
    for (int k = 0; k < D; k++) r.d[c][k] = 0.0f;
  return r;
}
template <int D>
MPM_HD Mat<D> mat_diag(float v) {  // taichi.h:7504
  Mat<D> r = mat_zero<D>();
#pragma unroll
  for (int c = 0; c < D; c++) r.d[c][c] = v;
  return r;
}
template <int D>
MPM_HD Mat<D> mat_add(const Mat<D> &a, const Mat<D> &b) {
  Mat<D> r;
#pragma unroll
  for (int c = 0; c < D; c++)
#pragma unroll
    for (int k = 0; k < D; k++) r.d[c][k] = a.d[c][k] + b.d[c][k];
  return r;
}
template <int D>
MPM_HD Mat<D> mat_sub(const Mat<D> &a, const Mat<D> &b) {
  Mat<D> r;
#pragma unroll
  for (int c = 0; c < D; c++)
#pragma unroll
    for (int k = 0; k < D; k++) r.d[c][k] = a.d[c][k] - b.d[c][k];
  return r;
}
template <int D>
MPM_HD Mat<D> mat_scale(float s, const Mat<D> &a) {  // taichi.h:7806
  Mat<D> r;
#pragma unroll
  for (int c = 0; c < D; c++)
#pragma unroll
    for (int k = 0; k < D; k++) r.d[c][k] = s * a.d[c][k];
  return r;
}
template <int D>
MPM_HD Mat<D> mat_transposed(const Mat<D> &a) {  // taichi.h:7696
  Mat<D> r;
#pragma unroll
  for (int i = 0; i < D; i++)
#pragma unroll
    for (int j = 0; j < D; j++) r.d[i][j] = a.d[j][i];
  return r;
}
// taichi.h:7591-7597: ret = d[0]*o[0]; ret += d[1]*o[1]; ...
template <int D>
MPM_HD void mat_mulvec(const Mat<D> &a, const float *v, float *out) {
#pragma unroll
  for (int k = 0; k < D; k++) {
    float r = a.d[0][k] * v[0];
#pragma unroll
    for (int c = 1; c < D; c++) r = r + a.d[c][k] * v[c];
    out[k] = r;
  }
}
template <int D>
MPM_HD Mat<D> mat_mul(const Mat<D> &a, const Mat<D> &b) {  // taichi.h:7639: column i = a * b[i]
  Mat<D> r;
#pragma unroll
  for (int c = 0; c < D; c++) mat_mulvec<D>(a, b.d[c], r.d[c]);
  return r;
}
MPM_HD float mat_det(const Mat<2> &m) {  // taichi.h:7850-7852
  return m.d[0][0] * m.d[1][1] - m.d[0][1] * m.d[1][0];
}
MPM_HD float mat_det(const Mat<3> &m) {  // taichi.h:7855-7859
  return m.d[0][0] * (m.d[1][1] * m.d[2][2] - m.d[2][1] * m.d[1][2]) -
         m.d[1][0] * (m.d[0][1] * m.d[2][2] - m.d[2][1] * m.d[0][2]) +
         m.d[2][0] * (m.d[0][1] * m.d[1][2] - m.d[1][1] * m.d[0][2]);
}
MPM_HD float clampf(float a, float lo, float hi) {  // taichi.h:6449-6455
  if (a < lo) return lo;
  if (a > hi) return hi;
  return a;
}

// taichi.h:8375-8385 (no guard for x == y == 0, like the reference)
MPM_HD void polar2(const Mat<2> &m, Mat<2> &R, Mat<2> &S) {
  float x = m.d[0][0] + m.d[1][1];
  float y = m.d[0][1] - m.d[1][0];  // m(1,0) - m(0,1)
  float scale = 1.0f / sqrtf(x * x + y * y);
  float c = x * scale, s = y * scale;
  R.d[0][0] = c;   // R(0,0)
  R.d[1][0] = -s;  // R(0,1)
  R.d[0][1] = s;   // R(1,0)
  R.d[1][1] = c;   // R(1,1)
  S = mat_mul<2>(mat_transposed<2>(R), m);
}

// taichi.h:8389-8420, both data-dependent branches kept
MPM_HD void svd2(const Mat<2> &m, Mat<2> &U, Mat<2> &sig, Mat<2> &V) {
  Mat<2> S;
  polar2(m, U, S);
  float c, s;
  // S(0,1) == S.d[1][0]
  if (fabsf(S.d[1][0]) < 1e-6f) {
    sig = S;
    c = 1.0f;
    s = 0.0f;
  } else {
    float tao = 0.5f * (S.d[0][0] - S.d[1][1]);
    float w = sqrtf(tao * tao + S.d[1][0] * S.d[1][0]);
    float t = tao > 0 ? S.d[1][0] / (tao + w) : S.d[1][0] / (tao - w);
    c = 1.0f / sqrtf(t * t + 1);
    s = -t * c;
    sig.d[0][0] = (c * c) * S.d[0][0] - 2 * c * s * S.d[1][0] + (s * s) * S.d[1][1];
    sig.d[1][1] = (s * s) * S.d[0][0] + 2 * c * s * S.d[1][0] + (c * c) * S.d[1][1];
  }
  // V is built as V(i,j) then transposed (:8406-8418)
  float v00, v01, v10, v11;
  if (sig.d[0][0] < sig.d[1][1]) {
    float t = sig.d[0][0];
    sig.d[0][0] = sig.d[1][1];
    sig.d[1][1] = t;
    v00 = -s;
    v01 = -c;
    v10 = c;
    v11 = -s;
  } else {
    v00 = c;
    v01 = -s;
    v10 = s;
    v11 = c;
  }
  // V = transposed: V'(i,j) = V(j,i); storage d[col][row] = V'(row,col) = V(col,row)
  V.d[0][0] = v00;
  V.d[0][1] = v01;
  V.d[1][0] = v10;
  V.d[1][1] = v11;
  U = mat_mul<2>(U, V);
}

// One-sided (Hestenes) Jacobi SVD, 4 fixed cyclic sweeps over column pairs (0,1),(0,2),(1,2);
// singular values sorted descending, det U = det V = +1 (sigma_2 carries the sign of det A).
// Identical operation order to the CPU oracle's svd3 (oracle/mpm_oracle.cpp) -- this routine has
// no counterpart in the reference (taichi.h has 2x2 decompositions only).
MPM_HD void svd3(const Mat<3> &A_in, Mat<3> &U, float sig[3], Mat<3> &V) {
  Mat<3> A = A_in;
  V = mat_diag<3>(1.0f);
#pragma unroll 1
  for (int sweep = 0; sweep < 4; sweep++) {
#pragma unroll
    for (int pr = 0; pr < 3; pr++) {
      const int p = pr == 2 ? 1 : 0, q = pr == 0 ? 1 : 2;
      float a = A.d[p][0] * A.d[p][0] + A.d[p][1] * A.d[p][1] + A.d[p][2] * A.d[p][2];
      float b = A.d[q][0] * A.d[q][0] + A.d[q][1] * A.d[q][1] + A.d[q][2] * A.d[q][2];
      float c = A.d[p][0] * A.d[q][0] + A.d[p][1] * A.d[q][1] + A.d[p][2] * A.d[q][2];
      // converged pair: the columns are orthogonal to fp32 resolution (|cos| <= 2 ulp); later sweeps then only pay the test
      if (fabsf(c) <= 2.4e-7f * sqrtf(a * b)) continue;
      float zeta = (b - a) / (2.0f * c);
      float t = (zeta >= 0.0f ? 1.0f : -1.0f) / (fabsf(zeta) + sqrtf(1.0f + zeta * zeta));
      float cs = 1.0f / sqrtf(1.0f + t * t);
      float sn = cs * t;
#pragma unroll
      for (int k = 0; k < 3; k++) {
        float ap = A.d[p][k], aq = A.d[q][k];
        A.d[p][k] = cs * ap - sn * aq;
        A.d[q][k] = sn * ap + cs * aq;
        float vp = V.d[p][k], vq = V.d[q][k];
        V.d[p][k] = cs * vp - sn * vq;
        V.d[q][k] = sn * vp + cs * vq;
      }
    }
  }
  float s[3];
#pragma unroll
  for (int c = 0; c < 3; c++) s[c] = sqrtf(A.d[c][0] * A.d[c][0] + A.d[c][1] * A.d[c][1] + A.d[c][2] * A.d[c][2]);
#define MPM_SWAP_COLS(i, j)                 \
  {                                         \
    float t_ = s[i];                        \
    s[i] = s[j];                            \
    s[j] = t_;                              \
    for (int k = 0; k < 3; k++) {           \
      t_ = A.d[i][k];                       \
      A.d[i][k] = A.d[j][k];                \
      A.d[j][k] = t_;                       \
      t_ = V.d[i][k];                       \
      V.d[i][k] = V.d[j][k];                \
      V.d[j][k] = t_;                       \
    }                                       \
  }
  if (s[0] < s[1]) MPM_SWAP_COLS(0, 1);
  if (s[0] < s[2]) MPM_SWAP_COLS(0, 2);
  if (s[1] < s[2]) MPM_SWAP_COLS(1, 2);
#undef MPM_SWAP_COLS
#pragma unroll
  for (int c = 0; c < 3; c++) {
    float inv = s[c] > 0.0f ? 1.0f / s[c] : 0.0f;
#pragma unroll
    for (int k = 0; k < 3; k++) U.d[c][k] = A.d[c][k] * inv;
  }
  if (mat_det(V) < 0.0f) {
#pragma unroll
    for (int k = 0; k < 3; k++) {
      V.d[2][k] = -V.d[2][k];
      U.d[2][k] = -U.d[2][k];
    }
  }
  if (mat_det(U) < 0.0f) {
#pragma unroll
    for (int k = 0; k < 3; k++) U.d[2][k] = -U.d[2][k];
    s[2] = -s[2];
  }
  sig[0] = s[0];
  sig[1] = s[1];
  sig[2] = s[2];
}

// rotation factor of F used by the stress (:75-76); 3D: U V^T of the Jacobi SVD
MPM_HD Mat<2> rotation_of(const Mat<2> &F) {
  Mat<2> r, s;
  polar2(F, r, s);
  return r;
}
MPM_HD Mat<3> rotation_of_svd(const Mat<3> &F) {
  Mat<3> U, V;
  float sg[3];
  svd3(F, U, sg, V);
  return mat_mul<3>(U, mat_transposed<3>(V));
}

// Device: reciprocal square root / division by the special-function unit (<= 2 ulp) in the 3D decompositions: the Jacobi rotations of
// plastic_project3 (a rotation angle that is off by an ulp leaves an off-diagonal of ~1e-7 |apq| behind, far below
// what the next sweep tests for).  The Newton polar iteration stays exact: it feeds P2G, which the deterministic mode
// promises bitwise.  Host (tests/host_check.cpp): the oracle's exact statements.
#if defined(__CUDA_ARCH__)
#define MPM_RSQRT(x) rsqrtf(x)
#define MPM_FDIV(a, b) __fdividef((a), (b))
#define MPM_SQRT_POS(x) ((x) * rsqrtf(x))
#else
#define MPM_RSQRT(x) (1.0f / sqrtf(x))
#define MPM_FDIV(a, b) ((a) / (b))
#define MPM_SQRT_POS(x) sqrtf(x)
#endif
// Rotation factor for the 3D stress (the 3D lift has no counterpart in the reference; the CPU oracle defines it with
// the very same statements, oracle/mpm_oracle.cpp polar_newton3 / rotation3): Newton's iteration for the polar
// decomposition, X <- (X + X^-T) / 2 from X0 = F, converges quadratically to the rotation R of F = R S.
// |X_{k+1} - X_k| <= 4e-4 means X_{k+1} is within ~8e-8 of R (fp32 rounding): 2 iterations for snow (F within
// 2.5 % of a rotation after the plastic clamp), 3-4 for a jelly under load, ~65 instructions each (cofactors as
// multiply + FMA) -- against ~1500 for
// the 4-sweep Jacobi SVD whose U V^T it replaced in round 2 (the two agree to ~1e-6, tests/test_host_math.py).
// Returns false for a near-singular or inverted F (det <= 1e-6 |F|^3): the callers then take the SVD.
MPM_HD float cof(float p, float q, float r, float t) { return fmaf(p, q, -(r * t)); }  // p q - r t
MPM_HD bool polar_newton3(const Mat<3> &F, Mat<3> &X) {
  X = F;
  float scale = 0.0f;
#pragma unroll
  for (int c = 0; c < 3; c++)
#pragma unroll
    for (int k = 0; k < 3; k++) scale = fmaxf(scale, fabsf(F.d[c][k]));
#pragma unroll 1
  for (int it = 0; it < 12; it++) {
    const float *a = X.d[0], *b = X.d[1], *c = X.d[2];
    Mat<3> K;  // cofactors: columns b x c, c x a, a x b  (X^-T = K / det)
    K.d[0][0] = cof(b[1], c[2], b[2], c[1]); K.d[0][1] = cof(b[2], c[0], b[0], c[2]); K.d[0][2] = cof(b[0], c[1], b[1], c[0]);
    K.d[1][0] = cof(c[1], a[2], c[2], a[1]); K.d[1][1] = cof(c[2], a[0], c[0], a[2]); K.d[1][2] = cof(c[0], a[1], c[1], a[0]);
    K.d[2][0] = cof(a[1], b[2], a[2], b[1]); K.d[2][1] = cof(a[2], b[0], a[0], b[2]); K.d[2][2] = cof(a[0], b[1], a[1], b[0]);
    const float det = fmaf(a[2], K.d[0][2], fmaf(a[1], K.d[0][1], a[0] * K.d[0][0]));
    if (!(det > 1e-6f * scale * scale * scale)) return false;
    const float h = 0.5f / det;  // exact: MPM_FLAG_DETERMINISTIC promises a P2G grid bitwise the oracle's
    float delta = 0.0f;
#pragma unroll
    for (int cc = 0; cc < 3; cc++)
#pragma unroll
      for (int k = 0; k < 3; k++) {
        const float y = fmaf(h, K.d[cc][k], 0.5f * X.d[cc][k]);
        delta = fmaxf(delta, fabsf(y - X.d[cc][k]));
        X.d[cc][k] = y;
      }
    if (delta <= 4e-4f) break;
  }
  return true;
}
MPM_HD Mat<3> rotation_of(const Mat<3> &F) {
  Mat<3> X;
  if (polar_newton3(F, X)) return X;
  return rotation_of_svd(F);
}

MPM_HD float dot3f(const float *a, const float *b) { return fmaf(a[2], b[2], fmaf(a[1], b[1], a[0] * b[0])); }

// One Jacobi rotation of the symmetric 3x3 (diagonal app, aqq; off-diagonal apq; arp, arq = the entries that couple
// the third index) with the eigenvector columns p, q of V.  Returns whether it rotated.
MPM_HD bool jacobi_rotate3(float &app, float &aqq, float &apq, float &arp, float &arq, float *vp, float *vq) {
  if (!(fabsf(apq) > 6e-8f * (fabsf(app) + fabsf(aqq)))) return false;
  const float h = 0.5f * (aqq - app);
  const float t = MPM_FDIV(apq, h + copysignf(MPM_SQRT_POS(fmaf(h, h, apq * apq)), h));
  const float c = MPM_RSQRT(fmaf(t, t, 1.0f)), s = t * c;
  app = fmaf(-t, apq, app);
  aqq = fmaf(t, apq, aqq);
  apq = 0.0f;
  const float x = arp, y = arq;
  arp = fmaf(c, x, -(s * y));
  arq = fmaf(s, x, c * y);
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const float a = vp[k], b = vq[k];
    vp[k] = fmaf(c, a, -(s * b));
    vq[k] = fmaf(s, a, c * b);
  }
  return true;
}

// Plastic projection of snow in 3D (:165-178 lifted): clamps the singular values of F to [lo, hi] and returns
// det(F) / det(F') for the Jp update (:175).  Same structure as the reference's 2x2 svd (taichi.h:8389-8420: polar
// decomposition first, then Jacobi rotations that diagonalise the symmetric factor):
//   F = R S (polar_newton3),  S = V diag(l) V^T (cyclic Jacobi),  F' = R V diag(clamp(l)) V^T.
// When Gershgorin's discs already place every eigenvalue of S inside [lo, hi] nothing is clamped: F stays, ratio 1.
// Near-singular or inverted F takes the one-sided Jacobi SVD.  Statement for statement oracle/mpm_oracle.cpp
// plastic_project3 (the definition; the reference has no 3D code).  `R_out` (optional) receives the rotation factor
// of F' -- the R the next P2G's stress needs (:75-76), free here.
MPM_HD float plastic_project3(float lo, float hi, Mat<3> &F, Mat<3> *R_out = nullptr) {
  Mat<3> R;
  if (!polar_newton3(F, R)) {
    Mat<3> U, V;
    float sg[3];
    svd3(F, U, sg, V);
    Mat<3> sig = mat_zero<3>();
#pragma unroll
    for (int i = 0; i < 3; i++) sig.d[i][i] = clampf(sg[i], lo, hi);
    const float oldJ = mat_det(F);
    F = mat_mul<3>(mat_mul<3>(U, sig), mat_transposed<3>(V));
    if (R_out) *R_out = mat_mul<3>(U, mat_transposed<3>(V));
    return oldJ / mat_det(F);
  }
  if (R_out) *R_out = R;
  float a00 = dot3f(R.d[0], F.d[0]), a11 = dot3f(R.d[1], F.d[1]), a22 = dot3f(R.d[2], F.d[2]);
  float a01 = 0.5f * (dot3f(R.d[0], F.d[1]) + dot3f(R.d[1], F.d[0]));
  float a02 = 0.5f * (dot3f(R.d[0], F.d[2]) + dot3f(R.d[2], F.d[0]));
  float a12 = 0.5f * (dot3f(R.d[1], F.d[2]) + dot3f(R.d[2], F.d[1]));
  {
    const float r0 = fabsf(a01) + fabsf(a02), r1 = fabsf(a01) + fabsf(a12), r2 = fabsf(a02) + fabsf(a12);
    if (a00 - r0 >= lo && a00 + r0 <= hi && a11 - r1 >= lo && a11 + r1 <= hi && a22 - r2 >= lo && a22 + r2 <= hi) return 1.0f;
  }
  Mat<3> V = mat_diag<3>(1.0f);
#pragma unroll 1
  for (int sweep = 0; sweep < 6; sweep++) {
    bool rotated = jacobi_rotate3(a00, a11, a01, a02, a12, V.d[0], V.d[1]);
    rotated |= jacobi_rotate3(a00, a22, a02, a01, a12, V.d[0], V.d[2]);
    rotated |= jacobi_rotate3(a11, a22, a12, a01, a02, V.d[1], V.d[2]);
    if (!rotated) break;
  }
  const float l0 = clampf(a00, lo, hi), l1 = clampf(a11, lo, hi), l2 = clampf(a22, lo, hi);
  float w0[3], w1[3], w2[3];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    w0[k] = l0 * V.d[0][k];
    w1[k] = l1 * V.d[1][k];
    w2[k] = l2 * V.d[2][k];
  }
  Mat<3> S;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = i; j < 3; j++) S.d[j][i] = S.d[i][j] = fmaf(w2[i], V.d[2][j], fmaf(w1[i], V.d[1][j], w0[i] * V.d[0][j]));
#pragma unroll
  for (int j = 0; j < 3; j++)
#pragma unroll
    for (int k = 0; k < 3; k++) F.d[j][k] = fmaf(R.d[2][k], S.d[j][2], fmaf(R.d[1][k], S.d[j][1], R.d[0][k] * S.d[j][0]));
  return (a00 * a11 * a22) / (l0 * l1 * l2);
}


// ---------------------------------------------------------------------------------------------
// Base cell + quadratic B-spline weights, :55-64.  The cast is C++ truncation (taichi.h:7185),
// and x*inv_dx - 0.5f must stay a separate multiply and subtract so the cell is bit-exact.
// ---------------------------------------------------------------------------------------------
template <int D>
struct Stencil {
  int base[D];
  float fx[D];
  float w[3][D];
};

// never contracted into an FMA, whatever -fmad says: the integer cell must be bit-exact
#if defined(__CUDA_ARCH__)
#define MPM_MUL_RN(a, b) __fmul_rn((a), (b))
#define MPM_SUB_RN(a, b) __fsub_rn((a), (b))
#else
#define MPM_MUL_RN(a, b) ((a) * (b))
#define MPM_SUB_RN(a, b) ((a) - (b))
#endif
MPM_HD int base_coord(float x, float inv_dx) { return (int)MPM_SUB_RN(MPM_MUL_RN(x, inv_dx), 0.5f); }

template <int D>
MPM_HD Stencil<D> make_stencil(const float *x, float inv_dx) {
  Stencil<D> s;
#pragma unroll
  for (int k = 0; k < D; k++) {
    s.base[k] = base_coord(x[k], inv_dx);       // :55
    s.fx[k] = MPM_SUB_RN(MPM_MUL_RN(x[k], inv_dx), (float)s.base[k]);  // :57
    s.w[0][k] = 0.5f * ((1.5f - s.fx[k]) * (1.5f - s.fx[k]));       // :61
    s.w[1][k] = 0.75f - ((s.fx[k] - 1.0f) * (s.fx[k] - 1.0f));      // :62
    s.w[2][k] = 0.5f * ((s.fx[k] - 0.5f) * (s.fx[k] - 0.5f));       // :63
  }
  return s;
}

MPM_HD int material_index(const Params &P, int c) {
  return (c >= 0 && c < P.n_materials) ? c : P.n_materials - 1;
}

// :67-89 -- the matrix `affine` such that the node contribution is
//   w * ( (mass_p * v, mass_p) + (affine * dpos, 0) ),  dpos = (node_offset - fx) * dx
// `rot`: the rotation factor of F (:75-76) when the caller already has it (the fused 3D kernel takes it from the snow
// projection of the same substep), else nullptr: computed here.  Not read for fluids.
template <int D>
MPM_HD Mat<D> p2g_affine(const Params &P, const Material &mat, float dt, const Mat<D> &F, const Mat<D> &C, float Jp,
                         const Mat<D> *rot = nullptr) {
  float e;
  if (mat.kind == KIND_SNOW) e = expf(mat.hardening * (1.0f - Jp));  // :67
  else if (mat.kind == KIND_JELLY) e = mat.hardening;
  else e = 1.0f;
  float mu = mat.mu_0 * e;          // :68
  float lambda = mat.lambda_0 * e;  // :69
  float J = mat_det(F);             // :72
  float Dinv = 4 * P.inv_dx * P.inv_dx;  // :79
  Mat<D> PF;
  if (mat.kind == KIND_FLUID) {
    PF = mat_diag<D>(lambda * (J - 1) * J);
  } else {
    Mat<D> r = rot ? *rot : rotation_of(F);  // :75-76
    PF = mat_add<D>(mat_mul<D>(mat_scale<D>(2 * mu, mat_sub<D>(F, r)), mat_transposed<D>(F)),
                    mat_diag<D>(lambda * (J - 1) * J));  // :81
  }
  Mat<D> stress = mat_scale<D>(-(dt * P.vol_p), mat_scale<D>(Dinv, PF));  // :84
  return mat_add<D>(stress, mat_scale<D>(P.mass_p, C));                    // :89
}

// :92-100 -- contribution of one particle to the node at stencil offset (a,b[,c]):
//   val[0..D-1] = w * (mass_p*v + affine*dpos), val[D] = w * mass_p,  dpos = (offset - fx) * dx
template <int D>
MPM_HD void p2g_node_value(const Params &P, const Stencil<D> &st, const Mat<D> &affine, const float *mv, int a, int b,
                           int c, float *val) {
  float dpos[D], ad[D];
  dpos[0] = ((float)a - st.fx[0]) * P.dx;  // :94
  dpos[1] = ((float)b - st.fx[1]) * P.dx;
  if (D == 3) dpos[D - 1] = ((float)c - st.fx[D - 1]) * P.dx;
  mat_mulvec<D>(affine, dpos, ad);
  float w = st.w[a][0] * st.w[b][1];
  if (D == 3) w = w * st.w[c][D - 1];
#pragma unroll
  for (int r = 0; r < D; r++) val[r] = w * (mv[r] + ad[r]);  // :97-100
  val[D] = w * (P.mass_p + 0.0f);
}

// :147-156 -- one node of the G2P gather.  gv = node velocity, vo = its pre-gravity value
// (only read when flip); accumulates v (:153), C (:154) and dv in the reference's order.
template <int D>
MPM_HD void g2p_accumulate(const Params &P, const Stencil<D> &st, int a, int b, int c, const float *gv, const float *vo,
                           bool flip, float *v, Mat<D> &C, float *dv) {
  float dpos[D];
  dpos[0] = (float)a - st.fx[0];  // :149
  dpos[1] = (float)b - st.fx[1];
  if (D == 3) dpos[D - 1] = (float)c - st.fx[D - 1];
  float w = st.w[a][0] * st.w[b][1];
  if (D == 3) w = w * st.w[c][D - 1];
  const float s4 = 4 * P.inv_dx;
  float wg[D];
#pragma unroll
  for (int r = 0; r < D; r++) {
    wg[r] = w * gv[r];
    v[r] = v[r] + wg[r];  // :153
  }
#pragma unroll
  for (int cc = 0; cc < D; cc++)
#pragma unroll
    for (int r = 0; r < D; r++) C.d[cc][r] = C.d[cc][r] + s4 * (wg[r] * dpos[cc]);  // :154
  if (flip) {
#pragma unroll
    for (int r = 0; r < D; r++) dv[r] = dv[r] + w * (gv[r] - vo[r]);
  }
}

// Fast form of the same gather: the constant 4*inv_dx is applied once after the loop (the caller
// multiplies C by it) and every multiply-add is an explicit FMA.  Algebraically identical to :153-154;
// rounding differs from the reference's association at the 1e-7 level (fewer roundings, not more).
// MPM_FLAG_STRICT keeps g2p_accumulate.
#if defined(__CUDACC__)
template <int D>
__device__ __forceinline__ void g2p_accumulate_fast(const Stencil<D> &st, int a, int b, int c, const float *gv,
                                                    const float *vo, bool flip, float *v, Mat<D> &Cu, float *dv) {
  float dpos[D];
  dpos[0] = (float)a - st.fx[0];
  dpos[1] = (float)b - st.fx[1];
  if (D == 3) dpos[D - 1] = (float)c - st.fx[D - 1];
  float w = st.w[a][0] * st.w[b][1];
  if (D == 3) w = w * st.w[c][D - 1];
#pragma unroll
  for (int r = 0; r < D; r++) {
    const float wg = w * gv[r];
    v[r] = v[r] + wg;
#pragma unroll
    for (int cc = 0; cc < D; cc++) Cu.d[cc][r] = __fmaf_rn(wg, dpos[cc], Cu.d[cc][r]);
    if (flip) dv[r] = __fmaf_rn(w, gv[r] - vo[r], dv[r]);
  }
}
#endif

// :105-131 -- one grid node: g = (m*v, m) in, (v, 1|0) out; vo = normalised pre-gravity velocity
// (the FLIP reference velocity).  Returns false when the node is empty (left untouched).
template <int D>
MPM_HD bool grid_node_update(const Params &P, float dt, int i, int j, int k, float *g, float *vo) {
  vo[0] = vo[1] = vo[2] = 0.0f;
  float m = g[D];
  if (!(m > 0)) return false;  // :109
#pragma unroll
  for (int c = 0; c <= D; c++) g[c] = g[c] / m;  // :111 (true division of every component)
#pragma unroll
  for (int c = 0; c < D; c++) vo[c] = g[c];
#pragma unroll
  for (int c = 0; c < D; c++) g[c] = g[c] + dt * P.gravity[c];  // :113
  g[D] = g[D] + dt * 0.0f;
  float boundary = P.boundary;    // :116
  float x = (float)i / P.n_grid;  // :118
  float y = (float)j / P.n_grid;
  bool sticky = x < boundary || x > 1 - boundary || y > 1 - boundary;  // :122
  if (D == 3) {
    float z = (float)k / P.n_grid;
    sticky = sticky || z < boundary || z > 1 - boundary;
  }
  if (sticky) {
#pragma unroll
    for (int c = 0; c <= D; c++) g[c] = 0.0f;
  }
  if (y < boundary) g[1] = fmaxf(0.0f, g[1]);  // :126-128
  return true;
}

// :165-173 -- SVD, clamp the singular values, rebuild F.  2D keeps the reference's quirk that
// `sig` carries S's tiny off-diagonals through the |S01| < 1e-6 branch (taichi.h:8393-8396).
MPM_HD void plastic_project(const Material &mat, Mat<2> &F) {
  Mat<2> U = mat_zero<2>(), sig = mat_zero<2>(), V = mat_zero<2>();
  svd2(F, U, sig, V);
  sig.d[0][0] = clampf(sig.d[0][0], mat.sig_lo, mat.sig_hi);
  sig.d[1][1] = clampf(sig.d[1][1], mat.sig_lo, mat.sig_hi);
  F = mat_mul<2>(mat_mul<2>(U, sig), mat_transposed<2>(V));
}
// fluid keeps only the volume change: F <- J^(1/d) * I
MPM_HD void fluid_project(Mat<2> &F) { F = mat_diag<2>(sqrtf(mat_det(F))); }
MPM_HD void fluid_project(Mat<3> &F) { F = mat_diag<3>(cbrtf(mat_det(F))); }

// :159-178 after the gather: advect, FLIP blend (alpha != 0 only), F update, plasticity.
// v_apic / C are the gathered values (:153-154); v_in the particle's previous velocity and dv the
// gathered grid velocity change (both only read when alpha != 0).
template <int D>
MPM_HD void g2p_finish(const Params &P, const Material &mat, float dt, float *x, float *v, const Mat<D> &C, Mat<D> &F,
                       float &Jp, const float *v_in, const float *dv) {
#pragma unroll
  for (int k = 0; k < D; k++) x[k] = x[k] + dt * v[k];  // :159
  if (P.alpha != 0.0f) {
    float a = P.alpha;
#pragma unroll
    for (int k = 0; k < D; k++) v[k] = (1.0f - a) * v[k] + a * (v_in[k] + dv[k]);
  }
  Mat<D> Fn = mat_mul<D>(mat_add<D>(mat_diag<D>(1.0f), mat_scale<D>(dt, C)), F);  // :162
  if (mat.kind == KIND_SNOW) {
    if constexpr (D == 2) {
      float oldJ = mat_det(Fn);  // :172 (F not yet rebuilt)
      plastic_project(mat, Fn);
      Jp = clampf(Jp * oldJ / mat_det(Fn), P.jp_min, P.jp_max);  // :175
    } else {
      const float ratio = plastic_project3(mat.sig_lo, mat.sig_hi, Fn);  // det(F) / det(F'), :172-175
      Jp = clampf(Jp * ratio, P.jp_min, P.jp_max);
    }
    F = Fn;
  } else if (mat.kind == KIND_JELLY) {
    F = Fn;
  } else {
    fluid_project(Fn);
    F = Fn;
  }
}

}  // namespace mpm
