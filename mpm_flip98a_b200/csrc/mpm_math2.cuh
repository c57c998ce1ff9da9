// mpm_math2.cuh -- the 2D per-particle arithmetic of mpm_math.cuh restated on float2 pairs, so that on
// sm_100a every 2-vector operation is ONE packed instruction (FMUL2 / FADD2 / FFMA2; a scalar operand
// is broadcast by the instruction itself).  Nearly every quantity of the 2D substep is a 2-vector: x, v,
// the columns of C and F (column-major like taichi.h:7575), the per-axis B-spline weights.
//
// Exactness: mul2 / add2 / sub2 are IEEE round-to-nearest per component, so a packed expression that keeps
// the reference's association is, as written, BITWISE the scalar one.  stencil2(), affine2() and g2p_finish2()
// keep it (same statements as make_stencil / p2g_affine / g2p_finish, cpp_validation/mls-mpm88-explained.cpp
// :55-64, :67-89, :159-178) and are checked bitwise against the oracle on the host
// (tests/test_host_math.py::test_packed2d_*).  ON THE DEVICE there is one licence: ptxas (CUDA 12.9) contracts a
// mul.rn.f32x2 feeding an add.rn.f32x2 into a single FFMA2 -- -fmad=false does not reach the packed type -- so a
// packed multiply-add may round once instead of twice (<= 1 ulp, the size of the reference's own summation-order
// noise).  The integer base cell is therefore formed with scalar mul.rn / sub.rn and stays bit-exact.
// gather2() is the fast separable form of :147-156 (fused multiply-adds by design, algebraically identical, ~1e-7
// relative from the reference association).  MPM_FLAG_STRICT never reaches this header: the scalar kernels are
// bit-faithful.
#pragma once
#include "mpm_math.cuh"

namespace mpm {

#if defined(__CUDACC__)
typedef float2 f2;
MPM_HD f2 mk2(float a, float b) { return make_float2(a, b); }
#else
struct f2 {
  float x, y;
};
MPM_HD f2 mk2(float a, float b) {
  f2 r;
  r.x = a;
  r.y = b;
  return r;
}
#endif
MPM_HD f2 sp2(float a) { return mk2(a, a); }
#if defined(__CUDA_ARCH__)
MPM_HD f2 mul2(f2 a, f2 b) { return __fmul2_rn(a, b); }
MPM_HD f2 add2(f2 a, f2 b) { return __fadd2_rn(a, b); }
MPM_HD f2 fma2(f2 a, f2 b, f2 c) { return __ffma2_rn(a, b, c); }
#else
MPM_HD f2 mul2(f2 a, f2 b) { return mk2(a.x * b.x, a.y * b.y); }
MPM_HD f2 add2(f2 a, f2 b) { return mk2(a.x + b.x, a.y + b.y); }
MPM_HD f2 fma2(f2 a, f2 b, f2 c) { return mk2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
#endif
MPM_HD f2 sub2(f2 a, f2 b) { return add2(a, mk2(-b.x, -b.y)); }  // a - b == a + (-b) exactly

// 2x2 matrix as its two columns: c0 = (m00, m10) = d[0][*], c1 = (m01, m11) = d[1][*]
struct M2c {
  f2 c0, c1;
};
MPM_HD M2c to_cols(const Mat<2> &m) {
  M2c r;
  r.c0 = mk2(m.d[0][0], m.d[0][1]);
  r.c1 = mk2(m.d[1][0], m.d[1][1]);
  return r;
}
MPM_HD Mat<2> to_mat(const M2c &m) {
  Mat<2> r;
  r.d[0][0] = m.c0.x; r.d[0][1] = m.c0.y; r.d[1][0] = m.c1.x; r.d[1][1] = m.c1.y;
  return r;
}
// taichi.h:7591-7597 per column: out = a.col0*v0 + a.col1*v1 (separate multiplies and add)
MPM_HD f2 mulvec2(const M2c &a, float v0, float v1) { return add2(mul2(a.c0, sp2(v0)), mul2(a.c1, sp2(v1))); }

// :55-64 on both axes at once.  base is NOT clamped here (fx must come from the unclamped cell, like
// make_stencil); w[k] = (w[k][0], w[k][1]).
struct Sten2 {
  int bx, by;
  f2 fx;
  f2 w[3];
};
MPM_HD Sten2 stencil2(f2 x, float inv_dx) {
  Sten2 s;
  // scalar on purpose: the integer cell must be bit-exact, and ptxas (CUDA 12.9) contracts a packed
  // mul.rn.f32x2 + add.rn.f32x2 pair into one FFMA2 whatever -fmad says; mul.rn.f32 / sub.rn.f32 it never touches
  const float tx = MPM_MUL_RN(x.x, inv_dx), ty = MPM_MUL_RN(x.y, inv_dx);  // rounded before the subtraction (:55, :57)
  s.bx = (int)MPM_SUB_RN(tx, 0.5f);      // truncation, taichi.h:7185
  s.by = (int)MPM_SUB_RN(ty, 0.5f);
  s.fx = mk2(MPM_SUB_RN(tx, (float)s.bx), MPM_SUB_RN(ty, (float)s.by));
  const f2 a = sub2(sp2(1.5f), s.fx), b = sub2(s.fx, sp2(1.0f)), c = sub2(s.fx, sp2(0.5f));
  s.w[0] = mul2(sp2(0.5f), mul2(a, a));         // :61
  s.w[1] = sub2(sp2(0.75f), mul2(b, b));        // :62
  s.w[2] = mul2(sp2(0.5f), mul2(c, c));         // :63
  return s;
}
// weights only, from a stored fx (the three expressions of :61-63 again)
MPM_HD void weights2(f2 fx, f2 *w) {
  const f2 a = sub2(sp2(1.5f), fx), b = sub2(fx, sp2(1.0f)), c = sub2(fx, sp2(0.5f));
  w[0] = mul2(sp2(0.5f), mul2(a, a));
  w[1] = sub2(sp2(0.75f), mul2(b, b));
  w[2] = mul2(sp2(0.5f), mul2(c, c));
}

// :67-89, the statements of p2g_affine<2> on columns (bitwise the same result)
MPM_HD M2c affine2(const Params &P, const Material &mat, float dt, const M2c &F, const M2c &C, float Jp) {
  float e;
  if (mat.kind == KIND_SNOW) e = expf(mat.hardening * (1.0f - Jp));  // :67
  else if (mat.kind == KIND_JELLY) e = mat.hardening;
  else e = 1.0f;
  const float mu = mat.mu_0 * e;          // :68
  const float lambda = mat.lambda_0 * e;  // :69
  const float J = F.c0.x * F.c1.y - F.c0.y * F.c1.x;  // :72, taichi.h:7850
  const float Dinv = 4 * P.inv_dx * P.inv_dx;         // :79
  const float dl = lambda * (J - 1) * J;
  M2c PF;
  if (mat.kind == KIND_FLUID) {
    PF.c0 = mk2(dl, 0.0f);
    PF.c1 = mk2(0.0f, dl);
  } else {
    // polar_decomp, taichi.h:8375-8385: R = [[c, -s], [s, c]]
    const float x = F.c0.x + F.c1.y;
    const float y = F.c0.y - F.c1.x;
    const float scale = 1.0f / sqrtf(x * x + y * y);
    const float c = x * scale, s = y * scale;
    M2c M;  // (2*mu) * (F - R)
    M.c0 = mul2(sp2(2 * mu), sub2(F.c0, mk2(c, s)));
    M.c1 = mul2(sp2(2 * mu), sub2(F.c1, mk2(-s, c)));
    // ... * transposed(F): column k of the product = M * (F.d[0][k], F.d[1][k])
    PF.c0 = add2(mulvec2(M, F.c0.x, F.c1.x), mk2(dl, 0.0f));  // :81 (+ scalar promoted to dl*I, taichi.h:7504)
    PF.c1 = add2(mulvec2(M, F.c0.y, F.c1.y), mk2(0.0f, dl));
  }
  const float k = -(dt * P.vol_p);
  M2c A;  // :84, :89
  A.c0 = add2(mul2(sp2(k), mul2(sp2(Dinv), PF.c0)), mul2(sp2(P.mass_p), C.c0));
  A.c1 = add2(mul2(sp2(k), mul2(sp2(Dinv), PF.c1)), mul2(sp2(P.mass_p), C.c1));
  return A;
}

// :159-178 after the gather (the statements of g2p_finish<2>); v = gathered APIC velocity on entry
MPM_HD void g2p_finish2(const Params &P, const Material &mat, float dt, f2 &x, f2 &v, const M2c &C, M2c &F, float &Jp,
                        f2 v_in, f2 dv) {
  x = add2(x, mul2(sp2(dt), v));  // :159
  if (P.alpha != 0.0f) {
    const float a = P.alpha;
    v = add2(mul2(sp2(1.0f - a), v), mul2(sp2(a), add2(v_in, dv)));
  }
  M2c A;  // Mat(1) + dt*C, :162
  A.c0 = add2(mk2(1.0f, 0.0f), mul2(sp2(dt), C.c0));
  A.c1 = add2(mk2(0.0f, 1.0f), mul2(sp2(dt), C.c1));
  M2c Fn;
  Fn.c0 = mulvec2(A, F.c0.x, F.c0.y);
  Fn.c1 = mulvec2(A, F.c1.x, F.c1.y);
  if (mat.kind == KIND_SNOW) {
    Mat<2> Fm = to_mat(Fn);
    const float oldJ = mat_det(Fm);  // :172
    plastic_project(mat, Fm);        // :165-173
    Jp = clampf(Jp * oldJ / mat_det(Fm), P.jp_min, P.jp_max);  // :175
    F = to_cols(Fm);
  } else if (mat.kind == KIND_JELLY) {
    F = Fn;
  } else {
    const float s = sqrtf(Fn.c0.x * Fn.c1.y - Fn.c0.y * Fn.c1.x);
    F.c0 = mk2(s, 0.0f);
    F.c1 = mk2(0.0f, s);
  }
}

// Fast separable form of the G2P gather (:147-156), one stencil row (fixed a) at a time:
//   t_a = sum_b wy_b g_ab,  u_a = sum_b (wy_b dy_b) g_ab
//   v += wx_a t_a,  C.col0 += (wx_a dx_a) t_a,  C.col1 += wx_a u_a      (C still lacks the 4*inv_dx of :154)
// `wd[k]` = w[k] * (k - fx) per axis.  FLIP: vo_sum += wx_a * sum_b wy_b vold_ab  (dv = v - vo_sum).
struct Gather2 {
  f2 v, c0, c1, vo;
};
MPM_HD void gather2_row(Gather2 &G, const Sten2 &s, const f2 *wd, int a, f2 g0, f2 g1, f2 g2) {
  f2 t = mul2(sp2(s.w[0].y), g0);
  t = fma2(sp2(s.w[1].y), g1, t);
  t = fma2(sp2(s.w[2].y), g2, t);
  f2 u = mul2(sp2(wd[0].y), g0);
  u = fma2(sp2(wd[1].y), g1, u);
  u = fma2(sp2(wd[2].y), g2, u);
  G.v = fma2(sp2(s.w[a].x), t, G.v);
  G.c0 = fma2(sp2(wd[a].x), t, G.c0);
  G.c1 = fma2(sp2(s.w[a].x), u, G.c1);
}
MPM_HD void gather2_row_old(Gather2 &G, const Sten2 &s, int a, f2 o0, f2 o1, f2 o2) {
  f2 t = mul2(sp2(s.w[0].y), o0);
  t = fma2(sp2(s.w[1].y), o1, t);
  t = fma2(sp2(s.w[2].y), o2, t);
  G.vo = fma2(sp2(s.w[a].x), t, G.vo);
}

}  // namespace mpm
