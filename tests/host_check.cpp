// tests/host_check.cpp -- TEST ONLY.  Runs the PRODUCT's per-particle arithmetic
// (mpm_flip98a_b200/csrc/mpm_math.cuh, the very functions the CUDA kernels call) on the CPU in the
// reference's sequential order, so that tests/test_host_math.py can demand BITWISE equality with the
// oracle without a GPU.  It is not a fallback: nothing in the product links or loads this file.
#include <cstring>
#include <vector>

#include "../mpm_flip98a_b200/csrc/mpm_math.cuh"
#include "../mpm_flip98a_b200/csrc/mpm_math2.cuh"

using namespace mpm;

template <int D>
static void advance(const Params &P, float dt, float *aos, long long n, float *grid /*(D+1) per node*/, float *tap) {
  constexpr int W = 2 * D + 2 * D * D + 2;
  const int N1 = P.n1;
  const size_t nodes = D == 2 ? (size_t)N1 * N1 : (size_t)N1 * N1 * N1;
  std::memset(grid, 0, sizeof(float) * (D + 1) * nodes);
  std::vector<float> vold(P.alpha != 0.0f ? nodes * 3 : 0);
  auto node_of = [&](const int *base, int a, int b, int c) {
    size_t nd = (size_t)(base[0] + a) * N1 + (base[1] + b);
    if (D == 3) nd = nd * N1 + (base[D - 1] + c);
    return nd;
  };
  for (long long pi = 0; pi < n; pi++) {
    float *r = aos + pi * W;
    float *x = r, *v = r + D;
    Mat<D> F, C;
    std::memcpy(&F, r + 2 * D, sizeof F);
    std::memcpy(&C, r + 2 * D + D * D, sizeof C);
    float Jp = r[2 * D + 2 * D * D];
    int c;
    std::memcpy(&c, r + 2 * D + 2 * D * D + 1, 4);
    Stencil<D> st = make_stencil<D>(x, P.inv_dx);
    const Material &mat = P.mat[material_index(P, c)];
    Mat<D> affine = p2g_affine<D>(P, mat, dt, F, C, Jp);
    float mv[D];
    for (int k = 0; k < D; k++) mv[k] = P.mass_p * v[k];
    for (int a = 0; a < 3; a++)
      for (int b = 0; b < 3; b++)
        for (int cc = 0; cc < (D == 3 ? 3 : 1); cc++) {
          float val[D + 1];
          p2g_node_value<D>(P, st, affine, mv, a, b, cc, val);
          float *g = grid + (D + 1) * node_of(st.base, a, b, cc);
          for (int k = 0; k <= D; k++) g[k] = g[k] + val[k];
        }
  }
  if (tap) std::memcpy(tap, grid, sizeof(float) * (D + 1) * nodes);
  for (size_t nd = 0; nd < nodes; nd++) {
    size_t rr = nd;
    int k = 0;
    if (D == 3) {
      k = (int)(rr % N1);
      rr /= N1;
    }
    int j = (int)(rr % N1), i = (int)(rr / N1);
    float g[4] = {0, 0, 0, 0}, vo[3];
    for (int q = 0; q <= D; q++) g[q] = grid[(D + 1) * nd + q];
    if (grid_node_update<D>(P, dt, i, j, k, g, vo))
      for (int q = 0; q <= D; q++) grid[(D + 1) * nd + q] = g[q];
    if (P.alpha != 0.0f)
      for (int q = 0; q < 3; q++) vold[3 * nd + q] = vo[q];
  }
  const bool flip = P.alpha != 0.0f;
  for (long long pi = 0; pi < n; pi++) {
    float *r = aos + pi * W;
    float *x = r, *v = r + D;
    Mat<D> F, C = mat_zero<D>();
    std::memcpy(&F, r + 2 * D, sizeof F);
    float Jp = r[2 * D + 2 * D * D];
    int c;
    std::memcpy(&c, r + 2 * D + 2 * D * D + 1, 4);
    Stencil<D> st = make_stencil<D>(x, P.inv_dx);
    const Material &mat = P.mat[material_index(P, c)];
    float v_in[D], dv[D], vv[D];
    for (int k = 0; k < D; k++) {
      v_in[k] = v[k];
      dv[k] = 0.0f;
      vv[k] = 0.0f;
    }
    for (int a = 0; a < 3; a++)
      for (int b = 0; b < 3; b++)
        for (int cc = 0; cc < (D == 3 ? 3 : 1); cc++) {
          size_t nd = node_of(st.base, a, b, cc);
          float gv[3] = {grid[(D + 1) * nd], grid[(D + 1) * nd + 1], grid[(D + 1) * nd + 2]};
          float vo[3] = {0, 0, 0};
          if (flip)
            for (int q = 0; q < 3; q++) vo[q] = vold[3 * nd + q];
          g2p_accumulate<D>(P, st, a, b, cc, gv, vo, flip, vv, C, dv);
        }
    for (int k = 0; k < D; k++) v[k] = vv[k];
    g2p_finish<D>(P, mat, dt, x, v, C, F, Jp, v_in, dv);
    std::memcpy(r + 2 * D, &F, sizeof F);
    std::memcpy(r + 2 * D + D * D, &C, sizeof C);
    r[2 * D + 2 * D * D] = Jp;
  }
}

// The packed 2D path of the default substep kernel (mpm_math2.cuh: stencil2, affine2, g2p_finish2) in the
// reference's sequential order.  exact_gather != 0: the gather keeps the reference association (g2p_accumulate),
// so the whole substep must be BITWISE the oracle's; exact_gather == 0: the kernel's separable FMA gather
// (gather2_row), which differs at the 1e-7 level.
static void advance_packed2(const Params &P, float dt, float *aos, long long n, float *grid, float *tap, int exact_gather) {
  constexpr int W = 14;
  const int N1 = P.n1;
  const size_t nodes = (size_t)N1 * N1;
  std::memset(grid, 0, sizeof(float) * 3 * nodes);
  const bool flip = P.alpha != 0.0f;
  std::vector<float> vold(flip ? nodes * 2 : 0);
  for (long long pi = 0; pi < n; pi++) {
    float *r = aos + pi * W;
    M2c F, C;
    F.c0 = mk2(r[4], r[5]); F.c1 = mk2(r[6], r[7]);
    C.c0 = mk2(r[8], r[9]); C.c1 = mk2(r[10], r[11]);
    int c;
    std::memcpy(&c, r + 13, 4);
    Sten2 s2 = stencil2(mk2(r[0], r[1]), P.inv_dx);
    const Material &mat = P.mat[material_index(P, c)];
    const M2c A = affine2(P, mat, dt, F, C, r[12]);
    const f2 mvp = mul2(sp2(P.mass_p), mk2(r[2], r[3]));
    Stencil<2> st;
    st.base[0] = s2.bx; st.base[1] = s2.by;
    st.fx[0] = s2.fx.x; st.fx[1] = s2.fx.y;
    for (int k = 0; k < 3; k++) { st.w[k][0] = s2.w[k].x; st.w[k][1] = s2.w[k].y; }
    const Mat<2> affine = to_mat(A);
    const float mv[2] = {mvp.x, mvp.y};
    for (int a = 0; a < 3; a++)
      for (int b = 0; b < 3; b++) {
        float val[3];
        p2g_node_value<2>(P, st, affine, mv, a, b, 0, val);
        float *g = grid + 3 * ((size_t)(st.base[0] + a) * N1 + st.base[1] + b);
        for (int k = 0; k < 3; k++) g[k] = g[k] + val[k];
      }
  }
  if (tap) std::memcpy(tap, grid, sizeof(float) * 3 * nodes);
  for (size_t nd = 0; nd < nodes; nd++) {
    int j = (int)(nd % N1), i = (int)(nd / N1);
    float g[4] = {grid[3 * nd], grid[3 * nd + 1], grid[3 * nd + 2], 0}, vo[3];
    if (grid_node_update<2>(P, dt, i, j, 0, g, vo))
      for (int q = 0; q < 3; q++) grid[3 * nd + q] = g[q];
    if (flip) { vold[2 * nd] = vo[0]; vold[2 * nd + 1] = vo[1]; }
  }
  for (long long pi = 0; pi < n; pi++) {
    float *r = aos + pi * W;
    f2 x = mk2(r[0], r[1]), v_in = mk2(r[2], r[3]);
    M2c F, C;
    F.c0 = mk2(r[4], r[5]); F.c1 = mk2(r[6], r[7]);
    float Jp = r[12];
    int c;
    std::memcpy(&c, r + 13, 4);
    Sten2 s2 = stencil2(x, P.inv_dx);
    const Material &mat = P.mat[material_index(P, c)];
    f2 v, dv;
    if (exact_gather) {
      Stencil<2> st;
      st.base[0] = s2.bx; st.base[1] = s2.by;
      st.fx[0] = s2.fx.x; st.fx[1] = s2.fx.y;
      for (int k = 0; k < 3; k++) { st.w[k][0] = s2.w[k].x; st.w[k][1] = s2.w[k].y; }
      float vv[2] = {0, 0}, dvv[2] = {0, 0};
      Mat<2> Cm = mat_zero<2>();
      for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++) {
          size_t nd = (size_t)(st.base[0] + a) * N1 + st.base[1] + b;
          float gv[3] = {grid[3 * nd], grid[3 * nd + 1], grid[3 * nd + 2]};
          float vo[3] = {flip ? vold[2 * nd] : 0.0f, flip ? vold[2 * nd + 1] : 0.0f, 0.0f};
          g2p_accumulate<2>(P, st, a, b, 0, gv, vo, flip, vv, Cm, dvv);
        }
      v = mk2(vv[0], vv[1]);
      dv = mk2(dvv[0], dvv[1]);
      C = to_cols(Cm);
    } else {
      f2 wd[3];
      wd[0] = mul2(s2.w[0], sub2(sp2(0.0f), s2.fx));
      wd[1] = mul2(s2.w[1], sub2(sp2(1.0f), s2.fx));
      wd[2] = mul2(s2.w[2], sub2(sp2(2.0f), s2.fx));
      Gather2 g;
      g.v = g.c0 = g.c1 = g.vo = sp2(0.0f);
      for (int a = 0; a < 3; a++) {
        f2 gg[3], oo[3];
        for (int b = 0; b < 3; b++) {
          size_t nd = (size_t)(s2.bx + a) * N1 + s2.by + b;
          gg[b] = mk2(grid[3 * nd], grid[3 * nd + 1]);
          oo[b] = flip ? mk2(vold[2 * nd], vold[2 * nd + 1]) : sp2(0.0f);
        }
        gather2_row(g, s2, wd, a, gg[0], gg[1], gg[2]);
        if (flip) gather2_row_old(g, s2, a, oo[0], oo[1], oo[2]);
      }
      const float s4 = 4 * P.inv_dx;
      v = g.v;
      dv = sub2(g.v, g.vo);
      C.c0 = mul2(sp2(s4), g.c0);
      C.c1 = mul2(sp2(s4), g.c1);
    }
    g2p_finish2(P, mat, dt, x, v, C, F, Jp, v_in, dv);
    r[0] = x.x; r[1] = x.y; r[2] = v.x; r[3] = v.y;
    r[4] = F.c0.x; r[5] = F.c0.y; r[6] = F.c1.x; r[7] = F.c1.y;
    r[8] = C.c0.x; r[9] = C.c0.y; r[10] = C.c1.x; r[11] = C.c1.y;
    r[12] = Jp;
  }
}

extern "C" {
int hostcheck_advance_packed2(const Params *P, float dt, float *aos, long long n, int n_steps, float *grid, float *tap,
                              int exact_gather) {
  for (int s = 0; s < n_steps; s++) advance_packed2(*P, dt, aos, n, grid, s == n_steps - 1 ? tap : nullptr, exact_gather);
  return 0;
}
int hostcheck_params_bytes() { return (int)sizeof(Params); }
// mats: n_materials x (kind, E, nu, hardening, sig_lo, sig_hi) as floats (kind as float value)
void hostcheck_make_params(Params *P, int n_grid, float mass_p, float vol_p, const float *gravity, float boundary,
                           float jp_min, float jp_max, float alpha, int n_materials, const float *mats) {
  std::memset(P, 0, sizeof *P);
  P->n_grid = n_grid;
  P->n1 = n_grid + 1;
  P->dx = 1.0f / n_grid;
  P->inv_dx = 1.0f / P->dx;
  P->mass_p = mass_p;
  P->vol_p = vol_p;
  for (int k = 0; k < 3; k++) P->gravity[k] = gravity[k];
  P->boundary = boundary;
  P->jp_min = jp_min;
  P->jp_max = jp_max;
  P->alpha = alpha;
  P->n_materials = n_materials;
  for (int m = 0; m < n_materials; m++) {
    const float *s = mats + 6 * m;
    Material &d = P->mat[m];
    d.kind = (int)s[0];
    volatile float E = s[1], nu = s[2];
    d.mu_0 = E / (2 * (1 + nu));
    d.lambda_0 = E * nu / ((1 + nu) * (1 - 2 * nu));
    d.hardening = s[3];
    d.sig_lo = s[4];
    d.sig_hi = s[5];
  }
  P->slab_lo = 0;
  P->slab_hi = n_grid;
  P->ncol = n_grid + 1;
}
int hostcheck_advance(const Params *P, int dim, float dt, float *aos, long long n, int n_steps, float *grid, float *tap) {
  for (int s = 0; s < n_steps; s++) {
    float *t = s == n_steps - 1 ? tap : nullptr;
    if (dim == 2) advance<2>(*P, dt, aos, n, grid, t);
    else advance<3>(*P, dt, aos, n, grid, t);
  }
  return 0;
}
void hostcheck_svd2(const float *m, float *U, float *sig, float *V) {
  Mat<2> M, u = mat_zero<2>(), s = mat_zero<2>(), v = mat_zero<2>();
  std::memcpy(&M, m, 16);
  svd2(M, u, s, v);
  std::memcpy(U, &u, 16);
  std::memcpy(sig, &s, 16);
  std::memcpy(V, &v, 16);
}
void hostcheck_polar2(const float *m, float *R, float *S) {
  Mat<2> M, r = mat_zero<2>(), s = mat_zero<2>();
  std::memcpy(&M, m, 16);
  polar2(M, r, s);
  std::memcpy(R, &r, 16);
  std::memcpy(S, &s, 16);
}
void hostcheck_rotation3(const float *m, float *R_svd, float *R_newton) {
  Mat<3> M;
  std::memcpy(&M, m, 36);
  Mat<3> a = rotation_of_svd(M), b = rotation_of(M);
  std::memcpy(R_svd, &a, 36);
  std::memcpy(R_newton, &b, 36);
}
float hostcheck_plastic_project3(float lo, float hi, float *m, float *R) {
  Mat<3> M, r;
  std::memcpy(&M, m, 36);
  const float ratio = plastic_project3(lo, hi, M, &r);
  std::memcpy(m, &M, 36);
  std::memcpy(R, &r, 36);
  return ratio;
}
void hostcheck_svd3(const float *m, float *U, float *sig3, float *V) {
  Mat<3> M, u, v;
  std::memcpy(&M, m, 36);
  svd3(M, u, sig3, v);
  std::memcpy(U, &u, 36);
  std::memcpy(V, &v, 36);
}
}
