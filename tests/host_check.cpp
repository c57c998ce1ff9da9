// tests/host_check.cpp -- TEST ONLY.  Runs the PRODUCT's per-particle arithmetic
// (mpm_flip98a_b200/csrc/mpm_math.cuh, the very functions the CUDA kernels call) on the CPU in the
// reference's sequential order, so that tests/test_host_math.py can demand BITWISE equality with the
// oracle without a GPU.  It is not a fallback: nothing in the product links or loads this file.
#include <cstring>
#include <vector>

#include "../mpm_flip98a_b200/csrc/mpm_math.cuh"

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

extern "C" {
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
void hostcheck_svd3(const float *m, float *U, float *sig3, float *V) {
  Mat<3> M, u, v;
  std::memcpy(&M, m, 36);
  svd3(M, u, sig3, v);
  std::memcpy(U, &u, 36);
  std::memcpy(V, &v, 36);
}
}
