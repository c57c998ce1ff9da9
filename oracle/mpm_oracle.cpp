// oracle/mpm_oracle.cpp -- CPU restatement of the reference substep.
//
// *** TEST INFRASTRUCTURE ONLY. ***  Nothing under mpm_flip98a_b200/ may include, link or
// call this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// `--impl reference` legs load the library built from it (oracle/build/liboracle.so).
//
// What it restates (all citations relative to /root/reference/):
//   * advance(real dt)                  cpp_validation/mls-mpm88-explained.cpp:49-180
//   * polar_decomp / svd for Matrix2    cpp_validation/taichi.h:8375-8420
//   * determinant, clamp, cast<int>     cpp_validation/taichi.h:7850-7852, 6449-6455, 7185-7187
//   * xorshift128 rand + Vec::rand      cpp_validation/taichi.h:6497-6513, 7317-7323
//   * scene seeding add_object          cpp_validation/mls-mpm88-explained.cpp:191-196
// Every expression keeps the reference's association order and is compiled with
// -ffp-contract=off on plain x86-64 (no FMA), like the reference's own build line
// (mls-mpm88-explained.cpp:233).  PINNING: tests/test_oracle_pinning.py requires this file to
// be BITWISE identical to the unmodified reference (oracle/_ref/libmpmref.so, built from the
// reference sources where they lie) on the shipped scene for 2500 substeps, and against the
// golden vectors under tests/golden/ that were generated from that library.
//
// Extensions that the reference does NOT contain and that are therefore "parity unpinned"
// (north_star asks for them):
//   * material kinds fluid / jelly (snow == the shipped path)         -> see material_* below
//   * FLIP/PIC blend alpha (alpha == 0 takes the reference statements verbatim)
//   * dim == 3 lift (27-node stencil, one-sided Jacobi 3x3 SVD)
#include <cmath>
#include <cstdint>
#include <cstring>
#include <algorithm>
#include <thread>
#include <vector>

namespace {

enum { KIND_FLUID = 0, KIND_JELLY = 1, KIND_SNOW = 2 };

struct OracleMaterial {
  int kind;
  float E, nu, hardening;
  float sig_lo, sig_hi;  // plastic clamp of the singular values (reference: 1-2.5e-2, 1+7.5e-3)
};

struct OracleParams {
  int dim;       // 2 or 3
  int n_grid;    // cells per axis; nodes per axis = n_grid + 1
  float mass_p;  // :17
  float vol_p;   // :18
  float gravity[3];  // :113 hard-codes (0,-200,0)
  float boundary;    // :116 hard-codes 0.05
  float jp_min, jp_max;  // :175 hard-codes 0.6, 20
  float alpha;           // FLIP blend; 0 == reference
  int n_materials;
  OracleMaterial mat[4];
};

// ---------------------------------------------------------------------------------------------
// 2x2 helpers.  Matrices are stored like the reference's MatrixND: column-major, m[c][r]
// (taichi.h:7575: operator()(i,j) == d[j][i]).
// ---------------------------------------------------------------------------------------------
struct M2 {
  float d[2][2];  // d[col][row]
  float &operator()(int i, int j) { return d[j][i]; }
  float operator()(int i, int j) const { return d[j][i]; }
};

inline M2 m2_zero() {
  M2 r;
  r.d[0][0] = r.d[0][1] = r.d[1][0] = r.d[1][1] = 0.0f;
  return r;
}
inline M2 m2_diag(float v) {  // taichi.h:7504 MatrixND(T v)
  M2 r = m2_zero();
  r.d[0][0] = v;
  r.d[1][1] = v;
  return r;
}
inline M2 m2_add(const M2 &a, const M2 &b) {
  M2 r;
  for (int c = 0; c < 2; c++)
    for (int k = 0; k < 2; k++) r.d[c][k] = a.d[c][k] + b.d[c][k];
  return r;
}
inline M2 m2_sub(const M2 &a, const M2 &b) {
  M2 r;
  for (int c = 0; c < 2; c++)
    for (int k = 0; k < 2; k++) r.d[c][k] = a.d[c][k] - b.d[c][k];
  return r;
}
inline M2 m2_scale(float s, const M2 &a) {  // taichi.h:7806 (a * M[i] == Vec(a) * M[i])
  M2 r;
  for (int c = 0; c < 2; c++)
    for (int k = 0; k < 2; k++) r.d[c][k] = s * a.d[c][k];
  return r;
}
inline M2 m2_transposed(const M2 &a) {  // taichi.h:7696
  M2 r;
  for (int i = 0; i < 2; i++)
    for (int j = 0; j < 2; j++) r.d[i][j] = a.d[j][i];
  return r;
}
// taichi.h:7591-7597  ret = d[0]*o[0]; ret += d[1]*o[1]
inline void m2_mulvec(const M2 &a, const float v[2], float out[2]) {
  for (int k = 0; k < 2; k++) {
    float r = a.d[0][k] * v[0];
    r = r + a.d[1][k] * v[1];
    out[k] = r;
  }
}
inline M2 m2_mul(const M2 &a, const M2 &b) {  // taichi.h:7639 column i = a * b[i]
  M2 r;
  for (int c = 0; c < 2; c++) m2_mulvec(a, b.d[c], r.d[c]);
  return r;
}
inline float m2_det(const M2 &m) {  // taichi.h:7850-7852
  return m.d[0][0] * m.d[1][1] - m.d[0][1] * m.d[1][0];
}
inline float clampf(float a, float lo, float hi) {  // taichi.h:6449-6455
  if (a < lo) return lo;
  if (a > hi) return hi;
  return a;
}

// taichi.h:8375-8385
inline void polar2(const M2 &m, M2 &R, M2 &S) {
  float x = m(0, 0) + m(1, 1);
  float y = m(1, 0) - m(0, 1);
  float scale = 1.0f / std::sqrt(x * x + y * y);
  float c = x * scale, s = y * scale;
  R(0, 0) = c;
  R(0, 1) = -s;
  R(1, 0) = s;
  R(1, 1) = c;
  S = m2_mul(m2_transposed(R), m);
}

// taichi.h:8389-8420
inline void svd2(const M2 &m, M2 &U, M2 &sig, M2 &V) {
  M2 S;
  polar2(m, U, S);
  float c, s;
  if (std::abs(S(0, 1)) < 1e-6f) {
    sig = S;
    c = 1;
    s = 0;
  } else {
    float tao = 0.5f * (S(0, 0) - S(1, 1));
    float w = std::sqrt(tao * tao + S(0, 1) * S(0, 1));
    float t = tao > 0 ? S(0, 1) / (tao + w) : S(0, 1) / (tao - w);
    c = 1.0f / std::sqrt(t * t + 1);
    s = -t * c;
    sig(0, 0) = (c * c) * S(0, 0) - 2 * c * s * S(0, 1) + (s * s) * S(1, 1);
    sig(1, 1) = (s * s) * S(0, 0) + 2 * c * s * S(0, 1) + (c * c) * S(1, 1);
  }
  if (sig(0, 0) < sig(1, 1)) {
    std::swap(sig(0, 0), sig(1, 1));
    V(0, 0) = -s;
    V(0, 1) = -c;
    V(1, 0) = c;
    V(1, 1) = -s;
  } else {
    V(0, 0) = c;
    V(0, 1) = -s;
    V(1, 0) = s;
    V(1, 1) = c;
  }
  V = m2_transposed(V);
  U = m2_mul(U, V);
}

// Lame parameters, fp32, the reference's expression order (mls-mpm88-explained.cpp:25-26).
inline void lame(float E, float nu, float &mu0, float &lambda0) {
  mu0 = E / (2 * (1 + nu));
  lambda0 = E * nu / ((1 + nu) * (1 - 2 * nu));
}

// ---------------------------------------------------------------------------------------------
// 2D substep.  Particle record = the reference's 56-byte struct (:28-42):
//   x[2] v[2] F[4 col-major] C[4 col-major] Jp  c(int; material id here)
// grid node = (vx, vy, m) like :46-47; `vold` (2 floats/node) only when alpha != 0.
// ---------------------------------------------------------------------------------------------
struct P2 {
  float x[2], v[2];
  M2 F, C;
  float Jp;
  int c;
};
static_assert(sizeof(P2) == 56, "2D particle record must match the reference (56 B)");

inline int material_of(const OracleParams &P, int c) {
  // material id == Particle::c when it is a valid table index, else the LAST table entry
  // (lets the shipped scene, whose c is a colour 0x2986CC, select the "shipped" material).
  if (c >= 0 && c < P.n_materials) return c;
  return P.n_materials - 1;
}

// ---------------------------------------------------------------------------------------------
// Optional host threading (oracle_advance_mt).  BITWISE identical to the serial loops:
//   * P2G: thread t owns the node columns [cut[t], cut[t+1]) and walks ALL particles in index order,
//     adding only the contributions that land in its columns -- so every node still receives its
//     contributions in particle-index order, exactly like the serial loop (:53-102);
//   * grid update and G2P touch disjoint nodes / particles and are split by range.
// n_threads == 1 runs the loops on the calling thread (the reference is single-threaded as shipped).
// ---------------------------------------------------------------------------------------------
template <class F>
void run_threads(int T, F f) {
  std::vector<std::thread> th;
  for (int t = 1; t < T; t++) th.emplace_back(f, t);
  f(0);
  for (auto &x : th) x.join();
}
// base x-column of every particle (:55) and node-column cuts that balance the particle counts
template <class PT>
void plan_columns(const PT *particles, long long n, float inv_dx, int N1, int T, std::vector<int> &bx,
                  std::vector<int> &cut) {
  bx.resize((size_t)n);
  run_threads(T, [&](int t) {
    long long lo = n * t / T, hi = n * (t + 1) / T;
    for (long long pi = lo; pi < hi; pi++) bx[(size_t)pi] = (int)(particles[pi].x[0] * inv_dx - 0.5f);
  });
  std::vector<long long> hist((size_t)N1 + 1, 0);
  for (long long pi = 0; pi < n; pi++) {
    int b = bx[(size_t)pi];
    b = b < 0 ? 0 : (b >= N1 ? N1 - 1 : b);
    hist[(size_t)b]++;
  }
  cut.assign((size_t)T + 1, N1);
  cut[0] = 0;
  long long run = 0;
  int t = 1;
  for (int c = 0; c < N1 && t < T; c++) {
    run += hist[(size_t)c];
    while (t < T && run >= n * t / T) cut[(size_t)t++] = c + 1;
  }
  for (int k = 1; k <= T; k++) cut[(size_t)k] = std::max(cut[(size_t)k], cut[(size_t)k - 1]);
  cut[(size_t)T] = N1;
}

void advance2(const OracleParams &P, float dt, P2 *particles, long long n, float *grid /*(n+1)^2*3*/,
              float *grid_post_p2g /*nullable*/, float *vold /*(n+1)^2*2, nullable unless alpha!=0*/,
              int n_threads = 1) {
  const int num_grid = P.n_grid;
  const int N1 = num_grid + 1;
  const float dx = 1.0f / num_grid;  // :12
  const float inv_dx = 1.0f / dx;    // :13
  const float mass_p = P.mass_p, vol_p = P.vol_p;
  float mu_0[4], lambda_0[4];
  for (int m = 0; m < P.n_materials; m++) lame(P.mat[m].E, P.mat[m].nu, mu_0[m], lambda_0[m]);
  const bool flip = P.alpha != 0.0f;
  const int T = n_threads > 1 ? n_threads : 1;
  std::vector<int> bx, cut;
  if (T > 1) plan_columns(particles, n, inv_dx, N1, T, bx, cut);
  else cut = {0, N1};

  run_threads(T, [&](int t) {
  const int col_lo = cut[(size_t)t], col_hi = cut[(size_t)t + 1];  // node columns this thread accumulates
  std::memset(grid + 3 * (size_t)col_lo * N1, 0, sizeof(float) * 3 * (size_t)(col_hi - col_lo) * N1);  // :50

  // P2G :53-102
  for (long long pi = 0; pi < n; pi++) {
    if (T > 1) {
      const int b = bx[(size_t)pi];
      if (b + 2 < col_lo || b >= col_hi) continue;
    }
    P2 &p = particles[pi];
    const int mid = material_of(P, p.c);
    const OracleMaterial &mat = P.mat[mid];
    int base[2];
    float fx[2];
    for (int k = 0; k < 2; k++) {
      base[k] = (int)(p.x[k] * inv_dx - 0.5f);      // :55 (static_cast == truncation)
      fx[k] = p.x[k] * inv_dx - (float)base[k];     // :57
    }
    float w[3][2];
    for (int k = 0; k < 2; k++) {                   // :60-64
      w[0][k] = 0.5f * ((1.5f - fx[k]) * (1.5f - fx[k]));
      w[1][k] = 0.75f - ((fx[k] - 1.0f) * (fx[k] - 1.0f));
      w[2][k] = 0.5f * ((fx[k] - 0.5f) * (fx[k] - 0.5f));
    }
    // :67-69 hardening.  snow: exp(h(1-Jp)); jelly: constant factor; fluid: 1 and mu = 0.
    float e;
    if (mat.kind == KIND_SNOW) e = std::exp(mat.hardening * (1.0f - p.Jp));
    else if (mat.kind == KIND_JELLY) e = mat.hardening;
    else e = 1.0f;
    float mu = mu_0[mid] * e;
    float lambda = lambda_0[mid] * e;
    float J = m2_det(p.F);  // :72
    float Dinv = 4 * inv_dx * inv_dx;  // :79
    M2 PF;
    if (mat.kind == KIND_FLUID) {
      // mu == 0: only the volumetric term survives (keeps the reference's "+ scalar" form)
      PF = m2_diag(lambda * (J - 1) * J);
    } else {
      M2 r, s;
      polar2(p.F, r, s);  // :75-76
      // :81  2*mu*(F-r)*F^T + lambda*(J-1)*J   (scalar promoted to v*I, taichi.h:7504)
      PF = m2_add(m2_mul(m2_scale(2 * mu, m2_sub(p.F, r)), m2_transposed(p.F)), m2_diag(lambda * (J - 1) * J));
    }
    M2 stress = m2_scale(-(dt * vol_p), m2_scale(Dinv, PF));  // :84
    M2 affine = m2_add(stress, m2_scale(mass_p, p.C));        // :89
    for (int i = 0; i < 3; i++) {
      for (int j = 0; j < 3; j++) {  // :92-101
        float dpos[2] = {((float)i - fx[0]) * dx, ((float)j - fx[1]) * dx};
        float mv[3] = {mass_p * p.v[0], mass_p * p.v[1], mass_p};
        float ad[2];
        m2_mulvec(affine, dpos, ad);
        float wgt = w[i][0] * w[j][1];
        if (T > 1 && (base[0] + i < col_lo || base[0] + i >= col_hi)) continue;  // another thread's column
        float *g = grid + 3 * ((size_t)(base[0] + i) * N1 + (base[1] + j));
        g[0] = g[0] + wgt * (mv[0] + ad[0]);
        g[1] = g[1] + wgt * (mv[1] + ad[1]);
        g[2] = g[2] + wgt * (mv[2] + 0.0f);
      }
    }
  }
  });
  if (grid_post_p2g) std::memcpy(grid_post_p2g, grid, sizeof(float) * 3 * (size_t)N1 * N1);

  // grid update :105-131
  run_threads(T, [&](int t) {
  for (int i = (int)((long long)N1 * t / T); i < (int)((long long)N1 * (t + 1) / T); i++) {
    for (int j = 0; j <= num_grid; j++) {
      float *g = grid + 3 * ((size_t)i * N1 + j);
      if (flip) {
        vold[2 * ((size_t)i * N1 + j) + 0] = 0.0f;
        vold[2 * ((size_t)i * N1 + j) + 1] = 0.0f;
      }
      if (g[2] > 0) {
        float m = g[2];
        g[0] = g[0] / m;  // :111
        g[1] = g[1] / m;
        g[2] = g[2] / m;
        if (flip) {
          vold[2 * ((size_t)i * N1 + j) + 0] = g[0];
          vold[2 * ((size_t)i * N1 + j) + 1] = g[1];
        }
        g[0] = g[0] + dt * P.gravity[0];  // :113
        g[1] = g[1] + dt * P.gravity[1];
        g[2] = g[2] + dt * 0.0f;
        float boundary = P.boundary;  // :116
        float x = (float)i / num_grid;  // :118
        float y = (float)j / num_grid;
        if (x < boundary || x > 1 - boundary || y > 1 - boundary) {  // :122
          g[0] = 0.0f;
          g[1] = 0.0f;
          g[2] = 0.0f;
        }
        if (y < boundary) {  // :126
          g[1] = std::max(0.0f, g[1]);
        }
      }
    }
  }
  });

  // G2P :134-179
  run_threads(T, [&](int t) {
  for (long long pi = n * t / T; pi < n * (t + 1) / T; pi++) {
    P2 &p = particles[pi];
    const int mid = material_of(P, p.c);
    const OracleMaterial &mat = P.mat[mid];
    int base[2];
    float fx[2];
    for (int k = 0; k < 2; k++) {
      base[k] = (int)(p.x[k] * inv_dx - 0.5f);
      fx[k] = p.x[k] * inv_dx - (float)base[k];
    }
    float w[3][2];
    for (int k = 0; k < 2; k++) {
      w[0][k] = 0.5f * ((1.5f - fx[k]) * (1.5f - fx[k]));
      w[1][k] = 0.75f - ((fx[k] - 1.0f) * (fx[k] - 1.0f));
      w[2][k] = 0.5f * ((fx[k] - 0.5f) * (fx[k] - 0.5f));
    }
    float v_in[2] = {p.v[0], p.v[1]};
    p.C = m2_zero();  // :144
    p.v[0] = 0.0f;    // :145
    p.v[1] = 0.0f;
    float dv[2] = {0.0f, 0.0f};
    for (int i = 0; i < 3; i++) {
      for (int j = 0; j < 3; j++) {  // :147-156
        float dpos[2] = {(float)i - fx[0], (float)j - fx[1]};
        size_t node = (size_t)(base[0] + i) * N1 + (base[1] + j);
        float gv[2] = {grid[3 * node + 0], grid[3 * node + 1]};
        float weight = w[i][0] * w[j][1];
        float wg[2] = {weight * gv[0], weight * gv[1]};
        p.v[0] = p.v[0] + wg[0];  // :153
        p.v[1] = p.v[1] + wg[1];
        // :154  C += 4*inv_dx * outer_product(weight*grid_v, dpos); column c = wg * dpos[c]
        float s4 = 4 * inv_dx;
        for (int c = 0; c < 2; c++)
          for (int k = 0; k < 2; k++) p.C.d[c][k] = p.C.d[c][k] + s4 * (wg[k] * dpos[c]);
        if (flip) {
          dv[0] = dv[0] + weight * (gv[0] - vold[2 * node + 0]);
          dv[1] = dv[1] + weight * (gv[1] - vold[2 * node + 1]);
        }
      }
    }
    // Advection :159 (with the grid (APIC/PIC) velocity, also when blending)
    p.x[0] = p.x[0] + dt * p.v[0];
    p.x[1] = p.x[1] + dt * p.v[1];
    if (flip) {
      float a = P.alpha;
      p.v[0] = (1.0f - a) * p.v[0] + a * (v_in[0] + dv[0]);
      p.v[1] = (1.0f - a) * p.v[1] + a * (v_in[1] + dv[1]);
    }
    // :162 F = (Mat(1) + dt*C) * F
    M2 F = m2_mul(m2_add(m2_diag(1.0f), m2_scale(dt, p.C)), p.F);
    if (mat.kind == KIND_SNOW) {
      M2 svd_u = m2_zero(), sig = m2_zero(), svd_v = m2_zero();
      svd2(F, svd_u, sig, svd_v);  // :165
      for (int i = 0; i < 2; i++) sig.d[i][i] = clampf(sig.d[i][i], mat.sig_lo, mat.sig_hi);  // :168-170
      float oldJ = m2_det(F);                                     // :172
      F = m2_mul(m2_mul(svd_u, sig), m2_transposed(svd_v));        // :173
      float Jp_new = clampf(p.Jp * oldJ / m2_det(F), P.jp_min, P.jp_max);  // :175
      p.Jp = Jp_new;
      p.F = F;
    } else if (mat.kind == KIND_JELLY) {
      p.F = F;  // elastic: no return mapping, Jp untouched
    } else {
      // fluid: keep only the volume change, F <- sqrt(J) * I (taichi mpm99 convention)
      float J = m2_det(F);
      p.F = m2_diag(std::sqrt(J));
    }
  }
  });
}

// ---------------------------------------------------------------------------------------------
// 3D lift (not in the reference; "parity unpinned").  Same statement order as advance2 with
// Vec3/Mat3, 27-node stencil; polar/SVD by a fixed-sweep one-sided Jacobi.
// Particle record (104 B): x[3] v[3] F[9 col-major] C[9 col-major] Jp c.
// ---------------------------------------------------------------------------------------------
struct M3 {
  float d[3][3];  // d[col][row]
  float &operator()(int i, int j) { return d[j][i]; }
  float operator()(int i, int j) const { return d[j][i]; }
};
struct P3 {
  float x[3], v[3];
  M3 F, C;
  float Jp;
  int c;
};
static_assert(sizeof(P3) == 104, "3D particle record is 104 B");

inline M3 m3_zero() {
  M3 r;
  std::memset(&r, 0, sizeof(r));
  return r;
}
inline M3 m3_diag(float v) {
  M3 r = m3_zero();
  r.d[0][0] = r.d[1][1] = r.d[2][2] = v;
  return r;
}
inline M3 m3_add(const M3 &a, const M3 &b) {
  M3 r;
  for (int c = 0; c < 3; c++)
    for (int k = 0; k < 3; k++) r.d[c][k] = a.d[c][k] + b.d[c][k];
  return r;
}
inline M3 m3_sub(const M3 &a, const M3 &b) {
  M3 r;
  for (int c = 0; c < 3; c++)
    for (int k = 0; k < 3; k++) r.d[c][k] = a.d[c][k] - b.d[c][k];
  return r;
}
inline M3 m3_scale(float s, const M3 &a) {
  M3 r;
  for (int c = 0; c < 3; c++)
    for (int k = 0; k < 3; k++) r.d[c][k] = s * a.d[c][k];
  return r;
}
inline M3 m3_transposed(const M3 &a) {
  M3 r;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) r.d[i][j] = a.d[j][i];
  return r;
}
inline void m3_mulvec(const M3 &a, const float v[3], float out[3]) {
  for (int k = 0; k < 3; k++) {
    float r = a.d[0][k] * v[0];
    r = r + a.d[1][k] * v[1];
    r = r + a.d[2][k] * v[2];
    out[k] = r;
  }
}
inline M3 m3_mul(const M3 &a, const M3 &b) {
  M3 r;
  for (int c = 0; c < 3; c++) m3_mulvec(a, b.d[c], r.d[c]);
  return r;
}
inline float m3_det(const M3 &mat) {  // taichi.h:7855-7859 (mat[c][r])
  return mat.d[0][0] * (mat.d[1][1] * mat.d[2][2] - mat.d[2][1] * mat.d[1][2]) -
         mat.d[1][0] * (mat.d[0][1] * mat.d[2][2] - mat.d[2][1] * mat.d[0][2]) +
         mat.d[2][0] * (mat.d[0][1] * mat.d[1][2] - mat.d[1][1] * mat.d[0][2]);
}

// One-sided (Hestenes) Jacobi SVD, 4 fixed cyclic sweeps over the column pairs (0,1),(0,2),(1,2):
// A V = U Sigma with A's columns made mutually orthogonal.  Afterwards the singular values are
// sorted descending and signs are fixed so that det U = det V = +1 (sigma_2 carries the sign of
// det A).  The CUDA kernel restates exactly this sequence.
inline void svd3(const M3 &A_in, M3 &U, float sig[3], M3 &V) {
  M3 A = A_in;
  V = m3_diag(1.0f);
  const int pairs[3][2] = {{0, 1}, {0, 2}, {1, 2}};
  for (int sweep = 0; sweep < 4; sweep++) {
    for (int pr = 0; pr < 3; pr++) {
      const int p = pairs[pr][0], q = pairs[pr][1];
      float a = A.d[p][0] * A.d[p][0] + A.d[p][1] * A.d[p][1] + A.d[p][2] * A.d[p][2];
      float b = A.d[q][0] * A.d[q][0] + A.d[q][1] * A.d[q][1] + A.d[q][2] * A.d[q][2];
      float c = A.d[p][0] * A.d[q][0] + A.d[p][1] * A.d[q][1] + A.d[p][2] * A.d[q][2];
      // converged pair: the columns are orthogonal to fp32 resolution (|cos| <= 2 ulp); later sweeps then only pay the test
      if (std::abs(c) <= 2.4e-7f * std::sqrt(a * b)) continue;
      float zeta = (b - a) / (2.0f * c);
      float t = (zeta >= 0.0f ? 1.0f : -1.0f) / (std::abs(zeta) + std::sqrt(1.0f + zeta * zeta));
      float cs = 1.0f / std::sqrt(1.0f + t * t);
      float sn = cs * t;
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
  for (int c = 0; c < 3; c++) s[c] = std::sqrt(A.d[c][0] * A.d[c][0] + A.d[c][1] * A.d[c][1] + A.d[c][2] * A.d[c][2]);
  // sort descending (3-element network), permuting columns of A and V alike
  auto swap_cols = [&](int i, int j) {
    std::swap(s[i], s[j]);
    for (int k = 0; k < 3; k++) {
      std::swap(A.d[i][k], A.d[j][k]);
      std::swap(V.d[i][k], V.d[j][k]);
    }
  };
  if (s[0] < s[1]) swap_cols(0, 1);
  if (s[0] < s[2]) swap_cols(0, 2);
  if (s[1] < s[2]) swap_cols(1, 2);
  for (int c = 0; c < 3; c++) {
    float inv = s[c] > 0.0f ? 1.0f / s[c] : 0.0f;
    for (int k = 0; k < 3; k++) U.d[c][k] = A.d[c][k] * inv;
  }
  if (m3_det(V) < 0.0f) {  // make V a rotation: flip its last column (and U's, keeping A = U S V^T)
    for (int k = 0; k < 3; k++) {
      V.d[2][k] = -V.d[2][k];
      U.d[2][k] = -U.d[2][k];
    }
  }
  if (m3_det(U) < 0.0f) {  // make U a rotation: the sign moves into sigma_2
    for (int k = 0; k < 3; k++) U.d[2][k] = -U.d[2][k];
    s[2] = -s[2];
  }
  sig[0] = s[0];
  sig[1] = s[1];
  sig[2] = s[2];
}

// Rotation factor of the 3D stress: Newton's iteration for the polar decomposition, X <- (X + X^-T)/2 from X0 = F,
// stopped when an iteration moves no entry by more than 4e-4 (the iterate it produced is then within 8e-8 of R: the
// convergence is quadratic) -- two iterations for snow, whose F is within 2.5 % of a rotation after the plastic clamp.
// Returns false for a near-singular or inverted F (det <= 1e-6 |F|^3), which the callers hand to the Jacobi SVD above.
// (Round 2: this replaced the SVD's U V^T in the stress -- 20x cheaper on the GPU; both are this builder's
// definitions, the reference has no 3D code, and they agree to ~1e-6.)
inline float cof(float p, float q, float r, float t) { return std::fma(p, q, -(r * t)); }  // p q - r t, one rounding on the first product
inline bool polar_newton3(const M3 &F, M3 &X) {
  X = F;
  float scale = 0.0f;
  for (int c = 0; c < 3; c++)
    for (int k = 0; k < 3; k++) scale = std::fmax(scale, std::fabs(F.d[c][k]));
  for (int it = 0; it < 12; it++) {
    const float *a = X.d[0], *b = X.d[1], *c = X.d[2];
    M3 K;  // cofactors: columns b x c, c x a, a x b  (X^-T = K / det)
    K.d[0][0] = cof(b[1], c[2], b[2], c[1]); K.d[0][1] = cof(b[2], c[0], b[0], c[2]); K.d[0][2] = cof(b[0], c[1], b[1], c[0]);
    K.d[1][0] = cof(c[1], a[2], c[2], a[1]); K.d[1][1] = cof(c[2], a[0], c[0], a[2]); K.d[1][2] = cof(c[0], a[1], c[1], a[0]);
    K.d[2][0] = cof(a[1], b[2], a[2], b[1]); K.d[2][1] = cof(a[2], b[0], a[0], b[2]); K.d[2][2] = cof(a[0], b[1], a[1], b[0]);
    const float det = std::fma(a[2], K.d[0][2], std::fma(a[1], K.d[0][1], a[0] * K.d[0][0]));
    if (!(det > 1e-6f * scale * scale * scale)) return false;
    const float h = 0.5f / det;
    float delta = 0.0f;
    for (int cc = 0; cc < 3; cc++)
      for (int k = 0; k < 3; k++) {
        const float y = std::fma(h, K.d[cc][k], 0.5f * X.d[cc][k]);
        delta = std::fmax(delta, std::fabs(y - X.d[cc][k]));
        X.d[cc][k] = y;
      }
    if (delta <= 4e-4f) break;
  }
  return true;
}
inline M3 rotation3(const M3 &F) {
  M3 X;
  if (polar_newton3(F, X)) return X;
  M3 U, V;
  float sg[3];
  svd3(F, U, sg, V);
  return m3_mul(U, m3_transposed(V));
}

inline float dot3f(const float *a, const float *b) { return std::fma(a[2], b[2], std::fma(a[1], b[1], a[0] * b[0])); }

// Plastic projection of snow in 3D (:165-178 lifted): clamp the singular values of F to [lo, hi], return
// det(F) / det(F') for the Jp update (:175).  Same structure as the reference's 2x2 svd (taichi.h:8389-8420: polar
// decomposition first, then Jacobi rotations that diagonalise the symmetric factor):
//   F = R S (polar_newton3),  S = V diag(l) V^T (cyclic Jacobi on the symmetric 3x3),  F' = R V diag(clamp(l)) V^T.
// When Gershgorin's discs already place every eigenvalue of S inside [lo, hi] nothing is clamped and F is returned
// untouched (ratio 1).  Near-singular or inverted F takes the one-sided Jacobi SVD (svd3).  Round 2: replaced
// "svd3, clamp, U S V^T" for every snow particle -- ~3x fewer operations; the two agree to ~1e-6 (tests/test_host_math.py).
inline float plastic_project3(float lo, float hi, M3 &F) {
  M3 R;
  if (!polar_newton3(F, R)) {
    M3 U, V;
    float sg[3];
    svd3(F, U, sg, V);
    M3 sig = m3_zero();
    for (int i = 0; i < 3; i++) sig.d[i][i] = clampf(sg[i], lo, hi);
    const float oldJ = m3_det(F);
    F = m3_mul(m3_mul(U, sig), m3_transposed(V));
    return oldJ / m3_det(F);
  }
  // S = R^T F: S(i,j) = <column i of R, column j of F>; symmetric part
  float a00 = dot3f(R.d[0], F.d[0]), a11 = dot3f(R.d[1], F.d[1]), a22 = dot3f(R.d[2], F.d[2]);
  float a01 = 0.5f * (dot3f(R.d[0], F.d[1]) + dot3f(R.d[1], F.d[0]));
  float a02 = 0.5f * (dot3f(R.d[0], F.d[2]) + dot3f(R.d[2], F.d[0]));
  float a12 = 0.5f * (dot3f(R.d[1], F.d[2]) + dot3f(R.d[2], F.d[1]));
  {
    const float r0 = std::fabs(a01) + std::fabs(a02), r1 = std::fabs(a01) + std::fabs(a12), r2 = std::fabs(a02) + std::fabs(a12);
    if (a00 - r0 >= lo && a00 + r0 <= hi && a11 - r1 >= lo && a11 + r1 <= hi && a22 - r2 >= lo && a22 + r2 <= hi) return 1.0f;
  }
  M3 V = m3_diag(1.0f);
  // one Jacobi rotation in the (p, q) plane; r = the third index; apq, arp, arq = the off-diagonal entries
#define ORACLE_JACOBI(app, aqq, apq, arp, arq, p, q)                                   \
  if (std::fabs(apq) > 6e-8f * (std::fabs(app) + std::fabs(aqq))) {                     \
    rotated = true;                                                                    \
    const float h_ = 0.5f * (aqq - app);                                               \
    const float t_ = apq / (h_ + std::copysign(std::sqrt(std::fma(h_, h_, apq * apq)), h_)); \
    const float c_ = 1.0f / std::sqrt(std::fma(t_, t_, 1.0f)), s_ = t_ * c_;           \
    app = std::fma(-t_, apq, app);                                                     \
    aqq = std::fma(t_, apq, aqq);                                                      \
    apq = 0.0f;                                                                        \
    const float x_ = arp, y_ = arq;                                                    \
    arp = std::fma(c_, x_, -(s_ * y_));                                                \
    arq = std::fma(s_, x_, c_ * y_);                                                   \
    for (int k = 0; k < 3; k++) {                                                      \
      const float vp_ = V.d[p][k], vq_ = V.d[q][k];                                    \
      V.d[p][k] = std::fma(c_, vp_, -(s_ * vq_));                                      \
      V.d[q][k] = std::fma(s_, vp_, c_ * vq_);                                         \
    }                                                                                  \
  }
  for (int sweep = 0; sweep < 6; sweep++) {
    bool rotated = false;
    ORACLE_JACOBI(a00, a11, a01, a02, a12, 0, 1)
    ORACLE_JACOBI(a00, a22, a02, a01, a12, 0, 2)
    ORACLE_JACOBI(a11, a22, a12, a01, a02, 1, 2)
    if (!rotated) break;
  }
#undef ORACLE_JACOBI
  const float l0 = clampf(a00, lo, hi), l1 = clampf(a11, lo, hi), l2 = clampf(a22, lo, hi);
  // S' = V diag(l') V^T (symmetric), F' = R S'
  float w0[3], w1[3], w2[3];
  for (int k = 0; k < 3; k++) {
    w0[k] = l0 * V.d[0][k];
    w1[k] = l1 * V.d[1][k];
    w2[k] = l2 * V.d[2][k];
  }
  M3 S;
  for (int i = 0; i < 3; i++)
    for (int j = i; j < 3; j++)
      S.d[j][i] = S.d[i][j] = std::fma(w2[i], V.d[2][j], std::fma(w1[i], V.d[1][j], w0[i] * V.d[0][j]));
  for (int j = 0; j < 3; j++)
    for (int k = 0; k < 3; k++)
      F.d[j][k] = std::fma(R.d[2][k], S.d[j][2], std::fma(R.d[1][k], S.d[j][1], R.d[0][k] * S.d[j][0]));
  return (a00 * a11 * a22) / (l0 * l1 * l2);
}

void advance3(const OracleParams &P, float dt, P3 *particles, long long n, float *grid /*(n+1)^3*4*/,
              float *grid_post_p2g, float *vold /*(n+1)^3*3*/, int n_threads = 1) {
  const int num_grid = P.n_grid;
  const int N1 = num_grid + 1;
  const float dx = 1.0f / num_grid;
  const float inv_dx = 1.0f / dx;
  const float mass_p = P.mass_p, vol_p = P.vol_p;
  float mu_0[4], lambda_0[4];
  for (int m = 0; m < P.n_materials; m++) lame(P.mat[m].E, P.mat[m].nu, mu_0[m], lambda_0[m]);
  const bool flip = P.alpha != 0.0f;
  const size_t NN = (size_t)N1 * N1 * N1;
  const int T = n_threads > 1 ? n_threads : 1;
  std::vector<int> bx, cut;
  if (T > 1) plan_columns(particles, n, inv_dx, N1, T, bx, cut);
  else cut = {0, N1};

  run_threads(T, [&](int t) {
  const int col_lo = cut[(size_t)t], col_hi = cut[(size_t)t + 1];  // node x-planes this thread accumulates
  std::memset(grid + 4 * (size_t)col_lo * N1 * N1, 0, sizeof(float) * 4 * (size_t)(col_hi - col_lo) * N1 * N1);

  for (long long pi = 0; pi < n; pi++) {
    if (T > 1) {
      const int b = bx[(size_t)pi];
      if (b + 2 < col_lo || b >= col_hi) continue;
    }
    P3 &p = particles[pi];
    const int mid = material_of(P, p.c);
    const OracleMaterial &mat = P.mat[mid];
    int base[3];
    float fx[3];
    for (int k = 0; k < 3; k++) {
      base[k] = (int)(p.x[k] * inv_dx - 0.5f);
      fx[k] = p.x[k] * inv_dx - (float)base[k];
    }
    float w[3][3];
    for (int k = 0; k < 3; k++) {
      w[0][k] = 0.5f * ((1.5f - fx[k]) * (1.5f - fx[k]));
      w[1][k] = 0.75f - ((fx[k] - 1.0f) * (fx[k] - 1.0f));
      w[2][k] = 0.5f * ((fx[k] - 0.5f) * (fx[k] - 0.5f));
    }
    float e;
    if (mat.kind == KIND_SNOW) e = std::exp(mat.hardening * (1.0f - p.Jp));
    else if (mat.kind == KIND_JELLY) e = mat.hardening;
    else e = 1.0f;
    float mu = mu_0[mid] * e;
    float lambda = lambda_0[mid] * e;
    float J = m3_det(p.F);
    float Dinv = 4 * inv_dx * inv_dx;
    M3 PF;
    if (mat.kind == KIND_FLUID) {
      PF = m3_diag(lambda * (J - 1) * J);
    } else {
      M3 r = rotation3(p.F);
      PF = m3_add(m3_mul(m3_scale(2 * mu, m3_sub(p.F, r)), m3_transposed(p.F)), m3_diag(lambda * (J - 1) * J));
    }
    M3 stress = m3_scale(-(dt * vol_p), m3_scale(Dinv, PF));
    M3 affine = m3_add(stress, m3_scale(mass_p, p.C));
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++)
        for (int k = 0; k < 3; k++) {
          float dpos[3] = {((float)i - fx[0]) * dx, ((float)j - fx[1]) * dx, ((float)k - fx[2]) * dx};
          float ad[3];
          m3_mulvec(affine, dpos, ad);
          float wgt = w[i][0] * w[j][1] * w[k][2];
          if (T > 1 && (base[0] + i < col_lo || base[0] + i >= col_hi)) continue;  // another thread's plane
          float *g = grid + 4 * (((size_t)(base[0] + i) * N1 + (base[1] + j)) * N1 + (base[2] + k));
          g[0] = g[0] + wgt * (mass_p * p.v[0] + ad[0]);
          g[1] = g[1] + wgt * (mass_p * p.v[1] + ad[1]);
          g[2] = g[2] + wgt * (mass_p * p.v[2] + ad[2]);
          g[3] = g[3] + wgt * (mass_p + 0.0f);
        }
  }
  });
  if (grid_post_p2g) std::memcpy(grid_post_p2g, grid, sizeof(float) * 4 * NN);

  run_threads(T, [&](int t) {
  for (int i = (int)((long long)N1 * t / T); i < (int)((long long)N1 * (t + 1) / T); i++)
    for (int j = 0; j <= num_grid; j++)
      for (int k = 0; k <= num_grid; k++) {
        size_t node = ((size_t)i * N1 + j) * N1 + k;
        float *g = grid + 4 * node;
        if (flip) vold[3 * node] = vold[3 * node + 1] = vold[3 * node + 2] = 0.0f;
        if (g[3] > 0) {
          float m = g[3];
          g[0] = g[0] / m;
          g[1] = g[1] / m;
          g[2] = g[2] / m;
          g[3] = g[3] / m;
          if (flip) {
            vold[3 * node + 0] = g[0];
            vold[3 * node + 1] = g[1];
            vold[3 * node + 2] = g[2];
          }
          g[0] = g[0] + dt * P.gravity[0];
          g[1] = g[1] + dt * P.gravity[1];
          g[2] = g[2] + dt * P.gravity[2];
          g[3] = g[3] + dt * 0.0f;
          float boundary = P.boundary;
          float x = (float)i / num_grid, y = (float)j / num_grid, z = (float)k / num_grid;
          if (x < boundary || x > 1 - boundary || y > 1 - boundary || z < boundary || z > 1 - boundary) {
            g[0] = g[1] = g[2] = g[3] = 0.0f;
          }
          if (y < boundary) g[1] = std::max(0.0f, g[1]);
        }
      }
  });

  run_threads(T, [&](int t) {
  for (long long pi = n * t / T; pi < n * (t + 1) / T; pi++) {
    P3 &p = particles[pi];
    const int mid = material_of(P, p.c);
    const OracleMaterial &mat = P.mat[mid];
    int base[3];
    float fx[3];
    for (int k = 0; k < 3; k++) {
      base[k] = (int)(p.x[k] * inv_dx - 0.5f);
      fx[k] = p.x[k] * inv_dx - (float)base[k];
    }
    float w[3][3];
    for (int k = 0; k < 3; k++) {
      w[0][k] = 0.5f * ((1.5f - fx[k]) * (1.5f - fx[k]));
      w[1][k] = 0.75f - ((fx[k] - 1.0f) * (fx[k] - 1.0f));
      w[2][k] = 0.5f * ((fx[k] - 0.5f) * (fx[k] - 0.5f));
    }
    float v_in[3] = {p.v[0], p.v[1], p.v[2]};
    p.C = m3_zero();
    p.v[0] = p.v[1] = p.v[2] = 0.0f;
    float dv[3] = {0.0f, 0.0f, 0.0f};
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++)
        for (int k = 0; k < 3; k++) {
          float dpos[3] = {(float)i - fx[0], (float)j - fx[1], (float)k - fx[2]};
          size_t node = ((size_t)(base[0] + i) * N1 + (base[1] + j)) * N1 + (base[2] + k);
          float gv[3] = {grid[4 * node + 0], grid[4 * node + 1], grid[4 * node + 2]};
          float weight = w[i][0] * w[j][1] * w[k][2];
          float wg[3] = {weight * gv[0], weight * gv[1], weight * gv[2]};
          for (int r = 0; r < 3; r++) p.v[r] = p.v[r] + wg[r];
          float s4 = 4 * inv_dx;
          for (int c = 0; c < 3; c++)
            for (int r = 0; r < 3; r++) p.C.d[c][r] = p.C.d[c][r] + s4 * (wg[r] * dpos[c]);
          if (flip)
            for (int r = 0; r < 3; r++) dv[r] = dv[r] + weight * (gv[r] - vold[3 * node + r]);
        }
    for (int r = 0; r < 3; r++) p.x[r] = p.x[r] + dt * p.v[r];
    if (flip) {
      float a = P.alpha;
      for (int r = 0; r < 3; r++) p.v[r] = (1.0f - a) * p.v[r] + a * (v_in[r] + dv[r]);
    }
    M3 F = m3_mul(m3_add(m3_diag(1.0f), m3_scale(dt, p.C)), p.F);
    if (mat.kind == KIND_SNOW) {
      const float ratio = plastic_project3(mat.sig_lo, mat.sig_hi, F);  // det(F) / det(F')
      p.Jp = clampf(p.Jp * ratio, P.jp_min, P.jp_max);
      p.F = F;
    } else if (mat.kind == KIND_JELLY) {
      p.F = F;
    } else {
      float J = m3_det(F);
      p.F = m3_diag(std::cbrt(J));
    }
  }
  });
}

// xorshift128, taichi.h:6497-6503 (state is process-wide there; explicit here)
struct XorShift {
  uint32_t x = 123456789, y = 362436069, z = 521288629, w = 88675123;
  uint32_t next() {
    uint32_t t = x ^ (x << 11);
    x = y;
    y = z;
    z = w;
    return (w = (w ^ (w >> 19)) ^ (t ^ (t >> 8)));
  }
  float rand() { return next() * (1.0f / 4294967296.0f); }  // taichi.h:6511-6513
};

}  // namespace

extern "C" {

int oracle_params_bytes() { return (int)sizeof(OracleParams); }

// n_steps substeps of the 2D/3D restatement on caller-owned AoS records (56 B / 104 B).
// grid_out: final grid of the last substep ((n+1)^d * (d+1) floats), may be NULL.
// grid_post_p2g: grid of the LAST substep tapped between P2G and the grid update, may be NULL.
// n_threads > 1: the same loops on several host threads, bitwise identical to n_threads == 1 (see run_threads).
int oracle_advance_mt(const void *params, float dt, void *particles, long long n, int n_steps, float *grid_out,
                      float *grid_post_p2g, int n_threads) {
  const OracleParams &P = *(const OracleParams *)params;
  if (P.dim != 2 && P.dim != 3) return -1;
  if (P.n_materials < 1 || P.n_materials > 4) return -1;
  const int N1 = P.n_grid + 1;
  size_t nodes = P.dim == 2 ? (size_t)N1 * N1 : (size_t)N1 * N1 * N1;
  std::vector<float> grid_local;
  float *grid = grid_out;
  if (!grid) {
    grid_local.resize(nodes * (P.dim + 1));
    grid = grid_local.data();
  }
  std::vector<float> vold;
  if (P.alpha != 0.0f) vold.resize(nodes * P.dim);
  for (int s = 0; s < n_steps; s++) {
    float *tap = (s == n_steps - 1) ? grid_post_p2g : nullptr;
    if (P.dim == 2) advance2(P, dt, (P2 *)particles, n, grid, tap, vold.data(), n_threads);
    else advance3(P, dt, (P3 *)particles, n, grid, tap, vold.data(), n_threads);
  }
  return 0;
}
int oracle_advance(const void *params, float dt, void *particles, long long n, int n_steps, float *grid_out,
                   float *grid_post_p2g) {
  return oracle_advance_mt(params, dt, particles, n, n_steps, grid_out, grid_post_p2g, 1);
}

void oracle_lame(float E, float nu, float *mu0, float *lambda0) { lame(E, nu, *mu0, *lambda0); }

void oracle_polar2(const float *m, float *R, float *S) {
  M2 M, r = m2_zero(), s = m2_zero();
  std::memcpy(&M, m, 16);
  polar2(M, r, s);
  std::memcpy(R, &r, 16);
  std::memcpy(S, &s, 16);
}
void oracle_svd2(const float *m, float *U, float *sig, float *V) {
  M2 M, u = m2_zero(), sg = m2_zero(), v = m2_zero();
  std::memcpy(&M, m, 16);
  svd2(M, u, sg, v);
  std::memcpy(U, &u, 16);
  std::memcpy(sig, &sg, 16);
  std::memcpy(V, &v, 16);
}
void oracle_rotation3(const float *m, float *R) {
  M3 M, r;
  std::memcpy(&M, m, 36);
  r = rotation3(M);
  std::memcpy(R, &r, 36);
}
float oracle_plastic_project3(float lo, float hi, float *m) {
  M3 M;
  std::memcpy(&M, m, 36);
  const float r = plastic_project3(lo, hi, M);
  std::memcpy(m, &M, 36);
  return r;
}
void oracle_svd3(const float *m, float *U, float *sig3, float *V) {
  M3 M, u = m3_zero(), v = m3_zero();
  std::memcpy(&M, m, 36);
  svd3(M, u, sig3, v);
  std::memcpy(U, &u, 36);
  std::memcpy(V, &v, 36);
}

// The shipped seeding (mls-mpm88-explained.cpp:191-196 as called from :207) into 56-byte records.
// `skip` RNG draws are discarded first so several blocks can continue one stream.
void oracle_seed_block2(void *aos56, int n, float cx, float cy, float half, int c, unsigned long long skip) {
  XorShift rng;
  for (unsigned long long i = 0; i < skip; i++) rng.next();
  P2 *p = (P2 *)aos56;
  for (int i = 0; i < n; i++) {
    float rx = rng.rand();
    float ry = rng.rand();
    std::memset(&p[i], 0, sizeof(P2));
    p[i].x[0] = (rx * 2.0f - 1.0f) * half + cx;
    p[i].x[1] = (ry * 2.0f - 1.0f) * half + cy;
    p[i].F = m2_diag(1.0f);
    p[i].Jp = 1.0f;
    p[i].c = c;
  }
}

// CPU binning oracle: cell (base) coordinate per particle exactly as :55, bin key = block of
// `bin_edge` cells (x-major linear id), and the STABLE permutation that sorts `order_in`
// (previous slot->particle map, or NULL for identity) by that key.
//   x           : n*dim floats (AoS positions)
//   cell_out    : n*dim ints   (clamped to [0, n_grid-2])
//   key_out     : n ints
//   order_out   : n ints, slot -> particle index after the stable sort
//   bin_start   : n_bins+1 ints
int oracle_bin(int dim, int n_grid, int bin_edge, const float *x, long long n, const int *order_in, int *cell_out,
               int *key_out, int *order_out, int *bin_start) {
  const float dx = 1.0f / n_grid;
  const float inv_dx = 1.0f / dx;
  const int nb = (n_grid - 1 + bin_edge - 1) / bin_edge;  // bases 0..n_grid-2
  long long n_bins = 1;
  for (int k = 0; k < dim; k++) n_bins *= nb;
  std::vector<int> count(n_bins + 1, 0);
  for (long long i = 0; i < n; i++) {
    int key = 0;
    for (int k = 0; k < dim; k++) {
      int b = (int)(x[i * dim + k] * inv_dx - 0.5f);
      if (b < 0) b = 0;
      if (b > n_grid - 2) b = n_grid - 2;
      cell_out[i * dim + k] = b;
      key = key * nb + b / bin_edge;
    }
    key_out[i] = key;
    count[key + 1]++;
  }
  for (long long b = 0; b < n_bins; b++) count[b + 1] += count[b];
  for (long long b = 0; b <= n_bins; b++) bin_start[b] = count[b];
  std::vector<int> cursor(count.begin(), count.end() - 1);
  for (long long s = 0; s < n; s++) {
    int pidx = order_in ? order_in[s] : (int)s;
    order_out[cursor[key_out[pidx]]++] = pidx;
  }
  return (int)n_bins;
}
}
