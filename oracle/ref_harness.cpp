// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Drives the UNMODIFIED reference translation unit
//   /root/reference/cpp_validation/mls-mpm88-explained.cpp
// by #including it where it lies (the path comes from -I on the command line,
// nothing is copied into this repository) with its main() renamed, and
// exposes its globals `particles` (:44), `grid` (:47), `add_object` (:191) and
// `advance` (:49) through a small C interface that tests/ and bench.py load
// with ctypes.  Output goes to oracle/_ref/libmpmref.so (git-ignored).
//
// The reference's constants (num_grid=80, dt=1e-4, E=1e2, nu=0.499, ...) are
// compile-time consts, so this library covers ONLY the scene as shipped
// (BASELINE config 1); everything else goes through oracle/mpm_oracle.cpp,
// which is gated on being bitwise identical to this library on that scene.
#include <cstring>

#define main mpm_reference_main
#include "mls-mpm88-explained.cpp"
#undef main

static_assert(sizeof(Particle) == 56, "reference Particle is expected to be 56 bytes");

extern "C" {

int ref_num_grid() { return num_grid; }
float ref_dt() { return dt; }
int ref_particle_bytes() { return (int)sizeof(Particle); }
float ref_mu0() { return mu_0; }
float ref_lambda0() { return lambda_0; }

void ref_clear() { particles.clear(); }

// The shipped seeding: add_object(Vec(0.5,0.5)) as main() calls it (:207).
// NOTE: the xorshift state in taichi.h:6497 is a function-local static, so the
// sequence continues across calls; the first call in a fresh process gives the
// scene the reference program runs.
void ref_seed_shipped() { add_object(Vec(0.5, 0.5)); }

int ref_count() { return (int)particles.size(); }

// Replace the particle set with caller-provided 56-byte records.
void ref_set(const void *aos56, int n) {
  particles.clear();
  particles.reserve(n);
  const Particle *src = (const Particle *)aos56;
  for (int i = 0; i < n; i++) particles.push_back(src[i]);
}

void ref_get(void *aos56_out) { std::memcpy(aos56_out, particles.data(), particles.size() * sizeof(Particle)); }

// grid after the LAST advance(): (vx, vy, 1|0) per node, [i][j] row-major, 12 B/node.
void ref_get_grid(float *out) { std::memcpy(out, grid, sizeof(grid)); }

void ref_advance(int n_steps) {
  for (int s = 0; s < n_steps; s++) advance(dt);
}

void ref_advance_dt(float step_dt, int n_steps) {
  for (int s = 0; s < n_steps; s++) advance(step_dt);
}

// 2x2 decompositions of taichi.h:8375-8420, column-major 4-float matrices in/out.
void ref_polar2(const float *m, float *R, float *S) {
  Mat M, r, s;
  std::memcpy(&M, m, 16);
  polar_decomp(M, r, s);
  std::memcpy(R, &r, 16);
  std::memcpy(S, &s, 16);
}
void ref_svd2(const float *m, float *U, float *sig, float *V) {
  Mat M, u, sg, v;
  std::memcpy(&M, m, 16);
  svd(M, u, sg, v);
  std::memcpy(U, &u, 16);
  std::memcpy(sig, &sg, 16);
  std::memcpy(V, &v, 16);
}

// First values of the reference RNG (taichi.h:6497-6513), for pinning the restated xorshift.
float ref_rand() { return taichi::rand(); }
}
