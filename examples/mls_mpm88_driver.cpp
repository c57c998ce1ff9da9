// examples/mls_mpm88_driver.cpp -- the reference program's main() loop on the C-ABI.
//
// What main() of /root/reference/cpp_validation/mls-mpm88-explained.cpp:203-227 does -- seed the scene
// (add_object, :191-196), 2500 x advance(dt) (:214-215), draw every int(frame_dt/dt) = 10 substeps
// (:217-225) -- with advance() replaced by mpm_substep() on a B200 and the GUI / PNG writer replaced by a
// headless point splat into binary PPM frames (no X11, no stb).  Plain C++14 + include/mpm.h.
//
//   mls_mpm88_driver [--steps N] [--frames DIR] [--dump FILE] [--three-blocks] [--devices 0,1,...]
//
// --dump writes the final 56-byte particle records (the reference's own struct layout) for the tests.
// --devices runs the same loop on several GPUs (one x-slab per listed CUDA ordinal, a device may repeat) through
// the mpm_group_* calls: same three statements -- upload, substep, read.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "mpm.h"

// The reference's particle record, :28-42 (Vec x, v; Mat F; Mat C; real Jp; int c) -- 56 bytes.
struct Particle {
  float x[2], v[2], F[4], C[4], Jp;
  int c;
};
static_assert(sizeof(Particle) == 56, "wire format of include/mpm.h");

// taichi.h:6497-6513: xorshift128 behind Vec::rand(); x then y per particle (taichi.h:7317-7323)
struct Rand {
  uint32_t x = 123456789, y = 362436069, z = 521288629, w = 88675123;
  float next() {
    uint32_t t = x ^ (x << 11);
    x = y;
    y = z;
    z = w;
    w = (w ^ (w >> 19)) ^ (t ^ (t >> 8));
    return w * (1.0f / 4294967296.0f);
  }
};

static void add_block(std::vector<Particle> &ps, Rand &rng, int n, float cx, float cy, float half, int c) {
  for (int i = 0; i < n; i++) {  // :193-195: (Vec::rand()*2.0f - Vec(1))*half + centre
    Particle p;
    memset(&p, 0, sizeof p);
    float rx = rng.next(), ry = rng.next();
    p.x[0] = (rx * 2.0f - 1.0f) * half + cx;
    p.x[1] = (ry * 2.0f - 1.0f) * half + cy;
    p.F[0] = p.F[3] = 1.0f;  // F(1), C(0), Jp(1): the constructor, :35-41
    p.Jp = 1.0f;
    p.c = c;
    ps.push_back(p);
  }
}

static void write_frame(const std::string &dir, int frame, const std::vector<Particle> &ps, int size) {
  std::vector<unsigned char> img((size_t)size * size * 3);
  for (size_t k = 0; k < img.size(); k += 3) {  // canvas.clear(0x112F41), :218
    img[k] = 0x11;
    img[k + 1] = 0x2F;
    img[k + 2] = 0x41;
  }
  for (const Particle &p : ps) {  // canvas.circle(p.x).radius(2).color(p.c), :220-222 (as a 3x3 splat)
    int px = (int)(p.x[0] * size), py = size - 1 - (int)(p.x[1] * size);
    unsigned c = p.c > 3 ? (unsigned)p.c : (p.c == 0 ? 0x068587u : p.c == 1 ? 0xED553Bu : 0xEEEEF0u);
    for (int dy = -1; dy <= 1; dy++)
      for (int dx = -1; dx <= 1; dx++) {
        int xx = px + dx, yy = py + dy;
        if (xx < 0 || yy < 0 || xx >= size || yy >= size) continue;
        unsigned char *q = &img[((size_t)yy * size + xx) * 3];
        q[0] = (c >> 16) & 255;
        q[1] = (c >> 8) & 255;
        q[2] = c & 255;
      }
  }
  char name[512];
  snprintf(name, sizeof name, "%s/%05d.ppm", dir.c_str(), frame);  // the reference writes tmp/%05d.png, :224
  if (FILE *f = fopen(name, "wb")) {
    fprintf(f, "P6\n%d %d\n255\n", size, size);
    fwrite(img.data(), 1, img.size(), f);
    fclose(f);
  }
}

int main(int argc, char **argv) {
  int steps = 2500;  // :214
  std::string frames, dump;
  bool three = false;
  std::vector<int> devices;
  for (int a = 1; a < argc; a++) {
    if (!strcmp(argv[a], "--devices") && a + 1 < argc) {
      for (char *t = strtok(argv[++a], ","); t; t = strtok(nullptr, ",")) devices.push_back(atoi(t));
      continue;
    }
    if (!strcmp(argv[a], "--steps") && a + 1 < argc) steps = atoi(argv[++a]);
    else if (!strcmp(argv[a], "--frames") && a + 1 < argc) frames = argv[++a];
    else if (!strcmp(argv[a], "--dump") && a + 1 < argc) dump = argv[++a];
    else if (!strcmp(argv[a], "--three-blocks")) three = true;
  }
  const float dt = 1e-4f, frame_dt = 1e-3f;  // :11-12
  std::vector<Particle> particles;
  Rand rng;
  if (three) {  // the seeding left commented in the reference, :182-188, :207-209
    add_block(particles, rng, 1000, 0.55f, 0.45f, 0.08f, MPM_KIND_FLUID);
    add_block(particles, rng, 1000, 0.45f, 0.65f, 0.08f, MPM_KIND_JELLY);
    add_block(particles, rng, 1000, 0.55f, 0.85f, 0.08f, MPM_KIND_SNOW);
  } else {  // as shipped, :191-196: 3000 particles, centre (0.05+0.08, 0.05+0.08), colour 0x2986CC
    add_block(particles, rng, 3000, 0.05f + 0.08f, 0.05f + 0.08f, 0.08f, 0x2986CC);
  }

  mpm_config cfg;
  mpm_default_config(&cfg, 2);  // the constants of :8-26
  cfg.capacity = (long long)particles.size();
  if (!devices.empty()) {  // several GPUs behind one handle: the same loop
    mpm_group *g = mpm_group_create(&cfg, devices.data(), (int)devices.size());
    if (!g || mpm_group_upload_particles(g, particles.data(), (long long)particles.size()) != MPM_OK) {
      fprintf(stderr, "group: %s\n", mpm_group_last_error(g));
      return 1;
    }
    int frame = 0;
    const int every = (int)(frame_dt / dt);  // :217
    for (int step = 0; step < steps; step += every) {
      const int k = steps - step < every ? steps - step : every;
      if (mpm_group_substep(g, dt, k) != MPM_OK) {  // `every` x advance(dt), :215
        fprintf(stderr, "substep: %s\n", mpm_group_last_error(g));
        return 1;
      }
      if (!frames.empty()) {
        mpm_group_read_particles(g, particles.data(), (long long)particles.size());
        write_frame(frames, frame++, particles, 800);
      }
    }
    if (mpm_group_read_particles(g, particles.data(), (long long)particles.size()) != MPM_OK) {
      fprintf(stderr, "read: %s\n", mpm_group_last_error(g));
      return 1;
    }
    const int st = mpm_group_poll_status(g);
    double cx = 0, cy = 0, ke = 0;
    for (const Particle &p : particles) {
      cx += p.x[0];
      cy += p.x[1];
      ke += 0.5 * ((double)p.v[0] * p.v[0] + (double)p.v[1] * p.v[1]);
    }
    printf("steps %d particles %zu status %d com %.7f %.7f ke %.4f slabs %zu:", steps, particles.size(), st,
           cx / particles.size(), cy / particles.size(), ke, devices.size());
    for (int k = 0; k < (int)devices.size(); k++) {
      int dev, lo, hi;
      long long cnt;
      mpm_group_slab(g, k, &dev, &lo, &hi, &cnt);
      printf(" [gpu %d: columns %d..%d, %lld particles]", dev, lo, hi, cnt);
    }
    printf("\n");
    if (!dump.empty())
      if (FILE *f = fopen(dump.c_str(), "wb")) {
        fwrite(particles.data(), sizeof(Particle), particles.size(), f);
        fclose(f);
      }
    mpm_group_destroy(g);
    return st == MPM_OK ? 0 : 2;
  }
  mpm_handle *h = mpm_create(&cfg);
  if (!h) {
    fprintf(stderr, "mpm_create: %s\n", mpm_last_error(nullptr));
    return 1;
  }
  if (mpm_upload_particles(h, particles.data(), (long long)particles.size(), 0) != MPM_OK) {
    fprintf(stderr, "upload: %s\n", mpm_last_error(h));
    return 1;
  }
  int frame = 0;
  const int every = (int)(frame_dt / dt);  // :217
  for (int step = 0; step < steps; step++) {
    if (mpm_substep(h, dt, 1) != MPM_OK) {  // advance(dt), :215
      fprintf(stderr, "substep: %s\n", mpm_last_error(h));
      return 1;
    }
    if (!frames.empty() && step % every == 0) {
      mpm_read_particles(h, particles.data(), (long long)particles.size(), 0);
      write_frame(frames, frame++, particles, 800);  // window_size, :8
    }
  }
  if (mpm_read_particles(h, particles.data(), (long long)particles.size(), 0) != MPM_OK) {
    fprintf(stderr, "read: %s\n", mpm_last_error(h));
    return 1;
  }
  int st = mpm_poll_status(h);
  double cx = 0, cy = 0, mx = 0, my = 0, ke = 0;
  for (const Particle &p : particles) {
    cx += p.x[0];
    cy += p.x[1];
    mx += p.v[0];
    my += p.v[1];
    ke += 0.5 * ((double)p.v[0] * p.v[0] + (double)p.v[1] * p.v[1]);
  }
  printf("steps %d particles %zu status %d com %.7f %.7f mom %.4f %.4f ke %.4f frames %d\n", steps, particles.size(), st,
         cx / particles.size(), cy / particles.size(), mx, my, ke, frame);
  if (!dump.empty())
    if (FILE *f = fopen(dump.c_str(), "wb")) {
      fwrite(particles.data(), sizeof(Particle), particles.size(), f);
      fclose(f);
    }
  mpm_destroy(h);
  return st == MPM_OK ? 0 : 2;
}
