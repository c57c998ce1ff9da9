// mpm_group.cu -- several GPUs behind ONE handle of the C-ABI (include/mpm.h, mpm_group_*): the reference's
// main() loop (cpp_validation/mls-mpm88-explained.cpp:203-227) can drive an x-slab decomposition from plain C++
// with the same three calls it would use for one GPU -- upload, substep, read.
//
// One slab handle per device (mpm_create with slab_lo/slab_hi), all driven from the calling thread.  The exchange of
// the x-slab protocol (one fixed-size message per neighbour and substep, see mpm_engine.cu) is a PULL over peer
// memory: each slab's stream waits for its neighbours' "staged" events, copies their messages into its own landing
// zones with cudaMemcpyPeerAsync (NVLink when peer access is available) and runs its substep; a "pulled" event
// keeps a neighbour from overwriting a message that is still being read.  No host synchronisation per substep,
// no collective: the pattern is nearest-neighbour only.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/mpm.h"

namespace {

struct Slab {
  mpm_handle *h = nullptr;
  int device = 0, lo = 0, hi = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t staged = nullptr, pulled = nullptr;
  mpm_slab_desc d;
  std::vector<float> host;  // upload / read-back staging (pinned would be faster; this is the convenience path)
  std::vector<int> ids;
};

}  // namespace

struct mpm_group {
  mpm_config cfg;
  std::vector<Slab> slabs;
  std::string err;
  long long n_total = 0;
  bool staged = false;  // messages of the last step / begin are waiting to be pulled
  float begun_dt = 0.0f;
  int words() const { return cfg.dim == 2 ? 14 : 26; }
  ~mpm_group() {
    for (Slab &s : slabs) {
      cudaSetDevice(s.device);
      if (s.h) mpm_destroy(s.h);
      if (s.staged) cudaEventDestroy(s.staged);
      if (s.pulled) cudaEventDestroy(s.pulled);
      if (s.stream) cudaStreamDestroy(s.stream);
    }
  }
  int fail(int rc, const std::string &what, const Slab *s = nullptr) {
    err = what;
    if (s && s->h) err += std::string(": ") + mpm_last_error(s->h);
    return rc;
  }
  int cuda(cudaError_t e, const char *what) {
    if (e == cudaSuccess) return MPM_OK;
    err = std::string(what) + ": " + cudaGetErrorString(e);
    return MPM_E_CUDA;
  }
  // pull the neighbours' staged messages into every slab's landing zones (stream-ordered, no host wait)
  int exchange() {
    const int N = (int)slabs.size();
    for (int i = 0; i < N; i++) {
      Slab &s = slabs[i];
      cudaSetDevice(s.device);
      if (i > 0) {
        Slab &lo = slabs[i - 1];
        cudaStreamWaitEvent(s.stream, lo.staged, 0);
        int rc = cuda(cudaMemcpyPeerAsync(s.d.recv_lo, s.device, lo.d.send_hi, lo.device, (size_t)s.d.bytes, s.stream),
                      "peer copy");
        if (rc) return rc;
      }
      if (i + 1 < N) {
        Slab &hi = slabs[i + 1];
        cudaStreamWaitEvent(s.stream, hi.staged, 0);
        int rc = cuda(cudaMemcpyPeerAsync(s.d.recv_hi, s.device, hi.d.send_lo, hi.device, (size_t)s.d.bytes, s.stream),
                      "peer copy");
        if (rc) return rc;
      }
      cudaEventRecord(s.pulled, s.stream);
    }
    // nobody restages a message before its readers have pulled it
    for (int i = 0; i < N; i++) {
      cudaSetDevice(slabs[i].device);
      if (i > 0) cudaStreamWaitEvent(slabs[i].stream, slabs[i - 1].pulled, 0);
      if (i + 1 < N) cudaStreamWaitEvent(slabs[i].stream, slabs[i + 1].pulled, 0);
    }
    staged = false;
    return MPM_OK;
  }
  int mark_staged() {
    for (Slab &s : slabs) {
      cudaSetDevice(s.device);
      cudaEventRecord(s.staged, s.stream);
    }
    staged = true;
    return MPM_OK;
  }
};

extern "C" {

mpm_group *mpm_group_create(const mpm_config *cfg, const int *devices, int n_devices) {
  if (!cfg || !devices || n_devices < 1) return nullptr;
  mpm_group *g = new (std::nothrow) mpm_group();
  if (!g) return nullptr;
  g->cfg = *cfg;
  g->slabs.resize((size_t)n_devices);
  for (int i = 0; i < n_devices; i++) g->slabs[(size_t)i].device = devices[i];
  return g;  // the slab handles are created by the upload, which knows where the particles are
}

void mpm_group_destroy(mpm_group *g) { delete g; }

const char *mpm_group_last_error(const mpm_group *g) { return g ? g->err.c_str() : "mpm_group: NULL handle"; }

// replaces add_object's push_back loop (:191-196) for N devices: cuts the grid into x-slabs of (nearly) equal
// particle counts, creates one slab handle per device and uploads each slab's particles; ids = upload indices
int mpm_group_upload_particles(mpm_group *g, const void *aos, long long n) {
  if (!g || n < 0 || (n > 0 && !aos)) return MPM_E_INVALID;
  const int N = (int)g->slabs.size(), W = g->words(), ng = g->cfg.n_grid;
  const int edge = g->cfg.bin_edge > 0 ? g->cfg.bin_edge : (g->cfg.dim == 2 ? 8 : 4);
  if (ng / edge < N) return g->fail(MPM_E_INVALID, "mpm_group: grid too small for that many slabs");
  const float *p = (const float *)aos;
  // base column exactly as the engine computes it (:55): fp32 multiply, subtract, truncate, clamp
  const float dx = 1.0f / ng, inv_dx = 1.0f / dx;
  std::vector<int> col((size_t)n);
  std::vector<long long> hist((size_t)ng + 1, 0);
  for (long long i = 0; i < n; i++) {
    volatile float t = p[i * W] * inv_dx;
    int b = (int)(t - 0.5f);
    b = b < 0 ? 0 : (b > ng - 2 ? ng - 2 : b);
    col[(size_t)i] = b;
    hist[(size_t)b]++;
  }
  // cuts at bin-edge multiples, balancing the particle counts
  std::vector<int> cut((size_t)N + 1, ng);
  cut[0] = 0;
  {
    long long run = 0;
    int k = 1;
    for (int c = 0; c < ng && k < N; c++) {
      run += hist[(size_t)c];
      if ((c + 1) % edge == 0 && run >= n * k / N) cut[(size_t)k++] = c + 1;
    }
    for (int k2 = 1; k2 < N; k2++) {  // every slab at least one bin wide, room left for the slabs above
      cut[(size_t)k2] = std::max(cut[(size_t)k2], cut[(size_t)k2 - 1] + edge);
      cut[(size_t)k2] = std::min(cut[(size_t)k2], (ng / edge - (N - k2)) * edge);
    }
  }
  std::vector<long long> count((size_t)N, 0);
  std::vector<int> owner((size_t)n);
  for (long long i = 0; i < n; i++) {
    int k = (int)(std::upper_bound(cut.begin() + 1, cut.end() - 1, col[(size_t)i]) - (cut.begin() + 1));
    owner[(size_t)i] = k;
    count[(size_t)k]++;
  }
  for (int k = 0; k < N; k++) {
    Slab &s = g->slabs[(size_t)k];
    cudaSetDevice(s.device);
    if (s.h) {
      mpm_destroy(s.h);
      s.h = nullptr;
    }
    if (!s.stream) {
      if (int rc = g->cuda(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking), "stream")) return rc;
      cudaEventCreateWithFlags(&s.staged, cudaEventDisableTiming);
      cudaEventCreateWithFlags(&s.pulled, cudaEventDisableTiming);
      for (int j = 0; j < N; j++) {  // NVLink / PCIe peer access where the platform offers it (else staged copies)
        int other = g->slabs[(size_t)j].device, can = 0;
        if (other != s.device && cudaDeviceCanAccessPeer(&can, s.device, other) == cudaSuccess && can)
          if (cudaDeviceEnablePeerAccess(other, 0) != cudaSuccess) cudaGetLastError();  // already enabled is fine
      }
    }
    mpm_config c = g->cfg;
    c.device = s.device;
    c.stream = (void *)s.stream;
    c.slab_lo = s.lo = cut[(size_t)k];
    c.slab_hi = s.hi = cut[(size_t)k + 1];
    // room for migration: the user's capacity is for the whole set; a slab gets its share plus a quarter
    long long share = std::max(count[(size_t)k], g->cfg.capacity / N);
    c.capacity = share + share / 4 + 65536;
    if (N == 1) {
      c.slab_lo = 0;
      c.slab_hi = ng;
    }
    s.h = mpm_create(&c);
    if (!s.h) return g->fail(MPM_E_CUDA, std::string("mpm_group: ") + mpm_last_error(nullptr));
    mpm_slab_describe(s.h, &s.d);
    s.host.resize((size_t)count[(size_t)k] * W);
    s.ids.resize((size_t)count[(size_t)k]);
  }
  std::vector<long long> fill((size_t)N, 0);
  for (long long i = 0; i < n; i++) {
    Slab &s = g->slabs[(size_t)owner[(size_t)i]];
    long long &f = fill[(size_t)owner[(size_t)i]];
    memcpy(&s.host[(size_t)f * W], p + i * W, (size_t)W * 4);
    s.ids[(size_t)f] = (int)i;
    f++;
  }
  for (Slab &s : g->slabs) {
    cudaSetDevice(s.device);
    int rc = mpm_upload_particles_ids(s.h, s.host.data(), s.ids.data(), (long long)s.ids.size(), 0);
    if (rc) return g->fail(rc, "mpm_group: upload", &s);
  }
  g->n_total = n;
  g->staged = false;
  g->begun_dt = 0.0f;
  return MPM_OK;
}

// replaces `for (...) advance(dt)` (:214-215) on all devices; asynchronous
int mpm_group_substep(mpm_group *g, float dt, int n_steps) {
  if (!g || n_steps < 0) return MPM_E_INVALID;
  if (g->slabs.empty() || !g->slabs[0].h) return g->fail(MPM_E_STATE, "mpm_group: substep before upload");
  if (!(dt > 0)) dt = g->cfg.dt;
  if (g->slabs.size() == 1) {
    int rc = mpm_substep(g->slabs[0].h, dt, n_steps);
    return rc ? g->fail(rc, "mpm_group: substep", &g->slabs[0]) : MPM_OK;
  }
  if (n_steps == 0) return MPM_OK;
  int began = 0;
  for (Slab &s : g->slabs) {
    int rc = mpm_slab_begin(s.h, dt);
    if (rc < 0) return g->fail(rc, "mpm_group: slab_begin", &s);
    began += rc;
  }
  if (began) {
    g->mark_staged();
    if (int rc = g->exchange()) return rc;
  }
  for (int k = 0; k < n_steps; k++) {
    for (Slab &s : g->slabs) {
      int rc = mpm_slab_step(s.h, dt);
      if (rc) return g->fail(rc, "mpm_group: slab_step", &s);
    }
    g->mark_staged();
    if (int rc = g->exchange()) return rc;
  }
  for (Slab &s : g->slabs) {
    int rc = mpm_slab_settle(s.h);
    if (rc) return g->fail(rc, "mpm_group: slab_settle", &s);
  }
  return MPM_OK;
}

int mpm_group_synchronize(mpm_group *g) {
  if (!g) return MPM_E_INVALID;
  for (Slab &s : g->slabs)
    if (s.h) {
      int rc = mpm_synchronize(s.h);
      if (rc) return g->fail(rc, "mpm_group: synchronize", &s);
    }
  return MPM_OK;
}

// replaces reading the global `particles` (:220-222): records in upload order; synchronises
int mpm_group_read_particles(mpm_group *g, void *aos_out, long long n) {
  if (!g || n < 0 || n > g->n_total || (n > 0 && !aos_out)) return MPM_E_INVALID;
  const int W = g->words();
  float *out = (float *)aos_out;
  long long seen = 0;
  for (Slab &s : g->slabs) {
    if (!s.h) return g->fail(MPM_E_STATE, "mpm_group: read before upload");
    cudaSetDevice(s.device);
    long long ext = mpm_storage_extent(s.h);
    if (ext < 0) return g->fail((int)ext, "mpm_group: extent", &s);
    s.host.resize((size_t)ext * W);
    s.ids.resize((size_t)ext);
    long long got = g->slabs.size() == 1 ? ext : mpm_read_particles_ids(s.h, s.host.data(), s.ids.data(), ext, 0);
    if (g->slabs.size() == 1) {
      int rc = mpm_read_particles(s.h, aos_out, n, 0);
      return rc ? g->fail(rc, "mpm_group: read", &s) : MPM_OK;
    }
    if (got < 0) return g->fail((int)got, "mpm_group: read", &s);
    for (long long i = 0; i < got; i++) {
      const int id = s.ids[(size_t)i];
      if (id < 0) continue;  // slot of a particle that emigrated
      seen++;
      if (id < n) memcpy(out + (long long)id * W, &s.host[(size_t)i * W], (size_t)W * 4);
    }
  }
  if (seen != g->n_total) {
    char b[128];
    snprintf(b, sizeof b, "mpm_group: %lld of %lld particles accounted for", seen, g->n_total);
    return g->fail(MPM_E_STATE, b);
  }
  return MPM_OK;
}

int mpm_group_poll_status(mpm_group *g) {
  if (!g) return MPM_E_INVALID;
  int worst = MPM_OK;
  for (Slab &s : g->slabs)
    if (s.h) {
      int rc = mpm_poll_status(s.h);
      if (rc) worst = g->fail(rc, "mpm_group: status", &s);
    }
  return worst;
}

int mpm_group_slab(const mpm_group *g, int k, int *device, int *slab_lo, int *slab_hi, long long *particles) {
  if (!g || k < 0 || k >= (int)g->slabs.size()) return MPM_E_INVALID;
  const Slab &s = g->slabs[(size_t)k];
  if (device) *device = s.device;
  if (slab_lo) *slab_lo = s.lo;
  if (slab_hi) *slab_hi = s.hi;
  if (particles) *particles = s.h ? mpm_particle_count(s.h) : 0;
  return MPM_OK;
}
}
