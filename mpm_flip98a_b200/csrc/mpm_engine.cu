// mpm_engine.cu -- host side of libmpm.so: the handle, its HBM arenas, the substep schedule and the
// C-ABI of include/mpm.h.  Replaces the reference's globals and main() loop
// (cpp_validation/mls-mpm88-explained.cpp:8-26 constants, :44-47 state, :214-215 loop).
// There is no CPU fallback here on purpose: without a CUDA device every call fails with MPM_E_CUDA.
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/mpm.h"
#include "mpm_kernels.cuh"

using namespace mpm;

static thread_local std::string g_create_error;

#define MPM_CUDA(call)                                                                      \
  do {                                                                                      \
    cudaError_t e_ = (call);                                                                \
    if (e_ != cudaSuccess) {                                                                \
      char b_[512];                                                                         \
      snprintf(b_, sizeof b_, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
      this->err = b_;                                                                       \
      return MPM_E_CUDA;                                                                    \
    }                                                                                       \
  } while (0)

struct mpm_handle {
  mpm_config cfg;
  Params P;
  int D;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  long long n = 0, cap = 0;
  long long steps_since_sort = 0;
  std::string err;

  // particle storage: two SoA buffers (re-sorts ping-pong between them).  Each buffer is ONE arena carved into
  // its arrays, so the idle one can hold a whole AoS image: uploads land there with a single host-to-device copy
  // and reads leave from there with a single device-to-host copy (no chunked staging).
  std::vector<void *> allocs;
  char *arena[2] = {nullptr, nullptr};
  size_t arena_bytes = 0;
  SoA<2> s2[2];
  SoA<3> s3[2];
  int cur = 0;
  bool plain_ids = true;  // ids are the upload indices 0..n-1 (mpm_upload_particles), not caller-chosen

  // grid
  float4 *grid = nullptr, *grid_tap = nullptr;
  // fused G2P->P2G (single-GPU binned path): the P2G of the NEXT substep lands in grid_next while G2P reads
  // `grid`; the two swap each substep.  grid_read = the buffer that holds the last UPDATED grid (mpm_read_grid).
  float4 *grid_next = nullptr, *grid_read = nullptr;
  bool fused = false, p2g_ready = false;
  // 2D default path: byte maps of the 8x8-node tiles each grid buffer was scattered into (k_grid_tiles); they travel
  // with the buffers when these swap
  unsigned char *touched_g = nullptr, *touched_n = nullptr;  // of `grid` / of `grid_next`
  int tiles_x = 0, tiles_y = 0;
  bool tiles_on() const { return touched_g != nullptr; }
  // overlapped slab schedule (MPM_FLAG_OVERLAP): the fused kernel of the bins >= 2 bin columns away from the slab
  // cuts runs on a side stream while the main stream finishes the boundary bins and the caller exchanges
  // emigrants / halo columns; act_lo_end / act_hi_begin split the (sorted) active-bin list into lo | interior | hi
  bool overlap = false, side_busy = false;
  cudaStream_t side = nullptr;
  cudaEvent_t ev_ready = nullptr, ev_side_done = nullptr;
  int act_lo_end = 0, act_hi_begin = 0;
  long long slot_lo_end = 0, slot_hi_begin = 0;  // storage slots of the boundary-lo bins | interior | boundary-hi bins
  void join_side() {  // everything issued on the main stream after this sees the side stream's work
    if (side_busy) {
      cudaStreamWaitEvent(stream, ev_side_done, 0);
      side_busy = false;
    }
  }
  float p2g_dt = 0.0f;
  void *vold = nullptr;
  long long nodes = 0;
  bool tap_valid = false;

  int *status_dev = nullptr;
  // [0] = binned-P2G fallback particles (total), [1] = since the last re-sort,
  // [2] = largest displacement of one substep since the last re-sort (float bits, in cells)
  unsigned long long *stats_dev = nullptr;
  // adaptive re-sort interval (binned path, cfg.rebin_every == 0), see rebin_storage()
  int rebin_interval = 16;
  unsigned long long *stats_host = nullptr;  // pinned copy of stats_dev taken at each re-sort
  cudaEvent_t stats_ev = nullptr;
  bool stats_pending = false;
  long long stats_particle_steps = 0;
  int current_interval() const {
    if (cfg.rebin_every != 0) return cfg.rebin_every;
    return binned ? rebin_interval : 32;
  }
  bool binned = false;                      // CTA-per-bin P2G (default) vs MPM_FLAG_NAIVE
  bool deterministic = false;               // MPM_FLAG_DETERMINISTIC (mpm_deterministic.cu)
  float *det_rec = nullptr;
  int *det_cell_start = nullptr;
  long long det_cells = 0;
  int det_key_bits = 0;
  int det_sort_and_p2g(float dt);

  // binning
  BinGeom G;
  SortBuffers sb;
  // bin ranges and active-bin lists are double-buffered: during a re-sort substep the kernel walks the OLD order
  // while it writes the NEW one (bs = index of the set that describes the current storage)
  int *bin_start_buf[2] = {nullptr, nullptr};
  int *active_bins_buf[2] = {nullptr, nullptr};  // compacted ids of the non-empty bins
  int4 *chunks_buf[2] = {nullptr, nullptr};      // work list of the 2D substep kernel (see launch_active_chunks)
  unsigned *chunk_offs = nullptr;                // its scan scratch, n_bins + 4
  long long chunks_cap = 0;
  int n_chunks = 0, chunk_lo_end = 0, chunk_hi_begin = 0;  // overlap: boundary-lo | interior | boundary-hi chunks
  int bs = 0;
  int *bin_start = nullptr;         // == bin_start_buf[bs]
  int *cell_start = nullptr;        // re-sort scratch: first slot of every cell (n_bins * cpb + 4), see begin_resort
  long long n_cells() const { return (long long)G.n_bins * G.cpb; }
  unsigned *active_offs = nullptr;  // scan scratch, n_bins + 2
  int *cell_dev = nullptr;
  int key_bits = 0;
  bool resort_due = false;          // the next fused substep re-sorts on the fly (RESORT kernel variant)
  int *resort_host = nullptr;       // pinned: n_active, live extent, overlap split (lo end, hi begin)
  cudaEvent_t resort_ev = nullptr;
  bool resort_pending = false;      // begin_resort() issued, end_resort() not yet
  bool fused_resort_now = false;    // rebin_storage() called from the fused substep: only the first half

  // x-slab exchange (multi == this handle owns a strict sub-range of base-cell columns)
  bool multi = false;
  MigPtrs mig = {nullptr, nullptr, nullptr, 0, 0, 0};
  // One message per neighbour and substep: [ghost partial sums of the 2 shared node columns | header: emigrant count |
  // up to K emigrant records].  Fixed size, so the caller's transport needs no counts and the host never waits:
  // the counts stay on the device (header), and so do the storage extent and the live count (dev_ext).
  char *msg_send[2] = {nullptr, nullptr}, *msg_recv[2] = {nullptr, nullptr};
  size_t msg_bytes = 0, msg_hdr = 0;  // msg_hdr = offset of the 16-byte header (== bytes of the ghost columns)
  int *dev_ext = nullptr;             // device: [0] storage extent, [1] live particles   (multi only)
  int *hdr(char *m) const { return (int *)(m + msg_hdr); }
  float *recs(char *m) const { return (float *)(m + msg_hdr + 16); }
  enum { SLAB_FRESH = 0, SLAB_STAGED = 1, SLAB_SETTLED = 2 };
  int slab_state = SLAB_FRESH;
  bool pipelined = false;  // the grid always holds the NEXT substep's P2G (fused kernel, or every x-slab handle)
  long long n_binned = 0;  // storage slots covered by bin_start (slots beyond are immigrants since the last re-sort)
  long long live = 0;      // particles owned (storage extent n also counts dead slots)
  long long halo_nodes() const { return 2LL * P.n1 * (D == 3 ? P.n1 : 1); }
  int mig_words() const { return D == 2 ? MigRec<2>::WORDS : MigRec<3>::WORDS; }


  // per-phase CUDA-event timing (mpm_profile_*)
  bool prof_on = false;
  mpm_profile prof = {};
  struct Span {
    int phase;
    cudaEvent_t a, b;
  };
  std::vector<Span> spans;
  std::vector<cudaEvent_t> ev_pool;
  cudaEvent_t get_event() {
    cudaEvent_t e = nullptr;
    if (!ev_pool.empty()) {
      e = ev_pool.back();
      ev_pool.pop_back();
    } else {
      cudaEventCreate(&e);
    }
    return e;
  }
  void flush_spans() {  // caller has synchronised the stream
    for (Span &sp : spans) {
      float ms = 0.0f;
      if (cudaEventElapsedTime(&ms, sp.a, sp.b) == cudaSuccess) prof.ms[sp.phase] += ms;
      ev_pool.push_back(sp.a);
      ev_pool.push_back(sp.b);
    }
    spans.clear();
  }
  // RAII: times everything enqueued on the stream during its lifetime as one phase
  struct Phase {
    mpm_handle *h;
    Span sp;
    cudaStream_t st;
    Phase(mpm_handle *h_, int phase, int launches, cudaStream_t on = nullptr, bool enabled = true)
        : h(h_), st(on ? on : h_->stream) {
      sp.phase = phase;
      sp.a = sp.b = nullptr;
      if (!h->prof_on || !enabled) return;
      h->prof.launches[phase] += launches;
      sp.a = h->get_event();
      sp.b = h->get_event();
      cudaEventRecord(sp.a, st);
    }
    ~Phase() {
      if (!sp.a) return;
      cudaEventRecord(sp.b, st);
      h->spans.push_back(sp);
      if (h->spans.size() >= 8192) {
        h->join_side();
        cudaStreamSynchronize(h->stream);
        if (st != h->stream) cudaStreamSynchronize(st);  // a side-stream span: its closing event was recorded just now
        h->flush_spans();
      }
    }
  };

  int record_words() const { return D == 2 ? 14 : 26; }

  template <typename T>
  int dalloc(T **p, size_t count) {
    void *q = nullptr;
    size_t bytes = count * sizeof(T);
    if (bytes == 0) bytes = 16;
    MPM_CUDA(cudaMalloc(&q, bytes));
    allocs.push_back(q);
    *p = (T *)q;
    return MPM_OK;
  }

  int init();
  int upload(const void *aos, const int *ids, long long count, int on_device);
  int read(void *aos_out, long long count, int to_device);
  long long read_ids(void *aos_out, int *ids_out, long long max_n, int to_device);
  int slab_begin(float dt);
  int slab_step(float dt);
  int slab_settle();
  int slab_consume();
  int slab_stage();
  int sync_extent();
  int resident_p2g(float dt);
  int rebin_storage();
  int begin_resort();
  int end_resort();
  bool fast2d() const { return fused && D == 2 && !(cfg.flags & MPM_FLAG_STRICT); }
  // 3D default: P2G (k_p2g_cells) and G2P (k_g2p3, mpm_substep3d.cu) stay separate kernels; the G2P re-sorts on the fly
  bool fast3d() const {
    return binned && D == 3 && !fused && !(cfg.flags & (MPM_FLAG_STRICT | MPM_FLAG_G2P_TILE));
  }
  // MPM_FLAG_FUSE_3D: G2P and the next P2G in ONE kernel (k_substep3d)
  bool fast3f() const { return fused && D == 3; }
  bool resorts_on_the_fly() const { return fast2d() || fast3d() || fast3f(); }
  int chunk_capacity() const { return D == 2 ? substep2d_chunk_capacity() : substep3d_chunk_capacity(); }
  void carve(int b);
  int substep(float dt, int n_steps);
  int step_p2g(float dt);
  int step_grid_g2p(float dt);
  int read_grid(int stage, float *out);
  int bin_particles(int *cell, int *key, int *order, int *bin_start_out);
  int poll_status();
  template <int DD>
  GridPtrs<DD> gp() {
    GridPtrs<DD> g;
    g.g = grid;
    g.vold = vold;
    g.nodes = nodes;
    return g;
  }
  ~mpm_handle() {
    cudaSetDevice(cfg.device);
    for (void *p : allocs) cudaFree(p);
    for (Span &sp : spans) {
      cudaEventDestroy(sp.a);
      cudaEventDestroy(sp.b);
    }
    for (cudaEvent_t e : ev_pool) cudaEventDestroy(e);
    if (side) {
      cudaStreamSynchronize(side);
      cudaStreamDestroy(side);
    }
    if (ev_ready) cudaEventDestroy(ev_ready);
    if (ev_side_done) cudaEventDestroy(ev_side_done);
    if (resort_host) cudaFreeHost(resort_host);
    if (resort_ev) cudaEventDestroy(resort_ev);
    if (stats_host) cudaFreeHost(stats_host);
    if (stats_ev) cudaEventDestroy(stats_ev);
    if (own_stream && stream) cudaStreamDestroy(stream);
  }
};

static int validate(const mpm_config &c, std::string &why) {
  char b[256];
#define BAD(...)                      \
  {                                   \
    snprintf(b, sizeof b, __VA_ARGS__); \
    why = b;                          \
    return MPM_E_INVALID;             \
  }
  if (c.abi_version != MPM_ABI_VERSION) BAD("abi_version %d != %d", c.abi_version, MPM_ABI_VERSION);
  if (c.dim != 2 && c.dim != 3) BAD("dim must be 2 or 3 (got %d)", c.dim);
  if (c.n_grid < 4) BAD("n_grid must be >= 4 (got %d)", c.n_grid);
  if (c.dim == 2 && c.n_grid > 32766) BAD("n_grid too large for 2D (%d)", c.n_grid);
  if (c.dim == 3 && c.n_grid > 2046) BAD("n_grid too large for 3D (%d)", c.n_grid);
  if (c.n_materials < 1 || c.n_materials > 4) BAD("n_materials must be 1..4 (got %d)", c.n_materials);
  for (int m = 0; m < c.n_materials; m++)
    if (c.materials[m].kind < 0 || c.materials[m].kind > 2) BAD("material %d: unknown kind %d", m, c.materials[m].kind);
  if (c.capacity < 1 || c.capacity > 2000000000LL) BAD("capacity must be in [1, 2e9] (got %lld)", c.capacity);
  if (c.slab_lo < 0 || c.slab_hi > c.n_grid || c.slab_lo >= c.slab_hi) BAD("bad slab [%d,%d)", c.slab_lo, c.slab_hi);
  if ((c.slab_lo > 0 || c.slab_hi < c.n_grid) && c.slab_hi - c.slab_lo < 2) BAD("a slab must own at least 2 columns [%d,%d)", c.slab_lo, c.slab_hi);
  if (c.bin_edge < 0 || c.bin_edge > 64) BAD("bin_edge must be 0..64 (got %d)", c.bin_edge);
  if ((c.flags & MPM_FLAG_DETERMINISTIC) && (c.slab_lo > 0 || c.slab_hi < c.n_grid))
    BAD("MPM_FLAG_DETERMINISTIC needs a whole-domain handle (an x-slab cut would split a node's sum)");
  return MPM_OK;
#undef BAD
}

int mpm_handle::init() {
  D = cfg.dim;
  MPM_CUDA(cudaSetDevice(cfg.device));
  if (cfg.stream) {
    stream = (cudaStream_t)cfg.stream;
  } else {
    MPM_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    own_stream = true;
  }
  // kernel-side constants, computed in fp32 in the reference's expression order (:12-13, :25-26)
  memset(&P, 0, sizeof P);
  P.n_grid = cfg.n_grid;
  P.n1 = cfg.n_grid + 1;
  P.dx = 1.0f / cfg.n_grid;
  P.inv_dx = 1.0f / P.dx;
  P.mass_p = cfg.mass_p;
  P.vol_p = cfg.vol_p;
  for (int k = 0; k < 3; k++) P.gravity[k] = cfg.gravity[k];
  P.boundary = cfg.boundary;
  P.jp_min = cfg.jp_min;
  P.jp_max = cfg.jp_max;
  P.alpha = cfg.alpha;
  P.n_materials = cfg.n_materials;
  for (int m = 0; m < cfg.n_materials; m++) {
    const mpm_material &s = cfg.materials[m];
    Material &d = P.mat[m];
    d.kind = s.kind;
    volatile float E = s.E, nu = s.nu;  // volatile: keep the divisions in fp32, unfused, as written
    d.mu_0 = E / (2 * (1 + nu));
    d.lambda_0 = E * nu / ((1 + nu) * (1 - 2 * nu));
    d.hardening = s.hardening;
    d.sig_lo = s.sig_lo;
    d.sig_hi = s.sig_hi;
  }
  P.slab_lo = cfg.slab_lo;
  P.slab_hi = cfg.slab_hi;
  int xhi = cfg.slab_hi < cfg.n_grid - 1 ? cfg.slab_hi : cfg.n_grid - 1;  // one past the last owned base column
  P.ncol = xhi - cfg.slab_lo + 2;
  multi = cfg.slab_lo > 0 || cfg.slab_hi < cfg.n_grid;
  P.multi = multi ? 1 : 0;
  cap = cfg.capacity;

  nodes = (long long)P.ncol * P.n1 * (D == 3 ? P.n1 : 1);
  int rc;
  if ((rc = dalloc(&grid, (size_t)nodes))) return rc;
  grid_read = grid;
  if (cfg.alpha != 0.0f) {
    char *v;
    if ((rc = dalloc(&v, (size_t)nodes * (D == 2 ? 8 : 16)))) return rc;
    vold = v;
  }
  if (cfg.flags & MPM_FLAG_CAPTURE_POST_P2G)
    if ((rc = dalloc(&grid_tap, (size_t)nodes))) return rc;
  if ((rc = dalloc(&status_dev, 4))) return rc;
  MPM_CUDA(cudaMemsetAsync(status_dev, 0, 16, stream));
  if ((rc = dalloc(&stats_dev, 4))) return rc;
  MPM_CUDA(cudaMemsetAsync(stats_dev, 0, 32, stream));
  MPM_CUDA(cudaHostAlloc((void **)&stats_host, 32, cudaHostAllocDefault));
  MPM_CUDA(cudaEventCreateWithFlags(&stats_ev, cudaEventDisableTiming));

  {
    auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t c = (size_t)cap;
    arena_bytes = D == 2 ? 2 * al(c * 8) + 2 * al(c * 16) + 3 * al(c * 4) : 2 * al(c * 16) + 19 * al(c * 4);
    arena_bytes += 4096;  // an AoS image (56 | 104 B per record) plus its id array always fits: 60 | 108 B per slot
    for (int b = 0; b < 2; b++) {
      if ((rc = dalloc(&arena[b], arena_bytes))) return rc;
      carve(b);
    }
  }

  int edge = cfg.bin_edge > 0 ? cfg.bin_edge : (D == 2 ? 8 : 4);
  G = make_bin_geom(P, D, edge);
  key_bits = 1;
  while ((1LL << key_bits) < (long long)G.n_bins + (multi ? 1 : 0)) key_bits++;  // slabs: + the bin of dead slots
  sb.capacity = cap;
  for (int b = 0; b < 2; b++)
    if ((rc = dalloc(&sb.key[b], cap)) || (rc = dalloc(&sb.val[b], cap))) return rc;
  if ((rc = dalloc(&sb.hist, sort_hist_elems(cap)))) return rc;
  {  // the scan scratch serves the radix histograms and the active-bin compaction (n_bins + 1 flags)
    long long longest = (long long)sort_hist_elems(cap);
    if (n_cells() + 4 > longest) longest = n_cells() + 4;
    if ((rc = dalloc(&sb.scan_tmp, scan_tmp_elems(longest)))) return rc;
  }
  for (int b = 0; b < 2; b++)
    if ((rc = dalloc(&bin_start_buf[b], (size_t)G.n_bins + 4)) || (rc = dalloc(&active_bins_buf[b], (size_t)G.n_bins + 1)))
      return rc;
  bin_start = bin_start_buf[0];
  if ((rc = dalloc(&cell_start, (size_t)n_cells() + 4))) return rc;
  MPM_CUDA(cudaHostAlloc((void **)&resort_host, 64, cudaHostAllocDefault));
  MPM_CUDA(cudaEventCreateWithFlags(&resort_ev, cudaEventDisableTiming));
  if (multi) {
    // K = records per message.  Every handle of a decomposition must arrive at the same number (the messages have a
    // fixed size), so it depends on the GLOBAL grid only: twice the nodes of a cut (a column / plane of cells at
    // ~8-16 particles per cell times a CFL number <= 0.1-0.25), unless mpm_config.mig_records says otherwise
    // (3D: a quarter of that -- a plane of 8 particles per cell at a CFL number of 0.05 sends ~0.12 n^2 records, the two
    // planes hold 2 n^2 nodes; the full count made the 3D message 38 MB at n = 512, most of it never filled)
    long long K = cfg.mig_records > 0 ? cfg.mig_records : (D == 3 ? halo_nodes() / 4 : halo_nodes());
    if (cfg.mig_records <= 0) {
      if (K < 4096) K = 4096;
      if (K > (1 << 18)) K = 1 << 18;
    }
    mig.cap = (int)K;
    mig.enabled = 1;
    msg_hdr = (size_t)halo_nodes() * sizeof(float4);
    msg_bytes = msg_hdr + 16 + (size_t)K * mig_words() * 4;
    for (int k = 0; k < 2; k++)
      if ((rc = dalloc(&msg_send[k], msg_bytes)) || (rc = dalloc(&msg_recv[k], msg_bytes))) return rc;
    if ((rc = dalloc(&mig.count, 4)) || (rc = dalloc(&dev_ext, 4))) return rc;
    mig.send_lo = recs(msg_send[0]);
    mig.send_hi = recs(msg_send[1]);
    for (int k = 0; k < 2; k++) {
      MPM_CUDA(cudaMemsetAsync(msg_send[k], 0, msg_hdr + 16, stream));
      MPM_CUDA(cudaMemsetAsync(msg_recv[k], 0, msg_hdr + 16, stream));
    }
    MPM_CUDA(cudaMemsetAsync(mig.count, 0, 16, stream));
    MPM_CUDA(cudaMemsetAsync(dev_ext, 0, 16, stream));
  }
  if ((rc = dalloc(&active_offs, (size_t)G.n_bins + 4))) return rc;
  deterministic = (cfg.flags & MPM_FLAG_DETERMINISTIC) != 0;
  binned = !(cfg.flags & (MPM_FLAG_NAIVE | MPM_FLAG_DETERMINISTIC)) &&
           (D == 2 ? p2g_cells_supported<2>(G) : p2g_cells_supported<3>(G));
  if (deterministic) {
    const long long nc = cfg.n_grid - 1;
    det_cells = D == 2 ? nc * nc : nc * nc * nc;
    det_key_bits = 1;
    while ((1LL << det_key_bits) < det_cells) det_key_bits++;
    if ((rc = dalloc(&det_rec, (size_t)cap * (D == 2 ? 8 : 16))) || (rc = dalloc(&det_cell_start, (size_t)det_cells + 2)))
      return rc;
  }
  // 2D only: in 3D the Jacobi-SVD-heavy kernels are compute-bound and fusing them costs occupancy (measured slower)
  // fused G2P->P2G: 2D k_substep2d (MPM_FLAG_STRICT: the generic k_p2g_cells<FUSED>); 3D only on request
  // (MPM_FLAG_FUSE_3D: k_substep3d, measured no faster than the two 3D kernels; 3D + STRICT always stays unfused)
  fused = binned && !(cfg.flags & (MPM_FLAG_NO_FUSE | MPM_FLAG_G2P_TILE)) &&
          (D == 2 || ((cfg.flags & MPM_FLAG_FUSE_3D) && !(cfg.flags & MPM_FLAG_STRICT)));
  pipelined = fused || multi;
  if (pipelined)
    if ((rc = dalloc(&grid_next, (size_t)nodes))) return rc;
  if (D == 3 && binned) {  // work list of the 3D chunk kernels (k_g2p3_tile, k_substep3d)
    chunks_cap = (long long)G.n_bins + cap / chunk_capacity() + 2;
    for (int b = 0; b < 2; b++)
      if ((rc = dalloc(&chunks_buf[b], (size_t)chunks_cap))) return rc;
    if ((rc = dalloc(&chunk_offs, (size_t)G.n_bins + 4))) return rc;
  }
  if (fast2d()) {
    tiles_x = (P.ncol + 7) / 8;
    tiles_y = (P.n1 + 7) / 8;
    const size_t tb = (size_t)tiles_x * tiles_y;
    if ((rc = dalloc(&touched_g, tb)) || (rc = dalloc(&touched_n, tb))) return rc;
    MPM_CUDA(cudaMemsetAsync(touched_g, 0, tb, stream));
    MPM_CUDA(cudaMemsetAsync(touched_n, 0, tb, stream));
    // an unmarked tile must hold zeros
    MPM_CUDA(cudaMemsetAsync(grid, 0, (size_t)nodes * sizeof(float4), stream));
    MPM_CUDA(cudaMemsetAsync(grid_next, 0, (size_t)nodes * sizeof(float4), stream));
    // one entry per chunk: at most one per non-empty bin plus one per full chunk of particles
    chunks_cap = (long long)G.n_bins + cap / substep2d_chunk_capacity() + 2;
    for (int b = 0; b < 2; b++)
      if ((rc = dalloc(&chunks_buf[b], (size_t)chunks_cap))) return rc;
    if ((rc = dalloc(&chunk_offs, (size_t)G.n_bins + 4))) return rc;
  }
  overlap = (fused || fast3d()) && multi && (cfg.flags & MPM_FLAG_OVERLAP);
  if (overlap) {
    {
      // lowest priority: when SM slots free up, boundary kernels (main stream) and the caller's NCCL kernels are
      // placed before further interior CTAs -- otherwise the many small interior CTAs starve them until the tail
      int least = 0, greatest = 0;
      MPM_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
      MPM_CUDA(cudaStreamCreateWithPriority(&side, cudaStreamNonBlocking, least));
    }
    MPM_CUDA(cudaEventCreateWithFlags(&ev_ready, cudaEventDisableTiming));
    MPM_CUDA(cudaEventCreateWithFlags(&ev_side_done, cudaEventDisableTiming));
  }

  MPM_CUDA(cudaStreamSynchronize(stream));
  return MPM_OK;
}

void mpm_handle::carve(int b) {
  auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
  char *p = arena[b];
  const size_t c = (size_t)cap;
  auto take = [&](size_t bytes) {
    char *q = p;
    p += al(bytes);
    return q;
  };
  if (D == 2) {
    SoA<2> &s = s2[b];
    s.C = (float4 *)take(c * 16);
    s.F = (float4 *)take(c * 16);
    s.x = (float2 *)take(c * 8);
    s.v = (float2 *)take(c * 8);
    s.Jp = (float *)take(c * 4);
    s.mat = (int *)take(c * 4);
    s.id = (int *)take(c * 4);
  } else {
    SoA<3> &s = s3[b];
    s.xj = (float4 *)take(c * 16);
    s.vm = (float4 *)take(c * 16);
    for (int k = 0; k < 9; k++) s.C[k] = (float *)take(c * 4);
    for (int k = 0; k < 9; k++) s.F[k] = (float *)take(c * 4);
    s.id = (int *)take(c * 4);
  }
}

int mpm_handle::upload(const void *aos, const int *ids, long long count, int on_device) {
  join_side();
  if (count < 0 || (count > 0 && !aos)) {
    err = "upload: bad arguments";
    return MPM_E_INVALID;
  }
  if (count > cap) {
    err = "upload: more particles than capacity";
    return MPM_E_CAPACITY;
  }
  MPM_CUDA(cudaSetDevice(cfg.device));
  if (resort_pending) end_resort();
  const int W = record_words();
  if (count > 0) {
    // one copy of the whole AoS image into the idle storage arena, one conversion kernel out of it
    const float *dev_src = (const float *)aos;
    const int *dev_ids = ids;
    if (!on_device) {
      char *img = arena[cur ^ 1];
      const size_t img_bytes = (size_t)count * W * 4, ids_off = (img_bytes + 255) & ~(size_t)255;
      MPM_CUDA(cudaMemcpyAsync(img, aos, img_bytes, cudaMemcpyHostToDevice, stream));
      dev_src = (const float *)img;
      if (ids) {
        MPM_CUDA(cudaMemcpyAsync(img + ids_off, ids, (size_t)count * 4, cudaMemcpyHostToDevice, stream));
        dev_ids = (const int *)(img + ids_off);
      }
    }
    if (D == 2) launch_aos_to_soa<2>(dev_src, 0, count, s2[cur], dev_ids, stream);
    else launch_aos_to_soa<3>(dev_src, 0, count, s3[cur], dev_ids, stream);
  }
  MPM_CUDA(cudaGetLastError());
  n = count;
  live = count;
  plain_ids = ids == nullptr;
  p2g_ready = false;
  resort_due = false;
  grid_read = grid;
  tap_valid = false;
  slab_state = SLAB_FRESH;
  if (multi) {
    const int e2[2] = {(int)count, (int)count};
    MPM_CUDA(cudaMemcpyAsync(dev_ext, e2, 8, cudaMemcpyHostToDevice, stream));
    MPM_CUDA(cudaStreamSynchronize(stream));  // e2 is on this stack frame
    MPM_CUDA(cudaMemsetAsync(mig.count, 0, 16, stream));
  }
  int rc = rebin_storage();
  if (rc) return rc;
  MPM_CUDA(cudaStreamSynchronize(stream));  // the caller may free `aos` on return
  return MPM_OK;
}

// First half of a storage re-sort: every slot of the current storage takes (new bin, rank in it) from its current
// position, the per-bin counts are scanned into the NEW bin starts, the non-empty bins are compacted into the NEW
// active list, and the few numbers the host needs start their way back.  Everything is enqueued on the stream and
// nothing waits: the consumer (k_reorder_scatter, or the RESORT substep kernel) is enqueued right behind and
// end_resort() only synchronises on the small read-back, not on the consumer.
int mpm_handle::begin_resort() {
  int *ns = bin_start_buf[bs ^ 1];
  int *na = active_bins_buf[bs ^ 1];
  const long long nc = n_cells();
  MPM_CUDA(cudaMemsetAsync(cell_start, 0, ((size_t)nc + 4) * sizeof(int), stream));
  if (D == 2) launch_count_rank<2>(P, G, s2[cur], n, (unsigned *)cell_start, sb.key[0], (unsigned *)sb.val[0], status_dev, stream, dev_ext);
  else launch_count_rank<3>(P, G, s3[cur], n, (unsigned *)cell_start, sb.key[0], (unsigned *)sb.val[0], status_dev, stream, dev_ext);
  // cell_start[c] = first slot of cell c (cells of a bin are consecutive); cell_start[nc] = live extent (dead slots
  // sort behind every cell); cell_start[nc + 1] = n.  Every cpb-th entry is a bin start.
  exclusive_scan_u32((unsigned *)cell_start, nc + 2, sb.scan_tmp, stream);
  launch_bin_starts_from_cells(cell_start, G.n_bins, G.cpb, ns, stream);
  launch_active_bins(ns, G.n_bins, active_offs, sb.scan_tmp, na, stream);
  MPM_CUDA(cudaMemcpyAsync(&resort_host[0], active_offs + G.n_bins, 4, cudaMemcpyDeviceToHost, stream));
  MPM_CUDA(cudaMemcpyAsync(&resort_host[1], ns + G.n_bins, 4, cudaMemcpyDeviceToHost, stream));
  if (chunk_offs) {
    launch_active_chunks(ns, G.n_bins, G.nb[1], chunk_capacity(), chunk_offs, sb.scan_tmp, chunks_buf[bs ^ 1], stream,
                         D == 3 ? G.nb[2] : 0);
    MPM_CUDA(cudaMemcpyAsync(&resort_host[4], chunk_offs + G.n_bins, 4, cudaMemcpyDeviceToHost, stream));
  }
  if (overlap) {
    // the active list is sorted by bin id (x-major): boundary-lo bins are a prefix, boundary-hi bins a suffix;
    // the scan scratch holds "active bins with a smaller id" for every bin
    const int col = G.nb[1] * G.nb[2];
    const int lo_bin = cfg.slab_lo > 0 ? (2 < G.nb[0] ? 2 : G.nb[0]) * col : 0;
    const int hi_bin = cfg.slab_hi < cfg.n_grid ? (G.nb[0] - 2 > 0 ? G.nb[0] - 2 : 0) * col : G.n_bins;
    MPM_CUDA(cudaMemcpyAsync(&resort_host[2], active_offs + lo_bin, 4, cudaMemcpyDeviceToHost, stream));
    MPM_CUDA(cudaMemcpyAsync(&resort_host[3], active_offs + hi_bin, 4, cudaMemcpyDeviceToHost, stream));
    if (chunk_offs) {
      MPM_CUDA(cudaMemcpyAsync(&resort_host[5], chunk_offs + lo_bin, 4, cudaMemcpyDeviceToHost, stream));
      MPM_CUDA(cudaMemcpyAsync(&resort_host[6], chunk_offs + hi_bin, 4, cudaMemcpyDeviceToHost, stream));
    }
    // ... and so are their storage slots (the storage is sorted by bin): first slot of the interior, of boundary-hi
    MPM_CUDA(cudaMemcpyAsync(&resort_host[7], ns + lo_bin, 4, cudaMemcpyDeviceToHost, stream));
    MPM_CUDA(cudaMemcpyAsync(&resort_host[8], ns + hi_bin, 4, cudaMemcpyDeviceToHost, stream));
  }
  MPM_CUDA(cudaEventRecord(resort_ev, stream));
  resort_pending = true;
  return MPM_OK;
}

// Second half: the consumer has been enqueued; adopt the new order on the host side.
int mpm_handle::end_resort() {
  if (!resort_pending) return MPM_OK;
  MPM_CUDA(cudaEventSynchronize(resort_ev));
  resort_pending = false;
  cur ^= 1;
  bs ^= 1;
  bin_start = bin_start_buf[bs];
  G.active = binned ? active_bins_buf[bs] : nullptr;
  G.n_active = binned ? resort_host[0] : 0;
  n = resort_host[1];  // dead (emigrated) slots are gone
  n_binned = n;
  act_lo_end = 0;
  act_hi_begin = G.n_active;
  slot_lo_end = 0;
  slot_hi_begin = n;
  if (overlap) {
    act_lo_end = resort_host[2];
    act_hi_begin = resort_host[3] > act_lo_end ? resort_host[3] : act_lo_end;
    slot_lo_end = resort_host[7];
    slot_hi_begin = resort_host[8] > slot_lo_end ? resort_host[8] : slot_lo_end;
  }
  if (chunk_offs) {
    n_chunks = resort_host[4];
    chunk_lo_end = 0;
    chunk_hi_begin = n_chunks;
    if (overlap) {
      chunk_lo_end = resort_host[5];
      chunk_hi_begin = resort_host[6] > chunk_lo_end ? resort_host[6] : chunk_lo_end;
    }
  }
  return MPM_OK;
}

// stand-alone re-sort: count + scan, then a scatter of the whole storage into the other buffer
int mpm_handle::rebin_storage() {
  join_side();
  if (resort_pending) end_resort();
  if (deterministic) {
    // the counting re-sort ranks with atomics (its order inside a cell is not reproducible); this mode keeps the
    // upload order and renews it with a STABLE sort by cell before every substep (det_sort_and_p2g)
    steps_since_sort = 0;
    n_binned = n;
    return MPM_OK;
  }
  if (binned && cfg.rebin_every == 0) {
    // Adaptive re-sort interval.  The binned kernels tolerate a particle that drifted up to MARGIN = 1 cell out of
    // its bin; beyond that it takes the per-particle scatter (correct, but as slow as the naive path), and once the
    // fastest particles have moved a cell the fallback share explodes.  So the interval follows the CFL number the
    // kernels measure (stats[2] = largest displacement of one substep, in cells, since the previous re-sort):
    //     interval = 0.75 * MARGIN / displacement,
    // growing at most 2x per re-sort; a measured fallback share above 0.1 % of particle-steps additionally cuts it
    // by a quarter (paths whose kernels do not report a displacement only use this second rule, additively
    // increasing after a clean interval).  The numbers arrive one interval late (asynchronous copy).
    if (stats_pending && cudaEventQuery(stats_ev) == cudaSuccess) {
      const double frac = stats_particle_steps > 0 ? (double)stats_host[1] / (double)stats_particle_steps : 0.0;
      float disp = 0.0f;
      {
        const unsigned bits = (unsigned)(stats_host[2] & 0xffffffffull);
        memcpy(&disp, &bits, 4);
      }
      int next = rebin_interval;
      if (disp > 0.0f && disp < 1e3f) {
        const double target = 0.75 / (double)disp;
        next = target > 512.0 ? 512 : (int)target;
        if (next > 2 * rebin_interval) next = 2 * rebin_interval;
      } else if (frac < 1e-4) {
        next = rebin_interval + (rebin_interval / 4 > 1 ? rebin_interval / 4 : 1);
      }
      if (frac > 1e-3 && next > rebin_interval - rebin_interval / 4) next = rebin_interval - rebin_interval / 4;
      rebin_interval = next < 2 ? 2 : (next > 512 ? 512 : next);
      stats_pending = false;
    }
    if (!stats_pending && steps_since_sort > 0) {
      MPM_CUDA(cudaMemcpyAsync(stats_host, stats_dev, 32, cudaMemcpyDeviceToHost, stream));
      MPM_CUDA(cudaEventRecord(stats_ev, stream));
      MPM_CUDA(cudaMemsetAsync(stats_dev + 1, 0, 16, stream));
      stats_particle_steps = (multi ? n : live) * steps_since_sort;
      stats_pending = true;
    }
  }
  steps_since_sort = 0;
  resort_due = false;
  if (n == 0) {  // nothing resident: no bin holds anything (do not leave the previous upload's ranges behind)
    n_binned = 0;
    G.n_active = 0;
    act_lo_end = act_hi_begin = 0;
    n_chunks = chunk_lo_end = chunk_hi_begin = 0;
    return MPM_OK;
  }
  if (!fused_resort_now) {
    Phase ph(this, MPM_PHASE_BIN, 7);
    int rc = begin_resort();
    if (rc) return rc;
    const int *ns = bin_start_buf[bs ^ 1];
    if (D == 2) launch_reorder_scatter<2>(s2[cur], s2[cur ^ 1], 0, n, (int)n_cells(), cell_start, sb.key[0], (unsigned *)sb.val[0], stream, dev_ext);
    else launch_reorder_scatter<3>(s3[cur], s3[cur ^ 1], 0, n, (int)n_cells(), cell_start, sb.key[0], (unsigned *)sb.val[0], stream, dev_ext);
    if (multi)  // the new extent (dead slots dropped) = first slot of the "dead" bin
      launch_slab_counters(dev_ext, nullptr, nullptr, nullptr, nullptr, 0, cap, ns + G.n_bins, stream);
    MPM_CUDA(cudaGetLastError());
    return end_resort();
  }
  // on-the-fly variant: the caller (step_grid_g2p) enqueues the RESORT substep kernel as the consumer
  Phase ph(this, MPM_PHASE_BIN, 6);
  return begin_resort();
}

int mpm_handle::read(void *aos_out, long long count, int to_device) {
  join_side();
  if (multi) sync_extent();
  if (count < 0 || count > live || (count > 0 && !aos_out)) {
    err = "read: bad arguments";
    return MPM_E_INVALID;
  }
  MPM_CUDA(cudaSetDevice(cfg.device));
  if (count > 0) {
    // one pass: every live particle whose id lies in [0, count) writes its record to image[id]; the image is the
    // caller's device buffer, or the idle storage arena followed by ONE device-to-host copy
    const int W = record_words();
    float *img = to_device ? (float *)aos_out : (float *)arena[cur ^ 1];
    if (D == 2) launch_soa_to_aos<2>(s2[cur], n, 0, count, img, nullptr, stream);
    else launch_soa_to_aos<3>(s3[cur], n, 0, count, img, nullptr, stream);
    if (!to_device) MPM_CUDA(cudaMemcpyAsync(aos_out, img, (size_t)count * W * 4, cudaMemcpyDeviceToHost, stream));
  }
  MPM_CUDA(cudaGetLastError());
  MPM_CUDA(cudaStreamSynchronize(stream));
  return MPM_OK;
}

// grid reset + P2G of every resident particle into `grid` (:50-102)
int mpm_handle::resident_p2g(float dt) {
  const bool strict = (cfg.flags & MPM_FLAG_STRICT) != 0;
  join_side();
  {
    Phase ph(this, MPM_PHASE_CLEAR, 0);
    MPM_CUDA(cudaMemsetAsync(grid, 0, (size_t)nodes * sizeof(float4), stream));  // :50
  }
  Phase ph(this, MPM_PHASE_P2G, n > 0 ? 1 : 0);
  if (binned) {
    if (D == 2) launch_p2g_cells<2>(P, G, dt, s2[cur], n_binned, bin_start, gp<2>(), status_dev, stats_dev, strict, stream);
    else launch_p2g_cells<3>(P, G, dt, s3[cur], n_binned, bin_start, gp<3>(), status_dev, stats_dev, strict, stream);
    // immigrants since the last re-sort sit behind the binned range: per-particle scatter
    if (D == 2) launch_p2g_naive<2>(P, dt, s2[cur], n_binned, n, gp<2>(), status_dev, stream, dev_ext);
    else launch_p2g_naive<3>(P, dt, s3[cur], n_binned, n, gp<3>(), status_dev, stream, dev_ext);
  } else {
    if (D == 2) launch_p2g_naive<2>(P, dt, s2[cur], 0, n, gp<2>(), status_dev, stream, dev_ext);
    else launch_p2g_naive<3>(P, dt, s3[cur], 0, n, gp<3>(), status_dev, stream, dev_ext);
  }
  if (tiles_on()) MPM_CUDA(cudaMemsetAsync(touched_g, 1, (size_t)tiles_x * tiles_y, stream));  // written everywhere
  p2g_ready = true;  // `grid` holds P2G(dt) of the current particle state
  p2g_dt = dt;
  grid_read = grid;
  return MPM_OK;
}

// Phase 1 of a substep: grid reset + P2G (:50-102).  On the pipelined schedule the grid usually already holds
// this P2G (the previous substep produced it) and nothing is launched.
int mpm_handle::step_p2g(float dt) {
  if (pipelined && p2g_ready && p2g_dt == dt) {
    grid_read = grid;
    return MPM_OK;
  }
  return resident_p2g(dt);
}

// Phases 2+3: grid update (:105-131) and G2P (:134-179).  Fused schedule: G2P runs in one kernel with the
// P2G of the NEXT substep, which lands in the other grid buffer; the buffers then swap.
int mpm_handle::step_grid_g2p(float dt) {
  if (!p2g_ready) {
    err = "step_grid_g2p: no P2G on the grid (call mpm_step_p2g first)";
    return MPM_E_STATE;
  }
  join_side();  // the previous substep's interior launch wrote particles and REDs into this grid
  if (grid_tap) {
    MPM_CUDA(cudaMemcpyAsync(grid_tap, grid, (size_t)nodes * sizeof(float4), cudaMemcpyDeviceToDevice, stream));
    tap_valid = true;
  }
  {
    Phase ph(this, MPM_PHASE_GRID, 1);
    if (tiles_on()) {
      // update of this substep's grid and reset of the next P2G target in one launch, touched tiles only
      launch_grid_tiles(P, dt, grid, vold, touched_g, grid_next, touched_n, tiles_x, tiles_y, stream, stats_dev,
                        overlap ? (int)steps_since_sort + 1 : 0, status_dev);
      if (multi) {
        // the two tile columns next to each cut: ghost sums, migrating particles and the immigrant tail land there
        // through kernels that do not mark tiles themselves
        const size_t col2 = (size_t)(tiles_x < 2 ? tiles_x : 2) * tiles_y;
        if (cfg.slab_lo > 0) MPM_CUDA(cudaMemsetAsync(touched_n, 1, col2, stream));
        if (cfg.slab_hi < cfg.n_grid)
          MPM_CUDA(cudaMemsetAsync(touched_n + (size_t)tiles_x * tiles_y - col2, 1, col2, stream));
      }
    } else if (D == 2) {
      launch_grid_update<2>(P, dt, gp<2>(), stream);
    } else {
      launch_grid_update<3>(P, dt, gp<3>(), stream);
    }
  }
  // exact association everywhere under MPM_FLAG_STRICT and on the naive path (the bit-faithful modes)
  const bool strict = (cfg.flags & (MPM_FLAG_STRICT | MPM_FLAG_NAIVE | MPM_FLAG_DETERMINISTIC)) != 0;
  if (multi) MPM_CUDA(cudaMemsetAsync(mig.count, 0, 8, stream));
  if (multi && !resorts_on_the_fly() && resort_due) {  // paths without the on-the-fly variant re-sort stand-alone, now
    int rc = rebin_storage();
    if (rc) return rc;
  }
  if (fused) {
    if (!tiles_on()) {
      Phase ph(this, MPM_PHASE_CLEAR, 0);
      MPM_CUDA(cudaMemsetAsync(grid_next, 0, (size_t)nodes * sizeof(float4), stream));  // :50 of the next substep
    }
    // 2D default: the packed-math kernel of mpm_substep2d.cu; MPM_FLAG_STRICT (and 3D, if ever fused) keeps
    // k_p2g_cells<FUSED>.  A due re-sort rides on the fast kernel (RESORT variant): first half here, the kernel
    // is the consumer, second half after it has been enqueued.
    const bool fast = fast2d(), fast3 = fast3f();
    const bool resort = (fast || fast3) && resort_due && n > 0;
    if (resort) {
      fused_resort_now = true;
      int rc = rebin_storage();
      fused_resort_now = false;
      if (rc) return rc;
    }
    Substep2dArgs sa;
    if (fast) {
      sa.P = P;
      sa.G = G;
      sa.dt_g2p = dt;
      sa.dt_p2g = dt;
      sa.s = s2[cur];
      sa.d = s2[cur ^ 1];
      sa.chunks = chunks_buf[bs];
      sa.n_chunks = n_chunks;
      sa.new_start = cell_start;
      sa.key = sb.key[0];
      sa.rank = (const unsigned *)sb.val[0];
      sa.grid_in = grid;
      sa.vold_in = (const float2 *)vold;
      sa.grid_out = grid_next;
      sa.touched_out = touched_n;
      sa.tiles_x = tiles_x;
      sa.tiles_y = tiles_y;
      sa.status = status_dev;
      sa.stats = stats_dev;
      sa.mig = mig;
    }
    Substep3dArgs sa3;
    if (fast3) {
      sa3.P = P;
      sa3.dt_g2p = dt;
      sa3.dt_p2g = dt;
      sa3.s = s3[cur];
      sa3.d = s3[cur ^ 1];
      sa3.chunks = chunks_buf[bs];
      sa3.n_chunks = n_chunks;
      sa3.new_start = cell_start;
      sa3.key = sb.key[0];
      sa3.rank = (const unsigned *)sb.val[0];
      sa3.grid_in = grid;
      sa3.vold_in = (const float4 *)vold;
      sa3.grid_out = grid_next;
      sa3.status = status_dev;
      sa3.stats = stats_dev;
      sa3.mig = mig;
    }
    const bool flip = P.alpha != 0.0f;
    // part: 0 = everything, 1 = boundary-lo, 2 = interior, 3 = boundary-hi (overlapped schedule)
    auto run_bins = [&](const BinGeom &Gx, MigPtrs mg, cudaStream_t st, int part) {
      if (fast3) {
        Substep3dArgs x = sa3;
        x.mig = mg;
        const int c0 = part == 0 || part == 1 ? 0 : (part == 2 ? chunk_lo_end : chunk_hi_begin);
        const int c1 = part == 0 || part == 3 ? n_chunks : (part == 1 ? chunk_lo_end : chunk_hi_begin);
        x.chunks = sa3.chunks + c0;
        x.n_chunks = c1 - c0;
        launch_substep3d(x, flip, mg.enabled != 0, resort, st);
      } else if (fast) {
        Substep2dArgs x = sa;
        x.G = Gx;
        x.mig = mg;
        const int c0 = part == 0 || part == 1 ? 0 : (part == 2 ? chunk_lo_end : chunk_hi_begin);
        const int c1 = part == 0 || part == 3 ? n_chunks : (part == 1 ? chunk_lo_end : chunk_hi_begin);
        x.chunks = sa.chunks + c0;
        x.n_chunks = c1 - c0;
        // the interior launch of the overlapped schedule runs the plain (single-GPU) variant: its chunks hold no dead
        // slots (only boundary chunks lose particles, immigrants sit in the tail) and none of its particles can reach
        // the cut -- guaranteed by the displacement guard in k_grid_tiles, not by a per-particle test
        launch_substep2d(x, flip, mg.enabled != 0 && part != 2, resort, st);
      } else if (D == 2) {
        launch_g2p2g<2>(P, Gx, dt, dt, s2[cur], n_binned, bin_start, gp<2>(), grid_next, status_dev, stats_dev, mg, strict, st);
      } else {
        launch_g2p2g<3>(P, Gx, dt, dt, s3[cur], n_binned, bin_start, gp<3>(), grid_next, status_dev, stats_dev, mg, strict, st);
      }
    };
    // immigrants since the last re-sort sit behind the binned range: G2P in place, their share of the next P2G,
    // and -- on a re-sort substep -- their move into the new order
    auto run_tail = [&]() {
      if (n <= n_binned) return;
      if (D == 2) {
        launch_g2p_naive<2>(P, dt, s2[cur], n_binned, n, gp<2>(), mig, status_dev, strict, stream, nullptr, dev_ext);
        GridPtrs<2> gn = gp<2>();
        gn.g = grid_next;
        launch_p2g_naive<2>(P, dt, s2[cur], n_binned, n, gn, status_dev, stream, dev_ext);
        if (resort)
          launch_reorder_scatter<2>(s2[cur], s2[cur ^ 1], n_binned, n, (int)n_cells(), cell_start, sb.key[0],
                                    (const unsigned *)sb.val[0], stream, dev_ext);
      } else {
        launch_g2p_naive<3>(P, dt, s3[cur], n_binned, n, gp<3>(), mig, status_dev, strict, stream, nullptr, dev_ext);
        GridPtrs<3> gn = gp<3>();
        gn.g = grid_next;
        launch_p2g_naive<3>(P, dt, s3[cur], n_binned, n, gn, status_dev, stream, dev_ext);
        if (resort)
          launch_reorder_scatter<3>(s3[cur], s3[cur ^ 1], n_binned, n, (int)n_cells(), cell_start, sb.key[0],
                                    (const unsigned *)sb.val[0], stream, dev_ext);
      }
    };
    if (overlap && act_hi_begin > act_lo_end && (fast || fast3)) {
      MPM_CUDA(cudaEventRecord(ev_ready, stream));  // grid updated, next grid cleared, (re-sort tables ready)
      {
        // boundary bins (both sides) FIRST and on the main stream: their emigrants and shared columns are what
        // the caller exchanges next.  (Enqueued before the interior launch on purpose: the hardware dispatches
        // CTAs of concurrent kernels in launch order, so the other order would run them after the interior.)
        Phase ph(this, MPM_PHASE_MIGRATE, 2);  // accounted with the migration work they feed
        BinGeom Gb = G;
        Gb.n_active = act_lo_end;
        if (Gb.n_active > 0) run_bins(Gb, mig, stream, 1);
        Gb.active = G.active + act_hi_begin;
        Gb.n_active = G.n_active - act_hi_begin;
        if (Gb.n_active > 0) run_bins(Gb, mig, stream, 3);
        run_tail();
      }
      // interior bins on the side stream: run while the caller exchanges the boundary
      MPM_CUDA(cudaStreamWaitEvent(side, ev_ready, 0));
      {
        Phase phs(this, MPM_PHASE_G2P, 1, side);
        BinGeom Gi = G;
        Gi.active = G.active + act_lo_end;
        Gi.n_active = act_hi_begin - act_lo_end;
        MigPtrs mi = mig;
        mi.interior = 1;
        run_bins(Gi, mi, side, 2);
      }
      MPM_CUDA(cudaEventRecord(ev_side_done, side));
      side_busy = true;
    } else {
      Phase ph(this, MPM_PHASE_G2P, n > 0 ? 1 : 0);
      if (n_binned > 0) run_bins(G, mig, stream, 0);
      run_tail();
    }
    if (resort) {
      if (multi)  // the new extent (dead slots dropped) = first slot of the "dead" bin of the new order
        launch_slab_counters(dev_ext, nullptr, nullptr, nullptr, nullptr, 0, cap, bin_start_buf[bs ^ 1] + G.n_bins, stream);
      int rc = end_resort();
      if (rc) return rc;
    }
    grid_read = grid;  // the updated grid of this substep stays readable (mpm_read_grid)
    float4 *t = grid;
    grid = grid_next;
    grid_next = t;
    unsigned char *tt = touched_g;
    touched_g = touched_n;
    touched_n = tt;
    p2g_ready = true;  // ... up to migrating particles, which the x-slab exchange adds (slab_stage / slab_consume)
    p2g_dt = dt;
    if (prof_on) prof.fused_substeps++;
  } else {
    if (pipelined) {
      Phase ph(this, MPM_PHASE_CLEAR, 0);
      MPM_CUDA(cudaMemsetAsync(grid_next, 0, (size_t)nodes * sizeof(float4), stream));  // :50 of the next substep
    }
    // 3D default: the G2P kernel of mpm_substep3d.cu; a due re-sort rides on it (RESORT variant): first half here,
    // the kernel is the consumer, second half after it has been enqueued
    const bool fast3 = fast3d();
    const bool resort3 = fast3 && resort_due && n > 0;
    bool p2g_overlapped = false;  // the overlapped 3D schedule below has already launched the next P2G
    if (resort3) {
      fused_resort_now = true;
      int rc = rebin_storage();
      fused_resort_now = false;
      if (rc) return rc;
    }
#ifndef MPM_G2P3_TILE
#define MPM_G2P3_TILE 0  // measured on c5 (B200): tile variant 1.83-2.23 ms at 5-7 CTAs/SM vs 1.48 ms thread-per-particle
#endif
    // Overlapped two-kernel schedule (MPM_FLAG_OVERLAP, no re-sort in this substep): see below.  Its launches are timed
    // by their own spans (boundary under MPM_PHASE_MIGRATE, interior under MPM_PHASE_G2P on the side stream), so the
    // enclosing G2P span stays off -- it would count the boundary work a second time.
    const bool ov3 = fast3 && !MPM_G2P3_TILE && overlap && !resort3 && n > 0 && act_hi_begin > act_lo_end &&
                     slot_hi_begin > slot_lo_end;
    if (fast3) {
      Phase ph(this, MPM_PHASE_G2P, n > 0 && !ov3 ? 1 : 0, nullptr, !ov3);
      G2p3Args ga;
      ga.P = P;
      ga.dt = dt;
      ga.s = s3[cur];
      ga.d = s3[cur ^ 1];
      ga.first = 0;
      ga.n = n;
      ga.grid = grid;
      ga.vold = (const float4 *)vold;
      ga.new_start = cell_start;
      ga.key = sb.key[0];
      ga.rank = (const unsigned *)sb.val[0];
      ga.mig = mig;
      ga.status = status_dev;
      ga.stats = stats_dev;
      ga.dev_n = dev_ext;
      ga.chunks = chunks_buf[bs];
      ga.n_chunks = n_chunks;
      if (MPM_G2P3_TILE && chunk_offs) {
        // binned range: CTA per chunk with the node tile in shared memory; immigrant tail: thread per particle
        launch_g2p3_tile(ga, P.alpha != 0.0f, mig.enabled != 0, resort3, stream);
        ga.first = n_binned;
      }
      // Overlapped two-kernel schedule (MPM_FLAG_OVERLAP, no re-sort in this substep): G2P and the next P2G of the two bin
      // columns next to each cut (a prefix and a suffix of the bin-sorted storage, plus the immigrant tail) run first on
      // the main stream -- their emigrants and shared node planes are what the caller exchanges next -- while G2P + P2G of
      // the interior follow on the low-priority side stream.  An interior particle can neither emigrate nor touch a
      // shared plane: checked per particle by the kernel (mig.interior), flagged as MPM_E_CFL.
      if (ov3) {
        const bool flip3 = P.alpha != 0.0f;
        const bool strict_p2g = false;  // fast3d() excludes MPM_FLAG_STRICT
        GridPtrs<3> gn = gp<3>();
        gn.g = grid_next;
        MPM_CUDA(cudaEventRecord(ev_ready, stream));  // grid updated, next grid cleared
        {
          Phase phb(this, MPM_PHASE_MIGRATE, 5);
          G2p3Args gb = ga;
          gb.first = 0;
          gb.n = slot_lo_end;
          launch_g2p3(gb, flip3, true, false, stream);
          gb.first = slot_hi_begin;
          gb.n = n;  // boundary-hi bins and the immigrant tail
          launch_g2p3(gb, flip3, true, false, stream);
          BinGeom Gb = G;
          Gb.n_active = act_lo_end;
          if (Gb.n_active > 0)
            launch_p2g_cells<3>(P, Gb, dt, s3[cur], n_binned, bin_start, gn, status_dev, stats_dev, strict_p2g, stream);
          Gb.active = G.active + act_hi_begin;
          Gb.n_active = G.n_active - act_hi_begin;
          if (Gb.n_active > 0)
            launch_p2g_cells<3>(P, Gb, dt, s3[cur], n_binned, bin_start, gn, status_dev, stats_dev, strict_p2g, stream);
          launch_p2g_naive<3>(P, dt, s3[cur], n_binned, n, gn, status_dev, stream, dev_ext);
        }
        MPM_CUDA(cudaStreamWaitEvent(side, ev_ready, 0));
        {
          Phase phs(this, MPM_PHASE_G2P, 2, side);
          G2p3Args gi = ga;
          gi.first = slot_lo_end;
          gi.n = slot_hi_begin;
          gi.dev_n = nullptr;  // the range lies inside the binned storage
          gi.mig.interior = 1;
          launch_g2p3(gi, flip3, true, false, side);
          BinGeom Gi = G;
          Gi.active = G.active + act_lo_end;
          Gi.n_active = act_hi_begin - act_lo_end;
          launch_p2g_cells<3>(P, Gi, dt, s3[cur], n_binned, bin_start, gn, status_dev, stats_dev, strict_p2g, side);
        }
        MPM_CUDA(cudaEventRecord(ev_side_done, side));
        side_busy = true;
        p2g_overlapped = true;
      } else
      launch_g2p3(ga, P.alpha != 0.0f, mig.enabled != 0, resort3, stream);
      if (resort3) {
        if (multi)  // the new extent (dead slots dropped) = first slot of the "dead" bin of the new order
          launch_slab_counters(dev_ext, nullptr, nullptr, nullptr, nullptr, 0, cap, bin_start_buf[bs ^ 1] + G.n_bins, stream);
        int rc = end_resort();
        if (rc) return rc;
      }
    } else {
      Phase ph(this, MPM_PHASE_G2P, n > 0 ? 1 : 0);
      if (binned && (cfg.flags & MPM_FLAG_G2P_TILE)) {
        if (D == 2) launch_g2p_bins<2>(P, G, dt, s2[cur], n_binned, bin_start, gp<2>(), mig, status_dev, strict, stream);
        else launch_g2p_bins<3>(P, G, dt, s3[cur], n_binned, bin_start, gp<3>(), mig, status_dev, strict, stream);
        if (D == 2) launch_g2p_naive<2>(P, dt, s2[cur], n_binned, n, gp<2>(), mig, status_dev, strict, stream, nullptr, dev_ext);
        else launch_g2p_naive<3>(P, dt, s3[cur], n_binned, n, gp<3>(), mig, status_dev, strict, stream, nullptr, dev_ext);
      } else {
        if (D == 2) launch_g2p_naive<2>(P, dt, s2[cur], 0, n, gp<2>(), mig, status_dev, strict, stream, stats_dev, dev_ext);
        else launch_g2p_naive<3>(P, dt, s3[cur], 0, n, gp<3>(), mig, status_dev, strict, stream, stats_dev, dev_ext);
      }
    }
    if (pipelined) {
      // x-slab handles run every path on the pipelined schedule: the P2G of the NEXT substep follows at once
      // (two kernels instead of the fused one), so that one exchange carries ghost sums and emigrants together
      Phase ph(this, MPM_PHASE_P2G, n > 0 && !p2g_overlapped ? 1 : 0);
      const bool strict_p2g = (cfg.flags & MPM_FLAG_STRICT) != 0;
      if (p2g_overlapped) {
        // launched above, boundary on the main stream and interior on the side stream
      } else if (D == 2) {
        GridPtrs<2> gn = gp<2>();
        gn.g = grid_next;
        if (binned) launch_p2g_cells<2>(P, G, dt, s2[cur], n_binned, bin_start, gn, status_dev, stats_dev, strict_p2g, stream);
        launch_p2g_naive<2>(P, dt, s2[cur], binned ? n_binned : 0, n, gn, status_dev, stream, dev_ext);
      } else {
        GridPtrs<3> gn = gp<3>();
        gn.g = grid_next;
        if (binned) launch_p2g_cells<3>(P, G, dt, s3[cur], n_binned, bin_start, gn, status_dev, stats_dev, strict_p2g, stream);
        launch_p2g_naive<3>(P, dt, s3[cur], binned ? n_binned : 0, n, gn, status_dev, stream, dev_ext);
      }
      grid_read = grid;
      float4 *t = grid;
      grid = grid_next;
      grid_next = t;
      p2g_ready = true;
      p2g_dt = dt;
    } else {
      p2g_ready = false;
    }
  }
  if (prof_on) prof.substeps++;
  return MPM_OK;
}

// MPM_FLAG_DETERMINISTIC: stable sort of the storage by cell, then the fixed-order P2G (mpm_deterministic.cu)
int mpm_handle::det_sort_and_p2g(float dt) {
  if (n > 0) {
    Phase ph(this, MPM_PHASE_BIN, 3 + 3 * ((det_key_bits + 7) / 8));
    if (D == 2) launch_det_cell_keys<2>(P, s2[cur], n, sb.key[0], status_dev, stream);
    else launch_det_cell_keys<3>(P, s3[cur], n, sb.key[0], status_dev, stream);
    launch_iota(sb.val[0], n, stream);
    const int r = radix_sort_pairs(sb, n, det_key_bits, stream);
    launch_bin_starts(sb.key[r], n, (int)det_cells, det_cell_start, stream);
    if (D == 2) launch_reorder<2>(s2[cur], s2[cur ^ 1], sb.val[r], n, stream);
    else launch_reorder<3>(s3[cur], s3[cur ^ 1], sb.val[r], n, stream);
    cur ^= 1;
  } else {
    MPM_CUDA(cudaMemsetAsync(det_cell_start, 0, ((size_t)det_cells + 2) * sizeof(int), stream));
  }
  Phase ph(this, MPM_PHASE_P2G, 2);
  if (D == 2) launch_det_p2g<2>(P, dt, s2[cur], n, det_rec, det_cell_start, grid, nodes, stream);
  else launch_det_p2g<3>(P, dt, s3[cur], n, det_rec, det_cell_start, grid, nodes, stream);
  p2g_ready = true;
  p2g_dt = dt;
  grid_read = grid;
  MPM_CUDA(cudaGetLastError());
  return MPM_OK;
}

int mpm_handle::substep(float dt, int n_steps) {
  if (n_steps < 0) {
    err = "substep: n_steps < 0";
    return MPM_E_INVALID;
  }
  MPM_CUDA(cudaSetDevice(cfg.device));
  if (multi) {
    err = "substep: this handle owns an x-slab; drive it with mpm_step_p2g / halo / grid_g2p / immigrate";
    return MPM_E_STATE;
  }
  if (!(dt > 0)) dt = cfg.dt;
  for (int s = 0; s < n_steps; s++) {
    int rc;
    if (deterministic) {
      if ((rc = det_sort_and_p2g(dt))) return rc;
      if ((rc = step_grid_g2p(dt))) return rc;
      continue;
    }
    const int every = current_interval();
    if (every > 0 && steps_since_sort >= every) {
      // the fast 2D kernel re-sorts on the fly inside this substep; every other path re-sorts stand-alone now
      if (resorts_on_the_fly() && n > 0) resort_due = true;
      else if ((rc = rebin_storage())) return rc;
    }
    if ((rc = step_p2g(dt))) return rc;
    if ((rc = step_grid_g2p(dt))) return rc;
    steps_since_sort++;
  }
  MPM_CUDA(cudaGetLastError());
  return MPM_OK;
}

int mpm_handle::read_grid(int stage_id, float *out) {
  join_side();
  if (!out || (stage_id != 0 && stage_id != 1)) {
    err = "read_grid: bad arguments";
    return MPM_E_INVALID;
  }
  if (stage_id == 1 && (!grid_tap || !tap_valid)) {
    err = "read_grid: stage 1 needs MPM_FLAG_CAPTURE_POST_P2G and at least one substep";
    return MPM_E_STATE;
  }
  MPM_CUDA(cudaSetDevice(cfg.device));
  std::vector<float4> host((size_t)nodes);
  MPM_CUDA(cudaMemcpyAsync(host.data(), stage_id == 1 ? grid_tap : grid_read, (size_t)nodes * sizeof(float4),
                           cudaMemcpyDeviceToHost, stream));
  MPM_CUDA(cudaStreamSynchronize(stream));
  for (long long i = 0; i < nodes; i++) {
    const float4 &g = host[(size_t)i];
    if (D == 2) {
      out[3 * i + 0] = g.x;
      out[3 * i + 1] = g.y;
      out[3 * i + 2] = g.z;
    } else {
      out[4 * i + 0] = g.x;
      out[4 * i + 1] = g.y;
      out[4 * i + 2] = g.z;
      out[4 * i + 3] = g.w;
    }
  }
  return MPM_OK;
}

int mpm_handle::bin_particles(int *cell, int *key, int *order, int *bin_start_out) {
  join_side();
  if (multi || !plain_ids) {
    // outputs are indexed by upload order: only defined when the ids ARE the upload indices of one whole set
    err = "bin_particles: needs a whole-domain handle filled by mpm_upload_particles (ids = upload indices)";
    return MPM_E_STATE;
  }
  MPM_CUDA(cudaSetDevice(cfg.device));
  if (n == 0) return G.n_bins;
  int rc;
  if (cell && !cell_dev)
    if ((rc = dalloc(&cell_dev, (size_t)cap * D))) return rc;
  // keys in UPLOAD order (scatter by id) so that the stable sort reproduces oracle_bin's permutation
  if (D == 2) launch_bin_keys<2>(P, G, s2[cur], n, cell ? cell_dev : nullptr, sb.key[0], status_dev, true, stream);
  else launch_bin_keys<3>(P, G, s3[cur], n, cell ? cell_dev : nullptr, sb.key[0], status_dev, true, stream);
  if (key) MPM_CUDA(cudaMemcpyAsync(key, sb.key[0], (size_t)n * 4, cudaMemcpyDeviceToHost, stream));
  if (cell) MPM_CUDA(cudaMemcpyAsync(cell, cell_dev, (size_t)n * D * 4, cudaMemcpyDeviceToHost, stream));
  launch_iota(sb.val[0], n, stream);
  int r = radix_sort_pairs(sb, n, key_bits, stream);
  // bin starts into the active-bin scan scratch (n_bins + 2 words; bin_start itself belongs to the storage order)
  int *tmp_start = (int *)active_offs;
  if (order) MPM_CUDA(cudaMemcpyAsync(order, sb.val[r], (size_t)n * 4, cudaMemcpyDeviceToHost, stream));
  launch_bin_starts(sb.key[r], n, G.n_bins, tmp_start, stream);
  if (bin_start_out)
    MPM_CUDA(cudaMemcpyAsync(bin_start_out, tmp_start, ((size_t)G.n_bins + 1) * 4, cudaMemcpyDeviceToHost, stream));
  MPM_CUDA(cudaGetLastError());
  MPM_CUDA(cudaStreamSynchronize(stream));
  return G.n_bins;
}

int mpm_handle::poll_status() {
  join_side();
  MPM_CUDA(cudaSetDevice(cfg.device));
  int st = 0;
  MPM_CUDA(cudaMemcpyAsync(&st, status_dev, 4, cudaMemcpyDeviceToHost, stream));
  MPM_CUDA(cudaMemsetAsync(status_dev, 0, 4, stream));
  MPM_CUDA(cudaStreamSynchronize(stream));
  if (st & STATUS_DOMAIN) {
    err = "a particle left the grid (base cell clamped)";
    return MPM_E_DOMAIN;
  }
  if (st & STATUS_MIGRATION_OVERFLOW) {
    err = "more emigrants in one substep than the migration buffers hold";
    return MPM_E_CAPACITY;
  }
  if (st & STATUS_CFL) {
    err = "overlapped slab schedule: a particle of an interior bin reached the slab cut (re-sort interval too long)";
    return MPM_E_CFL;
  }
  return MPM_OK;
}

long long mpm_handle::read_ids(void *aos_out, int *ids_out, long long max_n, int to_device) {
  join_side();
  if (multi) sync_extent();
  if (max_n < n || !aos_out || !ids_out) {
    err = "read_ids: buffers must hold mpm_storage_extent() records";
    return MPM_E_INVALID;
  }
  MPM_CUDA(cudaSetDevice(cfg.device));
  if (n > 0) {
    // storage order, one pass; image + ids in the caller's device buffers or in the idle storage arena
    const int W = record_words();
    const size_t img_bytes = (size_t)n * W * 4, ids_off = (img_bytes + 255) & ~(size_t)255;
    float *img = to_device ? (float *)aos_out : (float *)arena[cur ^ 1];
    int *dev_ids = to_device ? ids_out : (int *)(arena[cur ^ 1] + ids_off);
    if (D == 2) launch_soa_to_aos<2>(s2[cur], n, 0, n, img, dev_ids, stream);
    else launch_soa_to_aos<3>(s3[cur], n, 0, n, img, dev_ids, stream);
    if (!to_device) {
      MPM_CUDA(cudaMemcpyAsync(aos_out, img, img_bytes, cudaMemcpyDeviceToHost, stream));
      MPM_CUDA(cudaMemcpyAsync(ids_out, dev_ids, (size_t)n * 4, cudaMemcpyDeviceToHost, stream));
    }
  }
  MPM_CUDA(cudaGetLastError());
  MPM_CUDA(cudaStreamSynchronize(stream));
  return n;
}

// ------------------------------------------------------------------------------------------------
// x-slab protocol: ONE fixed-size message per neighbour and substep, no host synchronisation.
//
// Invariant between substeps (pipelined schedule): the particles hold the state after G2P(n) and `grid` holds
// this handle's partial sums of P2G(n+1).  A particle that changes slab in G2P(n) is packed into the message for
// that neighbour; its P2G(n+1) share goes to the node columns this handle holds on THIS side (so the two shared
// columns leave complete) and to the remaining columns on the receiver's side.  The message carries the partial
// sums of the two shared node columns, the emigrant count and the records; the receiver adds the sums
// (commutative: both sides end with bit-identical shared columns), appends the records behind its storage extent
// -- which lives on the device, like the counts -- and scatters their remaining P2G share.
//   slab_begin : (re)compute P2G of the resident particles, stage the messages              -> exchange
//   slab_step  : [consume the received messages] grid update, G2P + next P2G, stage         -> exchange
//   slab_settle: consume the received messages (state complete: every particle resident, grid whole)
// ------------------------------------------------------------------------------------------------
int mpm_handle::slab_stage() {
  const bool have_lo = cfg.slab_lo > 0, have_hi = cfg.slab_hi < cfg.n_grid;
  const int K = mig.cap;
  // emigrants of this substep: their share of the NEXT P2G on the node columns this handle holds
  if (p2g_ready) {
    const int c0 = cfg.slab_lo, c1 = cfg.slab_lo + P.ncol;
    if (D == 2) {
      if (have_lo) launch_scatter_records<2>(P, p2g_dt, mig.send_lo, mig.count + 0, K, grid, c0, c1, status_dev, stream);
      if (have_hi) launch_scatter_records<2>(P, p2g_dt, mig.send_hi, mig.count + 1, K, grid, c0, c1, status_dev, stream);
    } else {
      if (have_lo) launch_scatter_records<3>(P, p2g_dt, mig.send_lo, mig.count + 0, K, grid, c0, c1, status_dev, stream);
      if (have_hi) launch_scatter_records<3>(P, p2g_dt, mig.send_hi, mig.count + 1, K, grid, c0, c1, status_dev, stream);
    }
  }
  const long long hn = halo_nodes();
  if (have_lo) {
    MPM_CUDA(cudaMemcpyAsync(msg_send[0], grid, msg_hdr, cudaMemcpyDeviceToDevice, stream));
    MPM_CUDA(cudaMemcpyAsync(hdr(msg_send[0]), mig.count + 0, 4, cudaMemcpyDeviceToDevice, stream));
  }
  if (have_hi) {
    MPM_CUDA(cudaMemcpyAsync(msg_send[1], grid + (nodes - hn), msg_hdr, cudaMemcpyDeviceToDevice, stream));
    MPM_CUDA(cudaMemcpyAsync(hdr(msg_send[1]), mig.count + 1, 4, cudaMemcpyDeviceToDevice, stream));
  }
  // the emigrants are no longer this handle's
  launch_slab_counters(dev_ext, nullptr, nullptr, have_lo ? mig.count + 0 : nullptr, have_hi ? mig.count + 1 : nullptr, K,
                       cap, nullptr, stream);
  MPM_CUDA(cudaGetLastError());
  slab_state = SLAB_STAGED;
  return MPM_OK;
}

int mpm_handle::slab_consume() {
  const bool have_lo = cfg.slab_lo > 0, have_hi = cfg.slab_hi < cfg.n_grid;
  const int K = mig.cap;
  const long long hn = halo_nodes();
  {
    Phase ph(this, MPM_PHASE_HALO, (have_lo ? 1 : 0) + (have_hi ? 1 : 0));
    if (have_lo) launch_halo_add(grid, (const float4 *)msg_recv[0], hn, stream);
    if (have_hi) launch_halo_add(grid + (nodes - hn), (const float4 *)msg_recv[1], hn, stream);
  }
  // room for the arrivals?  `n` is the host's upper bound of the extent (it grows by K per neighbour and substep
  // until a re-sort reads the exact value back); compact before it would pass the capacity
  const long long room = (long long)K * ((have_lo ? 1 : 0) + (have_hi ? 1 : 0));
  if (n + room > cap) {
    int rc = sync_extent();  // the exact extent instead of the bound (small handles get here; big ones re-sort first)
    if (rc) return rc;
    if (n + room > cap && steps_since_sort > 0) {
      rc = rebin_storage();  // drops the slots of emigrated particles
      if (rc) return rc;
    }
  }
  {
    Phase ph(this, MPM_PHASE_MIGRATE, 3);
    const float *r_lo = have_lo ? recs(msg_recv[0]) : nullptr, *r_hi = have_hi ? recs(msg_recv[1]) : nullptr;
    const int *c_lo = have_lo ? hdr(msg_recv[0]) : nullptr, *c_hi = have_hi ? hdr(msg_recv[1]) : nullptr;
    // arrivals from below were scattered into the shared columns lo, lo+1 by their sender; from above into hi, hi+1
    const int lo = cfg.slab_lo, top = cfg.slab_lo + P.ncol;
    if (D == 2) {
      launch_immigrate<2>(r_lo, c_lo, r_hi, c_hi, K, s2[cur], dev_ext, cap, status_dev, stream);
      if (have_lo) launch_scatter_records<2>(P, p2g_dt, r_lo, c_lo, K, grid, lo + 2, top, status_dev, stream);
      if (have_hi) launch_scatter_records<2>(P, p2g_dt, r_hi, c_hi, K, grid, lo, top - 2, status_dev, stream);
    } else {
      launch_immigrate<3>(r_lo, c_lo, r_hi, c_hi, K, s3[cur], dev_ext, cap, status_dev, stream);
      if (have_lo) launch_scatter_records<3>(P, p2g_dt, r_lo, c_lo, K, grid, lo + 2, top, status_dev, stream);
      if (have_hi) launch_scatter_records<3>(P, p2g_dt, r_hi, c_hi, K, grid, lo, top - 2, status_dev, stream);
    }
    launch_slab_counters(dev_ext, c_lo, c_hi, nullptr, nullptr, K, cap, nullptr, stream);
  }
  n = n + room < cap ? n + room : cap;
  MPM_CUDA(cudaGetLastError());
  return MPM_OK;
}

int mpm_handle::slab_begin(float dt) {
  if (!multi) return 0;
  MPM_CUDA(cudaSetDevice(cfg.device));
  // a run that continues with the same dt has nothing to (re)compute: either the state is complete (SETTLED) or the
  // messages of the last substep were staged and -- by the contract "exchange after every call" -- exchanged, and the
  // next mpm_slab_step consumes them
  if (p2g_ready && p2g_dt == dt && slab_state != SLAB_FRESH) return 0;
  int rc;
  if (slab_state == SLAB_STAGED) {  // dt changes mid-run: take in the outstanding messages, then redo the P2G
    if ((rc = slab_consume())) return rc;
    slab_state = SLAB_SETTLED;
  }
  rc = resident_p2g(dt);
  if (rc) return rc;
  MPM_CUDA(cudaMemsetAsync(mig.count, 0, 8, stream));  // nobody emigrates in a P2G
  if ((rc = slab_stage())) return rc;
  return 1;
}

int mpm_handle::slab_step(float dt) {
  if (!multi) {
    err = "slab_step: this handle owns the whole domain (use mpm_substep)";
    return MPM_E_STATE;
  }
  MPM_CUDA(cudaSetDevice(cfg.device));
  if (slab_state == SLAB_FRESH || !p2g_ready || p2g_dt != dt) {
    err = "slab_step: call mpm_slab_begin first (after an upload or a change of dt) and exchange its messages";
    return MPM_E_STATE;
  }
  int rc;
  if (slab_state == SLAB_STAGED)
    if ((rc = slab_consume())) return rc;
  const int every = current_interval();
  if (every > 0 && steps_since_sort >= every) resort_due = true;  // consumed inside step_grid_g2p
  if ((rc = step_grid_g2p(dt))) return rc;
  steps_since_sort++;
  return slab_stage();
}

int mpm_handle::slab_settle() {
  if (!multi || slab_state != SLAB_STAGED) return MPM_OK;
  MPM_CUDA(cudaSetDevice(cfg.device));
  int rc = slab_consume();
  if (rc) return rc;
  slab_state = SLAB_SETTLED;
  return MPM_OK;
}

// exact storage extent / live count of an x-slab handle (they live on the device): synchronises
int mpm_handle::sync_extent() {
  if (!multi) return MPM_OK;
  MPM_CUDA(cudaSetDevice(cfg.device));
  join_side();
  int e2[2] = {0, 0};
  MPM_CUDA(cudaMemcpyAsync(e2, dev_ext, 8, cudaMemcpyDeviceToHost, stream));
  MPM_CUDA(cudaStreamSynchronize(stream));
  n = e2[0];
  live = e2[1];
  return MPM_OK;
}

// ================================================================================================
// C-ABI
// ================================================================================================
extern "C" {

int mpm_default_config(mpm_config *c, int dim) {
  if (!c || (dim != 2 && dim != 3)) return MPM_E_INVALID;
  memset(c, 0, sizeof *c);
  c->abi_version = MPM_ABI_VERSION;
  c->dim = dim;
  c->n_grid = 80;      // :9
  c->dt = 1e-4f;       // :11
  c->mass_p = 1.0f;    // :17
  c->vol_p = 1.0f;     // :18
  c->gravity[0] = 0.0f;
  c->gravity[1] = -200.0f;  // :113
  c->gravity[2] = 0.0f;
  c->boundary = 0.05f;  // :116
  c->jp_min = 0.6f;     // :175
  c->jp_max = 20.0f;
  c->alpha = 0.0f;
  c->n_materials = 4;
  const float lo = 1.0f - 2.5e-2f, hi = 1.0f + 7.5e-3f;  // :169
  mpm_material fluid = {MPM_KIND_FLUID, 1e4f, 0.2f, 0.0f, lo, hi};
  mpm_material jelly = {MPM_KIND_JELLY, 1e4f, 0.2f, 0.3f, lo, hi};
  mpm_material snow = {MPM_KIND_SNOW, 1e4f, 0.2f, 10.0f, lo, hi};
  mpm_material shipped = {MPM_KIND_SNOW, 1e2f, 0.499f, 1.0f, lo, hi};  // :18-20
  c->materials[0] = fluid;
  c->materials[1] = jelly;
  c->materials[2] = snow;
  c->materials[3] = shipped;
  c->capacity = 1 << 20;
  c->device = 0;
  c->flags = 0;
  c->slab_lo = 0;
  c->slab_hi = c->n_grid;
  return MPM_OK;
}

int mpm_config_bytes(void) { return (int)sizeof(mpm_config); }

mpm_handle *mpm_create(const mpm_config *cfg) {
  if (!cfg) {
    g_create_error = "mpm_create: NULL config";
    return nullptr;
  }
  std::string why;
  if (validate(*cfg, why) != MPM_OK) {
    g_create_error = "mpm_create: " + why;
    return nullptr;
  }
  mpm_handle *h = new (std::nothrow) mpm_handle();
  if (!h) {
    g_create_error = "mpm_create: out of host memory";
    return nullptr;
  }
  h->cfg = *cfg;
  int rc = h->init();
  if (rc != MPM_OK) {
    g_create_error = "mpm_create: " + h->err;
    delete h;
    return nullptr;
  }
  return h;
}

void mpm_destroy(mpm_handle *h) { delete h; }

const char *mpm_last_error(const mpm_handle *h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int mpm_upload_particles(mpm_handle *h, const void *aos, long long n, int on_device) {
  return h ? h->upload(aos, nullptr, n, on_device) : MPM_E_INVALID;
}
int mpm_upload_particles_ids(mpm_handle *h, const void *aos, const int *ids, long long n, int on_device) {
  return h ? h->upload(aos, ids, n, on_device) : MPM_E_INVALID;
}
long long mpm_read_particles_ids(mpm_handle *h, void *aos_out, int *ids_out, long long max_n, int to_device) {
  return h ? h->read_ids(aos_out, ids_out, max_n, to_device) : MPM_E_INVALID;
}
long long mpm_storage_extent(mpm_handle *h) {
  if (!h) return -1;
  if (h->sync_extent() != MPM_OK) return MPM_E_CUDA;
  return h->n;
}
int mpm_substep(mpm_handle *h, float dt, int n_steps) { return h ? h->substep(dt, n_steps) : MPM_E_INVALID; }
int mpm_read_particles(mpm_handle *h, void *aos_out, long long n, int to_device) {
  return h ? h->read(aos_out, n, to_device) : MPM_E_INVALID;
}
int mpm_read_grid(mpm_handle *h, int stage, float *out) { return h ? h->read_grid(stage, out) : MPM_E_INVALID; }
long long mpm_particle_count(mpm_handle *h) {
  if (!h) return -1;
  if (h->sync_extent() != MPM_OK) return MPM_E_CUDA;
  return h->live;
}
int mpm_resort(mpm_handle *h) {
  if (!h) return MPM_E_INVALID;
  cudaSetDevice(h->cfg.device);
  return h->rebin_storage();
}
int mpm_set_rebin_every(mpm_handle *h, int every) {
  if (!h) return MPM_E_INVALID;
  h->cfg.rebin_every = every;
  return MPM_OK;
}
int mpm_synchronize(mpm_handle *h) {
  if (!h) return MPM_E_INVALID;
  cudaSetDevice(h->cfg.device);
  h->join_side();
  cudaError_t e = cudaStreamSynchronize(h->stream);
  if (e != cudaSuccess) {
    h->err = std::string("synchronize: ") + cudaGetErrorString(e);
    return MPM_E_CUDA;
  }
  return MPM_OK;
}
int mpm_poll_status(mpm_handle *h) { return h ? h->poll_status() : MPM_E_INVALID; }
int mpm_profile_enable(mpm_handle *h, int on) {
  if (!h) return MPM_E_INVALID;
  cudaSetDevice(h->cfg.device);
  h->join_side();
  cudaStreamSynchronize(h->stream);
  h->flush_spans();
  memset(&h->prof, 0, sizeof h->prof);
  cudaMemsetAsync(h->stats_dev, 0, 32, h->stream);
  h->prof_on = on != 0;
  return MPM_OK;
}
int mpm_profile_read(mpm_handle *h, mpm_profile *out) {
  if (!h || !out) return MPM_E_INVALID;
  cudaSetDevice(h->cfg.device);
  h->join_side();
  cudaError_t e = cudaStreamSynchronize(h->stream);
  if (e != cudaSuccess) {
    h->err = std::string("profile_read: ") + cudaGetErrorString(e);
    return MPM_E_CUDA;
  }
  h->flush_spans();
  unsigned long long st[4] = {0, 0, 0, 0};
  cudaMemcpy(st, h->stats_dev, sizeof st, cudaMemcpyDeviceToHost);
  h->prof.fallback_particles = (long long)st[0];
  h->prof.rebin_interval = h->current_interval();
  *out = h->prof;
  return MPM_OK;
}
int mpm_bin_particles(mpm_handle *h, int *cell, int *key, int *order, int *bin_start) {
  return h ? h->bin_particles(cell, key, order, bin_start) : MPM_E_INVALID;
}

// ---- x-slab protocol ------------------------------------------------------------------------------
int mpm_slab_describe(mpm_handle *h, mpm_slab_desc *d) {
  if (!h || !d) return MPM_E_INVALID;
  memset(d, 0, sizeof *d);
  if (!h->multi) return MPM_OK;
  d->send_lo = h->msg_send[0];
  d->send_hi = h->msg_send[1];
  d->recv_lo = h->msg_recv[0];
  d->recv_hi = h->msg_recv[1];
  d->bytes = (long long)h->msg_bytes;
  d->halo_bytes = (long long)h->msg_hdr;
  d->record_bytes = h->mig_words() * 4;
  d->record_capacity = h->mig.cap;
  d->has_lo = h->cfg.slab_lo > 0;
  d->has_hi = h->cfg.slab_hi < h->cfg.n_grid;
  return MPM_OK;
}
int mpm_slab_begin(mpm_handle *h, float dt) {
  if (!h) return MPM_E_INVALID;
  if (!(dt > 0)) dt = h->cfg.dt;
  return h->slab_begin(dt);
}
int mpm_slab_step(mpm_handle *h, float dt) {
  if (!h) return MPM_E_INVALID;
  if (!(dt > 0)) dt = h->cfg.dt;
  return h->slab_step(dt);
}
int mpm_slab_settle(mpm_handle *h) { return h ? h->slab_settle() : MPM_E_INVALID; }
}
