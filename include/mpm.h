/* include/mpm.h -- C-ABI of the B200-native MLS-MPM substep engine (libmpm.so).
 *
 * Drop-in boundary for the `advance(dt)` path of the reference program
 *   /root/reference/cpp_validation/mls-mpm88-explained.cpp:49-180
 * whose state lives in two globals, `std::vector<Particle> particles` (:44) and
 * `Vector3 grid[num_grid+1][num_grid+1]` (:47), and whose parameters are file-scope consts (:8-26).
 * The reference has no FFI of its own; each entry point below names the reference statement(s)
 * it replaces.  Plain C types only: pointers, sizes, PODs.  No exceptions or signals cross this
 * boundary; every call returns 0 (MPM_OK) or a negative MPM_E_* code, and mpm_last_error() gives
 * the text.  A handle is not thread-safe; distinct handles are independent.
 *
 * Particle records cross the boundary in the reference's own memory layout (:28-42):
 *   2D, 56 bytes : x[2] v[2] F[4] C[4] Jp c      (F, C column-major like taichi.h:7575)
 *   3D, 104 bytes: x[3] v[3] F[9] C[9] Jp c      (the 3D lift of the same struct)
 * `c` (the reference's colour slot, :33) is the material id: an index into mpm_config.materials,
 * any other value selects the LAST table entry (so the shipped scene's colour 0x2986CC selects the
 * shipped constants when they are the last entry, which is the default table).
 */
#ifndef MPM_FLIP98A_B200_MPM_H
#define MPM_FLIP98A_B200_MPM_H

#ifdef __cplusplus
extern "C" {
#endif

#define MPM_ABI_VERSION 1

enum {
  MPM_OK = 0,
  MPM_E_INVALID = -1,   /* bad argument / config */
  MPM_E_CUDA = -2,      /* CUDA runtime error (text in mpm_last_error) */
  MPM_E_CAPACITY = -3,  /* more particles than the handle was created for */
  MPM_E_DOMAIN = -4,    /* a particle left the grid (the reference has no bounds check, :97,:150) */
  MPM_E_CFL = -5,       /* MPM_FLAG_OVERLAP only: a particle of an interior bin reached the slab cut */
  MPM_E_STATE = -6      /* call sequence error (e.g. read before upload) */
};

enum { MPM_KIND_FLUID = 0, MPM_KIND_JELLY = 1, MPM_KIND_SNOW = 2 };

/* Constitutive model of one material id.  snow == the reference's shipped path
 * (hardening :67-69, SVD clamp :165-178); fluid / jelly follow north_star. */
typedef struct mpm_material {
  int kind;        /* MPM_KIND_* */
  float E;         /* Young's modulus        (:19) */
  float nu;        /* Poisson ratio          (:20) */
  float hardening; /* snow: exp factor (:18); jelly: constant factor on mu/lambda; fluid: unused */
  float sig_lo;    /* plastic clamp of singular values, reference 1-2.5e-2 (:169) */
  float sig_hi;    /* reference 1+7.5e-3 (:169) */
} mpm_material;

enum {
  MPM_FLAG_CAPTURE_POST_P2G = 1 << 0, /* keep a copy of the grid between P2G and the grid update */
  MPM_FLAG_NAIVE = 1 << 1,            /* one-thread-per-particle kernels with global atomics
                                         (no binning); the small-scene / debugging path */
  MPM_FLAG_G2P_TILE = 1 << 3,         /* binned path: G2P stages each bin's node tile in shared memory
                                         (measured no faster than the read-only-path gather: off by default) */
  MPM_FLAG_NO_FUSE = 1 << 4,          /* keep P2G and G2P as separate kernels (2D default, single GPU and
                                         x-slabs alike: G2P of a substep and P2G of the next run as ONE
                                         kernel, one particle read + one write per substep; 3D runs two kernels
                                         unless MPM_FLAG_FUSE_3D) */
  MPM_FLAG_OVERLAP = 1 << 5,          /* x-slab handles (2D substep kernel; 3D two-kernel and fused schedules): bins >= 2
                                         bin columns away from the slab cuts run on a side stream while the boundary
                                         bins finish first, so the caller's migration / halo exchange overlaps the
                                         interior compute; guarded per substep (MPM_E_CFL) */
  MPM_FLAG_DETERMINISTIC = 1 << 6,    /* fixed summation order: the storage is stably sorted by cell before every substep
                                         and P2G runs one thread per grid node, adding the contributions of its 3^d
                                         cells' particles sequentially in storage order -- the order the reference's
                                         serial loop (:53-102) would use on that array.  Two runs are bit-identical and
                                         the grid after P2G (total mass included) is BITWISE the CPU oracle's on the
                                         same particle order.  Exact association throughout; whole-domain handles
                                         only; a validation mode, ~10x slower. */
  MPM_FLAG_FUSE_3D = 1 << 7,          /* 3D: G2P of a substep and P2G of the next as ONE kernel (k_substep3d: the snow
                                         projection hands its rotation factor to the stress).  Correct and tested, but
                                         measured no faster on B200 than the two 3D kernels (3.34 vs 3.23 ms on the 256^3
                                         scene: 96 registers leave 5 CTAs per SM): opt-in */
  MPM_FLAG_STRICT = 1 << 2            /* binned path: P2G node contributions (:92-100) and the G2P gather
                                         (:153-154) keep the reference's exact association instead of the
                                         separable / hoisted FMA forms (algebraically identical, ~1e-7
                                         relative apart).  MPM_FLAG_NAIVE is always exact. */
};

/* Everything the reference fixes at compile time (:8-26, :113, :116, :169, :175) plus the
 * engine's own sizing.  Fill with mpm_default_config() first, then override. */
typedef struct mpm_config {
  int abi_version;   /* MPM_ABI_VERSION */
  int dim;           /* 2 (reference) or 3 */
  int n_grid;        /* cells per axis (:9 num_grid); nodes per axis = n_grid+1 */
  float dt;          /* default substep when mpm_substep gets dt <= 0 (:11) */
  float mass_p;      /* :17 */
  float vol_p;       /* :18 */
  float gravity[3];  /* :113 hard-codes (0,-200) */
  float boundary;    /* :116 hard-codes 0.05 */
  float jp_min;      /* :175 hard-codes 0.6 */
  float jp_max;      /* :175 hard-codes 20 */
  float alpha;       /* FLIP/PIC blend; 0 == the reference's pure APIC (config.py:29) */
  int n_materials;   /* 1..4 */
  mpm_material materials[4];
  long long capacity; /* max particles resident on this handle */
  int device;         /* CUDA device ordinal */
  int flags;          /* MPM_FLAG_* */
  /* x-slab owned by this handle: base-cell columns [slab_lo, slab_hi) of the global grid.
   * 0 / n_grid == the whole domain (single GPU).  See the halo / migration calls below. */
  int slab_lo, slab_hi;
  void *stream;       /* cudaStream_t to launch on; NULL = a stream owned by the handle */
  int bin_edge;       /* cells per bin edge for the block binning; 0 = engine default */
  int rebin_every;    /* re-sort the particle storage by bin every this many substeps; <0 = never;
                         0 = engine default: 32 on the naive path, adaptive 2..512 on the binned path
                         (0.75 cells / the largest per-substep displacement the kernels measure) */
  int mig_records;    /* x-slab handles: emigrant records per message (must be the same on every handle of a
                         decomposition); 0 = engine default: 2 x the nodes of a cut (2D) / half the nodes of a cut (3D),
                         clamped to [4096, 262144] */
  int reserved[5];
} mpm_config;

typedef struct mpm_handle mpm_handle;

/* Defaults = the reference as shipped (2D: n_grid 80, dt 1e-4, E 1e2, nu 0.499, hardening 1 as
 * the LAST material; entries 0..2 = fluid / jelly / snow with the upstream mls-mpm88 constants). */
int mpm_default_config(mpm_config *cfg, int dim);
/* sizeof(mpm_config) as the library was compiled: lets an FFI binding check its struct layout */
int mpm_config_bytes(void);

/* replaces: the file-scope state and constants, :8-26, :44-47 */
mpm_handle *mpm_create(const mpm_config *cfg);
void mpm_destroy(mpm_handle *h);
const char *mpm_last_error(const mpm_handle *h); /* h may be NULL: error of the last failed create */

/* replaces: particles.push_back(...) in add_object, :191-196.  `aos` = n records of 56 B (2D) or
 * 104 B (3D) in HOST memory (or device memory if on_device != 0).  Replaces the particle set. */
int mpm_upload_particles(mpm_handle *h, const void *aos, long long n, int on_device);

/* replaces: `for (...) advance(dt)` in main(), :214-215.  Asynchronous on the handle's stream.
 * dt <= 0 uses cfg.dt. */
int mpm_substep(mpm_handle *h, float dt, int n_steps);

/* replaces: reading the global `particles` (:220-222).  Synchronises; records come back in the
 * ORIGINAL upload order (the engine carries a persistent id through its sorts and migration). */
int mpm_read_particles(mpm_handle *h, void *aos_out, long long n, int to_device);

/* replaces: reading the global `grid` (:47).  stage 0 = after the grid update of the last substep
 * ((vx,vy,1|0) per node 2D, (vx,vy,vz,1|0) 3D, exactly the reference's in-memory content);
 * stage 1 = between P2G and the grid update ((m*vx, m*vy, m) -- needs MPM_FLAG_CAPTURE_POST_P2G).
 * out = (slab columns) * (n_grid+1)^(dim-1) * (dim+1) floats, [i][j]([k]) row-major like :47. */
int mpm_read_grid(mpm_handle *h, int stage, float *out);

/* Same as the two calls above with caller-chosen persistent ids (x-slab runs: each handle holds a
 * subset of a global particle set).  read: records come back in STORAGE order, ids_out[i] = the id of
 * record i or -1 for a slot whose particle emigrated; buffers must hold mpm_storage_extent() records;
 * returns the number of records written. */
int mpm_upload_particles_ids(mpm_handle *h, const void *aos, const int *ids, long long n, int on_device);
long long mpm_read_particles_ids(mpm_handle *h, void *aos_out, int *ids_out, long long max_n, int to_device);
long long mpm_storage_extent(mpm_handle *h);  /* x-slab handles: synchronises (the extent lives on the device) */
long long mpm_particle_count(mpm_handle *h);
/* Re-sorts the particle storage by bin now (the engine does it on its own every rebin interval). */
int mpm_resort(mpm_handle *h);
/* Changes mpm_config.rebin_every of a live handle (0 = back to the adaptive interval). */
int mpm_set_rebin_every(mpm_handle *h, int every);
int mpm_synchronize(mpm_handle *h);
/* Sticky device-side status (MPM_E_DOMAIN, MPM_E_CAPACITY for a migration-buffer overflow) accumulated since
 * the last call; synchronises. */
int mpm_poll_status(mpm_handle *h);

/* ---- per-phase device timing (CUDA events on the handle's stream) and launch counts ---------- */
enum {
  MPM_PHASE_CLEAR = 0, /* grid reset, :50 */
  MPM_PHASE_P2G = 1,   /* :53-102 */
  MPM_PHASE_GRID = 2,  /* :105-131 */
  MPM_PHASE_G2P = 3,   /* :134-179 */
  MPM_PHASE_BIN = 4,   /* binning / sort (no counterpart in the reference) */
  MPM_PHASE_HALO = 5,
  MPM_PHASE_MIGRATE = 6, /* immigrant unpack; with MPM_FLAG_OVERLAP also the boundary bins' fused kernel */
  MPM_PHASE_COUNT = 8
};
typedef struct mpm_profile {
  double ms[MPM_PHASE_COUNT];          /* device milliseconds per phase since mpm_profile_enable */
  long long launches[MPM_PHASE_COUNT]; /* kernels launched per phase (memsets/copies not counted) */
  long long substeps;
  long long fallback_particles; /* binned P2G: particles that had drifted past the bin margin and
                                   took the per-particle scatter instead (correct, just slower) */
  long long rebin_interval;     /* substeps between storage re-sorts right now */
  long long fused_substeps;     /* substeps whose G2P ran fused with the next P2G (its time is under
                                   MPM_PHASE_G2P; MPM_PHASE_P2G then only holds stand-alone P2G launches) */
} mpm_profile;
int mpm_profile_enable(mpm_handle *h, int on); /* (re)starts accumulation from zero */
int mpm_profile_read(mpm_handle *h, mpm_profile *out); /* synchronises */

/* ---- binning: integer outputs, bit-exact against oracle_bin ------------- */
/* Bins the CURRENT particle positions by blocks of cfg.bin_edge cells (x-major block id of the
 * base cell, :55) with a stable sort.  Outputs (host pointers, any may be NULL):
 *   cell[n*dim] base coordinates clamped to [0,n_grid-2], key[n] bin id, order[n] slot->upload
 *   index, bin_start[n_bins+1].  Returns the number of bins or a negative error. */
int mpm_bin_particles(mpm_handle *h, int *cell, int *key, int *order, int *bin_start);

/* ---- x-slab multi-GPU protocol: one handle per GPU owns the base-cell columns [slab_lo, slab_hi) -------------
 * ONE fixed-size message per neighbour and substep, and no host synchronisation anywhere: the caller only moves
 * `bytes` from send_hi to the upper neighbour's recv_lo and from send_lo to the lower neighbour's recv_hi
 * (ncclSend/ncclRecv, torch.distributed P2P, cudaMemcpyPeerAsync ... stream-ordered with the handle's stream).
 * A message = [partial sums of the 2 node columns shared with that neighbour | 16-byte header: emigrant count |
 * up to record_capacity emigrant records (the reference's Particle record + persistent id)].  Counts, the storage
 * extent and the live-particle count stay on the device.
 *   rc = mpm_slab_begin(h, dt)   after an upload or a change of dt: P2G of the resident particles; returns 1 when
 *                                messages were staged (exchange them now), 0 when there is nothing to exchange
 *                                (in particular when the run simply continues with the same dt)
 *   mpm_slab_step(h, dt)         consumes the received messages (ghost-column sums -- commutative, so both sides end
 *                                with bit-identical shared columns; immigrants appended, their P2G share added),
 *                                then one substep: grid update (shared columns redundantly) and G2P + the NEXT
 *                                substep's P2G; particles whose new base cell left the slab are packed; stages the
 *                                next messages.  Exchange after every call.
 *   mpm_slab_settle(h)           consumes the last messages: every particle resident, grid complete (call it before
 *                                reading particles back; mpm_slab_step continues from there without an exchange)
 * MPM_FLAG_OVERLAP: the bins next to the cuts are computed and staged first, the interior follows on a side stream
 * while the caller's exchange is in flight. */
typedef struct mpm_slab_desc {
  void *send_lo, *send_hi; /* device: message for the lower / upper neighbour */
  void *recv_lo, *recv_hi; /* device: where the lower / upper neighbour's message must land */
  long long bytes;         /* bytes per message (fixed) */
  long long halo_bytes;    /* of which ghost-column sums (the header follows them) */
  int record_bytes;        /* 56+8 (2D) or 104+8 (3D): record + int32 id + pad */
  int record_capacity;     /* emigrants per message; more in one substep raise MPM_E_CAPACITY in mpm_poll_status */
  int has_lo, has_hi;      /* neighbours that exist */
} mpm_slab_desc;
int mpm_slab_describe(mpm_handle *h, mpm_slab_desc *d);
int mpm_slab_begin(mpm_handle *h, float dt);
int mpm_slab_step(mpm_handle *h, float dt);
int mpm_slab_settle(mpm_handle *h);

/* ---- several GPUs behind one handle (mpm_group.cu): the x-slab protocol above driven from the calling thread, one
 * slab per entry of `devices` (a device may appear more than once), messages pulled over peer memory
 * (cudaMemcpyPeerAsync: NVLink where peer access exists), no host synchronisation per substep.  `cfg` describes
 * the WHOLE domain (slab_lo/slab_hi/device/stream are ignored; capacity = all particles).
 * replaces: the same statements of main() as the single-GPU calls -- add_object (:191-196), the advance loop
 * (:214-215), reading `particles` (:220-222). */
typedef struct mpm_group mpm_group;
mpm_group *mpm_group_create(const mpm_config *cfg, const int *devices, int n_devices);
void mpm_group_destroy(mpm_group *g);
const char *mpm_group_last_error(const mpm_group *g);
/* cuts the grid into x-slabs of (nearly) equal particle counts and uploads each slab's particles (host records) */
int mpm_group_upload_particles(mpm_group *g, const void *aos, long long n);
int mpm_group_substep(mpm_group *g, float dt, int n_steps); /* asynchronous */
int mpm_group_synchronize(mpm_group *g);
int mpm_group_read_particles(mpm_group *g, void *aos_out, long long n); /* upload order; synchronises */
int mpm_group_poll_status(mpm_group *g);
/* slab k: its device, owned base-cell columns and current particle count (any pointer may be NULL) */
int mpm_group_slab(const mpm_group *g, int k, int *device, int *slab_lo, int *slab_hi, long long *particles);

#ifdef __cplusplus
}
#endif
#endif
