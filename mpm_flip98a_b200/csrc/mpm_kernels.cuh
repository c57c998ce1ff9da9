// mpm_kernels.cuh -- host-callable launchers of the substep kernels (defined in mpm_kernels.cu,
// mpm_sort.cu).  Internal to libmpm.so; the public boundary is include/mpm.h.
#pragma once
#include "mpm_common.cuh"

namespace mpm {

struct BinGeom {
  int edge;      // cells per bin edge
  int cpb;       // cells per bin = edge^dim
  int nb[3];     // bins per axis (x: over the owned slab)
  int n_bins;
  // bins that held particles at the last re-sort, compacted: the CTA-per-bin kernels launch one CTA per
  // ACTIVE bin (an empty CTA costs ~0.6 ns x 1M bins = 0.66 ms per launch on an 8192^2 grid).  NULL = all bins.
  const int *active;
  int n_active;
};
template <int D>
struct GridPtrs {
  float4 *g;     // node array (see mpm_common.cuh)
  void *vold;    // pre-gravity node velocity: float2 (2D) / float4 (3D); NULL when alpha == 0
  long long nodes;
};

// ---- naive path: one thread per particle, vector REDs straight into L2 -------------------------
template <int D>
void launch_p2g_naive(const Params &P, float dt, const SoA<D> &s, long long first, long long n, GridPtrs<D> g,
                      int *status, cudaStream_t st, const int *dev_n = nullptr);
template <int D>
void launch_g2p_naive(const Params &P, float dt, const SoA<D> &s, long long first, long long n, GridPtrs<D> g,
                      MigPtrs mig, int *status, bool strict, cudaStream_t st, unsigned long long *stats = nullptr,
                      const int *dev_n = nullptr);
// x-slab exchange helpers
void launch_halo_add(float4 *dst, const float4 *src, long long count, cudaStream_t st);
// message-driven immigration and the restricted P2G share of migrating particles (see mpm_kernels.cu)
template <int D>
void launch_immigrate(const float *recv_lo, const int *cnt_lo, const float *recv_hi, const int *cnt_hi, int K,
                      const SoA<D> &s, const int *ext, long long cap, int *status, cudaStream_t st);
template <int D>
void launch_scatter_records(const Params &P, float dt, const float *recs, const int *cnt, int K, float4 *grid, int col_lo,
                            int col_hi, int *status, cudaStream_t st);
void launch_slab_counters(int *ext, const int *add_a, const int *add_b, const int *sub_a, const int *sub_b, int K,
                          long long cap, const int *set_extent, cudaStream_t st);
template <int D>
void launch_grid_update(const Params &P, float dt, GridPtrs<D> g, cudaStream_t st);
// 2D: update `upd` and reset `clr` on the tiles (8x8 nodes) their byte maps mark; see k_grid_tiles
// guard_steps > 0 (overlapped slab schedule): flags STATUS_CFL when (guard_steps + 2) x the largest per-substep
// displacement in stats[2] exceeds 8 cells -- the condition under which the interior launch may skip the migration code
void launch_grid_tiles(const Params &P, float dt, float4 *upd, void *vold, const unsigned char *t_upd, float4 *clr,
                       unsigned char *t_clr, int tiles_x, int tiles_y, cudaStream_t st,
                       const unsigned long long *stats = nullptr, int guard_steps = 0, int *status = nullptr);

// ---- binned path: one CTA per bin, in-CTA cell sort + register accumulation (see mpm_kernels.cu) --
template <int D>
bool p2g_cells_supported(const BinGeom &G);
template <int D>
void launch_p2g_cells(const Params &P, const BinGeom &G, float dt, const SoA<D> &s, long long n, const int *bin_start,
                      GridPtrs<D> g, int *status, unsigned long long *stats, bool strict, cudaStream_t st);

template <int D>
void launch_g2p2g(const Params &P, const BinGeom &G, float dt_g2p, float dt_p2g, const SoA<D> &s, long long n,
                  const int *bin_start, GridPtrs<D> g_in, float4 *grid_out, int *status, unsigned long long *stats,
                  MigPtrs mig, bool strict, cudaStream_t st);
template <int D>
void launch_g2p_bins(const Params &P, const BinGeom &G, float dt, const SoA<D> &s, long long n, const int *bin_start,
                     GridPtrs<D> g, MigPtrs mig, int *status, bool strict, cudaStream_t st);

// ---- the 2D default substep kernel (mpm_substep2d.cu): fused G2P -> P2G, optional on-the-fly re-sort ----
struct Substep2dArgs {
  Params P;
  BinGeom G;
  float dt_g2p, dt_p2g;
  SoA<2> s;                   // particle storage (updated in place unless RESORT)
  SoA<2> d;                   // RESORT: the other storage buffer
  const int4 *chunks;         // work list: (bin, first slot, particles, bin x << 16 | bin y), one CTA each
  int n_chunks;
  const int *new_start;       // RESORT: bin starts of the new order
  const unsigned *key;        // RESORT: new bin of slot i (k_count_rank, positions before this substep)
  const unsigned *rank;       // RESORT: rank of slot i inside its new bin
  const float4 *grid_in;      // updated grid of this substep
  const float2 *vold_in;      // FLIP: pre-gravity node velocity
  float4 *grid_out;           // P2G target of the next substep (zeroed)
  unsigned char *touched_out; // its tile map (8x8-node tiles, x-major, tiles_y per column): marked by every CTA
  int tiles_x, tiles_y;
  int *status;
  unsigned long long *stats;
  MigPtrs mig;
};
void launch_substep2d(const Substep2dArgs &a, bool flip, bool mig, bool resort, cudaStream_t st);
int substep2d_chunk_capacity();  // particles per work-list entry (the kernel's shared-memory chunk)

// ---- the 3D default G2P kernel (mpm_substep3d.cu): early stores, optional on-the-fly re-sort ----
struct G2p3Args {
  Params P;
  float dt;
  SoA<3> s;                   // particle storage (updated in place unless RESORT)
  SoA<3> d;                   // RESORT: the other storage buffer
  long long first, n;         // storage slots [first, n)
  const float4 *grid;         // updated grid of this substep
  const float4 *vold;         // FLIP: pre-gravity node velocity
  const int *new_start;       // RESORT: cell starts of the new order
  const unsigned *key;        // RESORT: new cell of slot i (k_count_rank, positions before this substep)
  const unsigned *rank;       // RESORT: rank of slot i inside its new cell
  MigPtrs mig;
  int *status;
  unsigned long long *stats;
  const int *dev_n;           // x-slab handles: exact storage extent on the device
  const int4 *chunks;         // launch_g2p3_tile: work list (bin, first slot, particles, bin x << 20 | y << 10 | z)
  int n_chunks;
};
void launch_g2p3(const G2p3Args &a, bool flip, bool mig, bool resort, cudaStream_t st);       // thread per particle
void launch_g2p3_tile(const G2p3Args &a, bool flip, bool mig, bool resort, cudaStream_t st);  // CTA per chunk, smem tile

// ---- the fused 3D substep kernel (mpm_substep3d.cu): G2P -> P2G in one pass, optional on-the-fly re-sort ----
struct Substep3dArgs {
  Params P;
  float dt_g2p, dt_p2g;
  SoA<3> s;                   // particle storage (updated in place unless RESORT)
  SoA<3> d;                   // RESORT: the other storage buffer
  const int4 *chunks;         // work list: (bin, first slot, particles, bin x << 20 | bin y << 10 | bin z), one CTA each
  int n_chunks;
  const int *new_start;       // RESORT: cell starts of the new order
  const unsigned *key;        // RESORT: new cell of slot i (k_count_rank, positions before this substep)
  const unsigned *rank;       // RESORT: rank of slot i inside its new cell
  const float4 *grid_in;      // updated grid of this substep
  const float4 *vold_in;      // FLIP: pre-gravity node velocity
  float4 *grid_out;           // P2G target of the next substep (zeroed)
  int *status;
  unsigned long long *stats;
  MigPtrs mig;
};
void launch_substep3d(const Substep3dArgs &a, bool flip, bool mig, bool resort, cudaStream_t st);
int substep3d_chunk_capacity();

// ---- MPM_FLAG_DETERMINISTIC (mpm_deterministic.cu): fixed-order P2G without atomics ----
template <int D>
void launch_det_cell_keys(const Params &P, const SoA<D> &s, long long n, unsigned *key, int *status, cudaStream_t st);
// records (8 | 16 floats per slot) + one thread per node summing its 3^d cells' particles in storage order
template <int D>
void launch_det_p2g(const Params &P, float dt, const SoA<D> &s, long long n, float *rec, const int *cell_start,
                    float4 *grid, long long nodes, cudaStream_t st);

// ---- AoS <-> SoA at the C-ABI ------------------------------------------------------------------
// records [first, first+count) of the caller's AoS -> SoA slots [first, first+count), id = index
template <int D>
void launch_aos_to_soa(const float *aos, long long first, long long count, const SoA<D> &s, const int *ids,
                       cudaStream_t st);
// ids_out == NULL: every live particle whose id lies in [id0, id0+count) writes its record to aos[id - id0];
// ids_out != NULL: storage slots [id0, id0+count) are written in storage order with their ids (-1 = dead)
template <int D>
void launch_soa_to_aos(const SoA<D> &s, long long n, long long id0, long long count, float *aos, int *ids_out,
                       cudaStream_t st);
// dst[slot] = src[order[slot]] for all fields
template <int D>
void launch_reorder(const SoA<D> &src, const SoA<D> &dst, const int *order, long long n, cudaStream_t st);

// ---- binning -----------------------------------------------------------------------------------
BinGeom make_bin_geom(const Params &P, int dim, int edge);
// cell[n*D] (may be NULL), key[n]; flags STATUS_DOMAIN
template <int D>
void launch_bin_keys(const Params &P, const BinGeom &G, const SoA<D> &s, long long n, int *cell, unsigned *key,
                     int *status, bool by_id, cudaStream_t st);
// stable LSD radix sort of (key, val) pairs on `bits` key bits; returns which buffer holds the result
struct SortBuffers {
  unsigned *key[2];
  int *val[2];
  unsigned *hist;       // 256 * n_tiles
  unsigned *scan_tmp;   // scratch for the scan
  long long capacity;
};
size_t sort_hist_elems(long long n);
size_t scan_tmp_elems(long long n);
int radix_sort_pairs(SortBuffers &B, long long n, int bits, cudaStream_t st);
// storage re-sort by counting (see mpm_sort.cu): the key is the CELL inside the bin (bin * cpb + local cell), so the
// storage is cell-ordered inside every bin; counts[n_bins*cpb + 2] zeroed by the caller and scanned afterwards;
// launch_bin_starts_from_cells then takes every cpb-th entry as the bin ranges
template <int D>
void launch_count_rank(const Params &P, const BinGeom &G, const SoA<D> &s, long long n, unsigned *counts, unsigned *key,
                       unsigned *rank, int *status, cudaStream_t st, const int *dev_n = nullptr);
template <int D>
void launch_reorder_scatter(const SoA<D> &src, const SoA<D> &dst, long long first, long long n, int n_bins,
                            const int *start, const unsigned *key, const unsigned *rank, cudaStream_t st,
                            const int *dev_n = nullptr);
void launch_bin_starts_from_cells(const int *cell_start, int n_bins, int cpb, int *bin_start, cudaStream_t st);
// bin_start[n_bins+1] from sorted keys
void launch_bin_starts(const unsigned *sorted_key, long long n, int n_bins, int *bin_start, cudaStream_t st);
void launch_iota(int *v, long long n, cudaStream_t st);
// active[0..count) = ids of the non-empty bins, ascending; offs = scratch of n_bins+1 u32; count is written to offs[n_bins]
void launch_active_bins(const int *bin_start, int n_bins, unsigned *offs, unsigned *scan_tmp, int *active, cudaStream_t st);
// chunks[0..count) = (bin, first slot, particles, bin x << 16 | bin y) per chunk of <= cap particles of a non-empty
// bin, ascending bin; offs = scratch of n_bins+1 u32 (afterwards: chunks before bin b); count -> offs[n_bins]
// nb_z > 0: 3D bins, the last word is bin x << 20 | bin y << 10 | bin z
void launch_active_chunks(const int *bin_start, int n_bins, int nb_y, int cap, unsigned *offs, unsigned *scan_tmp,
                          int4 *chunks, cudaStream_t st, int nb_z = 0);
// exclusive scan of unsigned data[n] in place (tmp: scan_tmp_elems(n))
void exclusive_scan_u32(unsigned *data, long long n, unsigned *tmp, cudaStream_t st);

}  // namespace mpm
