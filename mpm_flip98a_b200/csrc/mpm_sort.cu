// mpm_sort.cu -- binning of particles by grid block: keys, a stable LSD radix sort, bin starts.
//
// The bin key is the x-major linear id of the block of `bin_edge` cells that contains the particle's
// BASE cell, base = trunc(x*inv_dx - 0.5) (cpp_validation/mls-mpm88-explained.cpp:55; truncation,
// taichi.h:7185-7187), clamped into the grid.  All outputs are integers and are bit-exact against
// the CPU binning oracle (oracle_bin in oracle/mpm_oracle.cpp): same cells, same keys, and -- the
// sort being stable -- the same permutation.
#include "mpm_kernels.cuh"

namespace mpm {

BinGeom make_bin_geom(const Params &P, int dim, int edge) {
  BinGeom G;
  G.edge = edge;
  // bases run over [0, n_grid-2] globally; x over the owned slab [slab_lo, min(slab_hi, n_grid-1))
  int xhi = P.slab_hi < P.n_grid - 1 ? P.slab_hi : P.n_grid - 1;
  G.nb[0] = (xhi - P.slab_lo + edge - 1) / edge;
  G.nb[1] = (P.n_grid - 1 + edge - 1) / edge;
  G.nb[2] = dim == 3 ? G.nb[1] : 1;
  G.n_bins = G.nb[0] * G.nb[1] * G.nb[2];
  G.cpb = dim == 3 ? edge * edge * edge : edge * edge;
  G.active = nullptr;
  G.n_active = 0;
  return G;
}

template <int D>
__global__ void k_bin_keys(Params P, BinGeom G, SoA<D> s, long long n, int *__restrict__ cell,
                           unsigned *__restrict__ key, int *__restrict__ status, bool by_id) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float x[D];
  load_pos(s, i, x);
  int base[D];
#pragma unroll
  for (int k = 0; k < D; k++) base[k] = base_coord(x[k], P.inv_dx);
  int bad = clamp_base<D>(P, base);
  const bool dead = P.multi && load_mat(s, i) == DEAD;
  if (bad && !dead) atomicOr(status, bad);
  unsigned kk = (unsigned)((base[0] - P.slab_lo) / G.edge);
#pragma unroll
  for (int k = 1; k < D; k++) kk = kk * (unsigned)G.nb[k] + (unsigned)(base[k] / G.edge);
  if (dead) kk = (unsigned)G.n_bins;  // emigrated: sorts behind every live particle
  const long long o = by_id ? (long long)s.id[i] : i;  // by_id: outputs indexed by upload order
  key[o] = kk;
  if (cell) {
#pragma unroll
    for (int k = 0; k < D; k++) cell[o * D + k] = base[k];
  }
}
template <int D>
void launch_bin_keys(const Params &P, const BinGeom &G, const SoA<D> &s, long long n, int *cell, unsigned *key,
                     int *status, bool by_id, cudaStream_t st) {
  if (n <= 0) return;
  k_bin_keys<D><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(P, G, s, n, cell, key, status, by_id);
}
template void launch_bin_keys<2>(const Params &, const BinGeom &, const SoA<2> &, long long, int *, unsigned *, int *,
                                 bool, cudaStream_t);
template void launch_bin_keys<3>(const Params &, const BinGeom &, const SoA<3> &, long long, int *, unsigned *, int *,
                                 bool, cudaStream_t);

// ------------------------------------------------------------------------------------------------
// exclusive scan (u32), three-phase, recursive on the block sums.  1024 elements per block.
// ------------------------------------------------------------------------------------------------
static const int SCAN_BLOCK = 1024;  // elements per CTA (256 threads x 4)

__device__ __forceinline__ unsigned warp_incl_scan(unsigned v) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    unsigned t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// scans one block of 1024 in place; writes the block total to sums[blockIdx] (if sums)
__global__ void __launch_bounds__(256) k_scan_block(unsigned *__restrict__ data, long long n, unsigned *__restrict__ sums) {
  __shared__ unsigned wsum[8];
  long long base = (long long)blockIdx.x * SCAN_BLOCK + threadIdx.x * 4;
  unsigned v[4];
#pragma unroll
  for (int k = 0; k < 4; k++) v[k] = base + k < n ? data[base + k] : 0u;
  unsigned tsum = v[0] + v[1] + v[2] + v[3];
  unsigned inc = warp_incl_scan(tsum);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 31) wsum[w] = inc;
  __syncthreads();
  if (w == 0) {
    unsigned s = lane < 8 ? wsum[lane] : 0u;
    unsigned si = warp_incl_scan(s);
    if (lane < 8) wsum[lane] = si - s;
    if (lane == 7 && sums) sums[blockIdx.x] = si;
  }
  __syncthreads();
  unsigned run = wsum[w] + inc - tsum;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    if (base + k < n) data[base + k] = run;
    run += v[k];
  }
}
__global__ void __launch_bounds__(256) k_scan_add(unsigned *__restrict__ data, long long n, const unsigned *__restrict__ sums) {
  long long base = (long long)blockIdx.x * SCAN_BLOCK + threadIdx.x * 4;
  unsigned add = sums[blockIdx.x];
#pragma unroll
  for (int k = 0; k < 4; k++)
    if (base + k < n) data[base + k] += add;
}
size_t scan_tmp_elems(long long n) {
  size_t total = 0;
  while (n > 1) {
    n = (n + SCAN_BLOCK - 1) / SCAN_BLOCK;
    total += (size_t)n + 1;
    if (n == 1) break;
  }
  return total + 4;
}
void exclusive_scan_u32(unsigned *data, long long n, unsigned *tmp, cudaStream_t st) {
  if (n <= 0) return;
  long long blocks = (n + SCAN_BLOCK - 1) / SCAN_BLOCK;
  if (blocks == 1) {
    k_scan_block<<<1, 256, 0, st>>>(data, n, nullptr);
    return;
  }
  k_scan_block<<<(unsigned)blocks, 256, 0, st>>>(data, n, tmp);
  exclusive_scan_u32(tmp, blocks, tmp + blocks + 1, st);
  k_scan_add<<<(unsigned)blocks, 256, 0, st>>>(data, n, tmp);
}

// ------------------------------------------------------------------------------------------------
// LSD radix sort, 8 bits per pass, stable.  Tile = 256 threads x 16 keys, warp w owns the
// contiguous segment [w*512, (w+1)*512) of its tile and walks it in rounds of 32 so that
// "earlier in memory" == "earlier (round, lane)" -- which is what makes the in-tile rank stable.
// ------------------------------------------------------------------------------------------------
static const int RS_THREADS = 256, RS_ITEMS = 16, RS_TILE = RS_THREADS * RS_ITEMS, RS_WARPS = RS_THREADS / 32;

size_t sort_hist_elems(long long n) { return (size_t)256 * (size_t)((n + RS_TILE - 1) / RS_TILE) + 1; }

__global__ void __launch_bounds__(RS_THREADS) k_rs_hist(const unsigned *__restrict__ key, long long n, int shift,
                                                        unsigned *__restrict__ hist, unsigned n_tiles) {
  __shared__ unsigned cnt[256];
  cnt[threadIdx.x] = 0;
  __syncthreads();
  long long base = (long long)blockIdx.x * RS_TILE;
#pragma unroll 4
  for (int r = 0; r < RS_ITEMS; r++) {
    long long i = base + r * RS_THREADS + threadIdx.x;
    if (i < n) atomicAdd(&cnt[(key[i] >> shift) & 255u], 1u);
  }
  __syncthreads();
  hist[(size_t)threadIdx.x * n_tiles + blockIdx.x] = cnt[threadIdx.x];  // digit-major for the scan
}

__global__ void __launch_bounds__(RS_THREADS) k_rs_scatter(const unsigned *__restrict__ key_in, const int *__restrict__ val_in,
                                                           unsigned *__restrict__ key_out, int *__restrict__ val_out,
                                                           long long n, int shift, const unsigned *__restrict__ hist,
                                                           unsigned n_tiles) {
  __shared__ unsigned cnt[RS_WARPS][256];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int k = threadIdx.x; k < RS_WARPS * 256; k += RS_THREADS) (&cnt[0][0])[k] = 0;
  __syncthreads();
  const long long seg = (long long)blockIdx.x * RS_TILE + (long long)w * (32 * RS_ITEMS);
  unsigned kreg[RS_ITEMS];
  unsigned rank[RS_ITEMS];
#pragma unroll
  for (int r = 0; r < RS_ITEMS; r++) {
    long long i = seg + r * 32 + lane;
    bool valid = i < n;
    unsigned k = valid ? key_in[i] : 0xffffffffu;
    kreg[r] = k;
    unsigned d = (k >> shift) & 255u;
    // invalid lanes get a digit of their own (256) so they never join a valid group
    unsigned m = __match_any_sync(0xffffffffu, valid ? d : 256u);
    int leader = __ffs(m) - 1;
    unsigned before = __popc(m & ((1u << lane) - 1u));
    unsigned old = 0;
    if (valid && lane == leader) {
      old = cnt[w][d];
      cnt[w][d] = old + __popc(m);
    }
    old = __shfl_sync(0xffffffffu, old, leader);
    rank[r] = old + before;
    __syncwarp();
  }
  __syncthreads();
  {  // exclusive prefix over the warps, per digit, plus the tile's global base
    unsigned d = threadIdx.x;
    unsigned run = hist[(size_t)d * n_tiles + blockIdx.x];
#pragma unroll
    for (int ww = 0; ww < RS_WARPS; ww++) {
      unsigned c = cnt[ww][d];
      cnt[ww][d] = run;
      run += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < RS_ITEMS; r++) {
    long long i = seg + r * 32 + lane;
    if (i < n) {
      unsigned d = (kreg[r] >> shift) & 255u;
      unsigned pos = cnt[w][d] + rank[r];
      key_out[pos] = kreg[r];
      val_out[pos] = val_in[i];
    }
  }
}

int radix_sort_pairs(SortBuffers &B, long long n, int bits, cudaStream_t st) {
  int cur = 0;
  if (n <= 0) return cur;
  unsigned n_tiles = (unsigned)((n + RS_TILE - 1) / RS_TILE);
  for (int shift = 0; shift < bits; shift += 8) {
    k_rs_hist<<<n_tiles, RS_THREADS, 0, st>>>(B.key[cur], n, shift, B.hist, n_tiles);
    exclusive_scan_u32(B.hist, (long long)256 * n_tiles, B.scan_tmp, st);
    k_rs_scatter<<<n_tiles, RS_THREADS, 0, st>>>(B.key[cur], B.val[cur], B.key[cur ^ 1], B.val[cur ^ 1], n, shift,
                                                 B.hist, n_tiles);
    cur ^= 1;
  }
  return cur;
}

// bin_start[b] = first slot whose key >= b  (bin_start[n_bins] = n)
__global__ void k_bin_starts(const unsigned *__restrict__ key, long long n, int n_bins, int *__restrict__ bin_start) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  long long prev = i == 0 ? -1 : (long long)key[i - 1];
  long long cur = i == n ? (long long)n_bins : (long long)key[i];
  for (long long b = prev + 1; b <= cur; b++) bin_start[b] = (int)i;
}
void launch_bin_starts(const unsigned *sorted_key, long long n, int n_bins, int *bin_start, cudaStream_t st) {
  k_bin_starts<<<(unsigned)((n + 1 + 255) / 256), 256, 0, st>>>(sorted_key, n, n_bins, bin_start);
}

__global__ void k_bin_starts_from_cells(const int *__restrict__ cell_start, int n_bins, int cpb, int *__restrict__ bin_start) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b <= n_bins) bin_start[b] = cell_start[(size_t)b * cpb];
  if (b == n_bins) bin_start[b + 1] = cell_start[(size_t)n_bins * cpb + 1];
}
void launch_bin_starts_from_cells(const int *cell_start, int n_bins, int cpb, int *bin_start, cudaStream_t st) {
  k_bin_starts_from_cells<<<(unsigned)((n_bins + 1 + 255) / 256), 256, 0, st>>>(cell_start, n_bins, cpb, bin_start);
}

__global__ void k_flag_active(const int *__restrict__ bin_start, int n_bins, unsigned *__restrict__ offs) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b <= n_bins) offs[b] = (b < n_bins && bin_start[b + 1] > bin_start[b]) ? 1u : 0u;
}
__global__ void k_compact_active(const int *__restrict__ bin_start, int n_bins, const unsigned *__restrict__ offs,
                                 int *__restrict__ active) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < n_bins && bin_start[b + 1] > bin_start[b]) active[offs[b]] = b;
}
void launch_active_bins(const int *bin_start, int n_bins, unsigned *offs, unsigned *scan_tmp, int *active, cudaStream_t st) {
  unsigned blocks = (unsigned)((n_bins + 1 + 255) / 256);
  k_flag_active<<<blocks, 256, 0, st>>>(bin_start, n_bins, offs);
  exclusive_scan_u32(offs, (long long)n_bins + 1, scan_tmp, st);  // offs[n_bins] = number of active bins
  k_compact_active<<<blocks, 256, 0, st>>>(bin_start, n_bins, offs, active);
}

// ------------------------------------------------------------------------------------------------
// Storage re-sort by counting (the engine's periodic re-binning; mpm_bin_particles keeps the stable radix
// sort above because ITS permutation is compared bit-exactly with the CPU binning oracle).
// Particles are already nearly in order, so one pass suffices: every slot takes a rank inside its new
// CELL (key = bin * cells-per-bin + cell inside the bin) from a warp-aggregated counter (MATCH.ANY groups the
// lanes of a warp by key: about one atomic per warp and key), the counters are scanned into cell starts (every
// cpb-th of which is a bin start), and the consumer -- k_reorder_scatter here, or the substep kernel itself
// (RESORT) -- writes each particle to start[key] + rank in the other buffer.  12 bytes of traffic per particle
// instead of three 16-byte radix passes.  Ordering by cell INSIDE the bin is what makes the substep kernel's
// grid gather coalesce: the lanes of a warp then read the same few node rows (the round-2 profile had the L1
// data pipe at 78 % of its wavefront peak with bin-only order: ~12 sectors per gather instruction).
// ------------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(256) k_count_rank(Params P, BinGeom G, SoA<D> s, long long n, unsigned *__restrict__ counts,
                                                    unsigned *__restrict__ key_out, unsigned *__restrict__ rank_out,
                                                    int *__restrict__ status, const int *__restrict__ dev_n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (dev_n && n > *dev_n) n = *dev_n;  // x-slab handles: exact extent on the device
  const bool valid = i < n;
  unsigned kk = 0xffffffffu;
  if (valid) {
    float x[D];
    load_pos(s, i, x);
    int base[D];
#pragma unroll
    for (int k = 0; k < D; k++) base[k] = base_coord(x[k], P.inv_dx);
    const int bad = clamp_base<D>(P, base);
    const bool dead = P.multi && load_mat(s, i) == DEAD;
    if (bad && !dead) atomicOr(status, bad);
    const int x0 = base[0] - P.slab_lo;
    kk = (unsigned)(x0 / G.edge);
    unsigned local = (unsigned)(x0 % G.edge);
#pragma unroll
    for (int k = 1; k < D; k++) {
      kk = kk * (unsigned)G.nb[k] + (unsigned)(base[k] / G.edge);
      local = local * (unsigned)G.edge + (unsigned)(base[k] % G.edge);
    }
    kk = kk * (unsigned)G.cpb + local;
    if (dead) kk = (unsigned)G.n_bins * (unsigned)G.cpb;  // emigrated: behind every live particle, dropped by the consumer
  }
  const unsigned lane = threadIdx.x & 31u;
  const unsigned m = __match_any_sync(0xffffffffu, kk);
  const int leader = __ffs(m) - 1;
  unsigned old = 0;
  if (valid && (int)lane == leader) old = atomicAdd(&counts[kk], (unsigned)__popc(m));
  old = __shfl_sync(0xffffffffu, old, leader);
  if (valid) {
    key_out[i] = kk;
    rank_out[i] = old + (unsigned)__popc(m & ((1u << lane) - 1u));
  }
}
template <int D>
void launch_count_rank(const Params &P, const BinGeom &G, const SoA<D> &s, long long n, unsigned *counts, unsigned *key,
                       unsigned *rank, int *status, cudaStream_t st, const int *dev_n) {
  if (n <= 0) return;
  k_count_rank<D><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(P, G, s, n, counts, key, rank, status, dev_n);
}
template void launch_count_rank<2>(const Params &, const BinGeom &, const SoA<2> &, long long, unsigned *, unsigned *,
                                   unsigned *, int *, cudaStream_t, const int *);
template void launch_count_rank<3>(const Params &, const BinGeom &, const SoA<3> &, long long, unsigned *, unsigned *,
                                   unsigned *, int *, cudaStream_t, const int *);

// slots [first, n) of `src` -> dst[start[key] + rank]; dead slots (key == n_bins) are dropped
template <int D>
__global__ void __launch_bounds__(256) k_reorder_scatter(SoA<D> src, SoA<D> dst, long long first, long long n, int n_bins,
                                                         const int *__restrict__ start, const unsigned *__restrict__ key,
                                                         const unsigned *__restrict__ rank, const int *__restrict__ dev_n) {
  const long long i = first + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (dev_n && n > *dev_n) n = *dev_n;
  if (i >= n) return;
  const unsigned kk = key[i];
  if (kk >= (unsigned)n_bins) return;
  const long long d = (long long)start[kk] + rank[i];
  PState<D> p;
  load_full(src, i, p);
  store_state(dst, d, p);
  store_tags(dst, d, p.mat, src.id[i]);
}
template <int D>
void launch_reorder_scatter(const SoA<D> &src, const SoA<D> &dst, long long first, long long n, int n_bins,
                            const int *start, const unsigned *key, const unsigned *rank, cudaStream_t st, const int *dev_n) {
  if (n - first <= 0) return;
  k_reorder_scatter<D><<<(unsigned)((n - first + 255) / 256), 256, 0, st>>>(src, dst, first, n, n_bins, start, key, rank,
                                                                            dev_n);
}
template void launch_reorder_scatter<2>(const SoA<2> &, const SoA<2> &, long long, long long, int, const int *,
                                        const unsigned *, const unsigned *, cudaStream_t, const int *);
template void launch_reorder_scatter<3>(const SoA<3> &, const SoA<3> &, long long, long long, int, const int *,
                                        const unsigned *, const unsigned *, cudaStream_t, const int *);

// Work list of the 2D substep kernel: one entry per CHUNK of at most `cap` particles of a non-empty bin,
// (bin, first slot, particles, bin x << 16 | bin y).  A CTA reads its entry with one load (no bin -> range -> particle chain of
// dependent loads at CTA start), and a bin that a collapsing scene has filled far beyond the average is spread
// over several CTAs instead of serialising its chunks in one.
__global__ void k_flag_chunks(const int *__restrict__ bin_start, int n_bins, int cap, unsigned *__restrict__ offs) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b <= n_bins) offs[b] = b < n_bins ? (unsigned)((bin_start[b + 1] - bin_start[b] + cap - 1) / cap) : 0u;
}
__global__ void k_fill_chunks(const int *__restrict__ bin_start, int n_bins, int nb_y, int nb_z, int cap,
                              const unsigned *__restrict__ offs, int4 *__restrict__ chunks) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_bins) return;
  const int s0 = bin_start[b], cnt = bin_start[b + 1] - s0;
  // bin coordinates for the substep kernels: 2D x << 16 | y; 3D x << 20 | y << 10 | z
  const int xy = nb_z > 0 ? (((b / nb_z) / nb_y) << 20) | (((b / nb_z) % nb_y) << 10) | (b % nb_z)
                          : ((b / nb_y) << 16) | (b % nb_y);
  unsigned o = offs[b];
  // a bin that needs several chunks is split EVENLY (1024 particles -> 512 + 512, not 768 + 256): the CTAs of the
  // substep kernel then run for similar times, which is what its static (persistent, strided) work distribution assumes
  const int parts = (cnt + cap - 1) / cap;  // == what k_flag_chunks reserved
  for (int j = 0; j < parts; j++) {
    const int a = (int)((long long)cnt * j / parts), e = (int)((long long)cnt * (j + 1) / parts);
    chunks[o++] = make_int4(b, s0 + a, e - a, xy);
  }
}
void launch_active_chunks(const int *bin_start, int n_bins, int nb_y, int cap, unsigned *offs, unsigned *scan_tmp,
                          int4 *chunks, cudaStream_t st, int nb_z) {
  unsigned blocks = (unsigned)((n_bins + 1 + 255) / 256);
  k_flag_chunks<<<blocks, 256, 0, st>>>(bin_start, n_bins, cap, offs);
  exclusive_scan_u32(offs, (long long)n_bins + 1, scan_tmp, st);  // offs[n_bins] = number of chunks
  k_fill_chunks<<<blocks, 256, 0, st>>>(bin_start, n_bins, nb_y, nb_z, cap, offs, chunks);
}

__global__ void k_iota(int *v, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = (int)i;
}
void launch_iota(int *v, long long n, cudaStream_t st) {
  if (n <= 0) return;
  k_iota<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(v, n);
}

}  // namespace mpm
