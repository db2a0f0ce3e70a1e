// cm_radix_sort.cu -- in-house onesweep LSD radix sort of (voxel key, point index) pairs for sm_100a.
//
// Replaces the std::sort(index_vector) at the heart of pcl::VoxelGrid::applyFilter (PCL 1.8.1 voxel_grid.hpp, reached
// from the reference's voxelgrid(), pc_preprocessing_main.cpp:168-177). No CUB / Thrust.
//
// One launch per 8-bit digit. The digit histograms of all passes were produced up front by k_voxel_key_hist, so a pass
// is a single sweep: each CTA takes a tile, ranks its keys by digit (ballot-based match, stable), obtains for
// each of the 256 digits the number of equal-digit keys in all earlier tiles (chained scan; see "scanner CTAs"), and scatters
// keys and values to their final place of this pass through shared memory so that global stores are digit-contiguous.
// The number of passes is decided on the device (SortInfo.num_passes, from the significant key bits); a pass beyond it
// returns immediately, and every kernel derives the ping-pong buffer it reads from the pass number.
//
// Roofline: HBM. Algorithmic bytes per pass = n * 2 * (key_bytes + 4).
#include "cm_kernels.h"

namespace cm {

namespace {

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;

template <typename KeyT>
struct SortCfg;
template <>
struct SortCfg<uint32_t> {
  static constexpr int IPT = 16;
};
template <>
struct SortCfg<unsigned long long> {
  static constexpr int IPT = 12;
};

// ---- scanner CTAs -------------------------------------------------------------------------------------------------------
// The first RS_SCANNERS CTAs of a pass do not sort: scanner h owns 64 digits and walks the tile rows in order, four
// threads per digit, 32 rows per round trip: it waits until the rows' aggregates are published, and rewrites every row
// with the inclusive prefix. A worker tile then reads exactly one row -- its predecessor's inclusive prefix -- instead
// of walking back over every tile in flight (that walk, 256 digits x dozens of rows per tile, cost as much L2 bandwidth
// as the keys themselves). Scanners are CTAs 0..3 of the grid, i.e. resident before any worker; workers publish their
// aggregate before they wait, so the pair cannot deadlock.
constexpr int RS_SCANNERS = 4;
constexpr int RS_SCAN_ROWS = 8;  // rows per thread per round trip (x4 threads per digit = 32 rows)

__device__ __forceinline__ void scanner_load(unsigned long long (&w)[RS_SCAN_ROWS], const unsigned long long* st,
                                             uint32_t r0, uint32_t n_tiles, uint32_t d) {
#pragma unroll
  for (int k = 0; k < RS_SCAN_ROWS; ++k) w[k] = (r0 + k < n_tiles) ? ld_cg_u64(st + (size_t)(r0 + k) * CM_RADIX + d) : 0ull;
}

__device__ __forceinline__ void scanner_cta(unsigned long long* st, uint32_t n_tiles, uint32_t epoch, uint32_t* err) {
  const uint32_t tid = threadIdx.x;
  const uint32_t d = blockIdx.x * (CM_RADIX / RS_SCANNERS) + (tid >> 2);  // digit
  const uint32_t q = tid & 3u;                                            // which quarter of the 32-row batch
  uint32_t run = 0;                                                       // inclusive prefix of everything before the batch
  unsigned long long w[RS_SCAN_ROWS], wn[RS_SCAN_ROWS];
  scanner_load(w, st, q * RS_SCAN_ROWS, n_tiles, d);
  for (uint32_t t0 = 0; t0 < n_tiles; t0 += 4 * RS_SCAN_ROWS) {
    const uint32_t r0 = t0 + q * RS_SCAN_ROWS;
    scanner_load(wn, st, r0 + 4 * RS_SCAN_ROWS, n_tiles, d);  // next batch in flight while this one is processed
    uint32_t v[RS_SCAN_ROWS];
    uint32_t sum = 0;
#pragma unroll
    for (int k = 0; k < RS_SCAN_ROWS; ++k) {
      v[k] = 0;
      if (r0 + k < n_tiles) {
        unsigned long long x = w[k];
        if (!lb_ready(x, epoch)) x = lb_wait(st + (size_t)(r0 + k) * CM_RADIX + d, epoch, err);
        v[k] = (uint32_t)x;
      }
      sum += v[k];
      v[k] = sum;  // inclusive within this thread's rows
    }
    // inclusive prefix of `sum` over the 4 threads of the digit (lanes 4g .. 4g+3)
    uint32_t incl = sum;
    uint32_t o = __shfl_up_sync(0xFFFFFFFFu, incl, 1, 4);
    if (q >= 1) incl += o;
    o = __shfl_up_sync(0xFFFFFFFFu, incl, 2, 4);
    if (q >= 2) incl += o;
    const uint32_t base = run + incl - sum;
#pragma unroll
    for (int k = 0; k < RS_SCAN_ROWS; ++k)
      if (r0 + k < n_tiles) st_cg_u64(st + (size_t)(r0 + k) * CM_RADIX + d, lb_pack(epoch, CM_LB_INCL, base + v[k]));
    run += __shfl_sync(0xFFFFFFFFu, incl, 3, 4);  // batch total
#pragma unroll
    for (int k = 0; k < RS_SCAN_ROWS; ++k) w[k] = wn[k];
  }
}

template <typename KeyT>
__global__ void __launch_bounds__(RS_THREADS, 3) k_onesweep_pass(const VoxelParams p, const int pass) {
  constexpr int IPT = SortCfg<KeyT>::IPT;
  constexpr int TILE = RS_THREADS * IPT;
  constexpr int WARP_ITEMS = 32 * IPT;

  __shared__ uint32_t s_warp_hist[RS_WARPS][CM_RADIX];
  __shared__ uint32_t s_cnt[CM_RADIX];                       // early per-tile digit counts
  __shared__ uint32_t s_bin_start[CM_RADIX];                 // first position of digit d inside the sorted tile
  __shared__ uint32_t s_scatter[CM_RADIX];                   // global position of sorted-tile position 0 of digit d, minus s_bin_start
  __shared__ uint32_t s_scan[9];
  __shared__ __align__(16) KeyT s_keys[TILE];
  __shared__ uint32_t s_vals[TILE];

  const long long tr0 = clock64();
  const SortInfo si = *p.info;
  if ((uint32_t)pass >= si.num_passes) return;
  const uint32_t M = si.n_keys;
  const uint32_t n_tiles = (M + TILE - 1) / TILE;

  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  // tile id = blockIdx.x: CTAs of a 1-D grid are dispatched in index order (what CUB's single-pass scan relies on too;
  // the look-back watchdog covers the case that this ever fails to hold)
  const uint32_t epoch = p.epoch + 1u + (uint32_t)pass;
  if (blockIdx.x < RS_SCANNERS) {
    scanner_cta(p.lb_sort, n_tiles, epoch, &p.ctrl->error);
    return;
  }
  const uint32_t tile = blockIdx.x - RS_SCANNERS;
  if (tile >= n_tiles) return;
  const uint32_t gcount = (tid < CM_RADIX) ? p.hist[pass * CM_RADIX + tid] : 0u;  // needed late: fetch it now
  for (uint32_t i = tid; i < RS_WARPS * CM_RADIX; i += RS_THREADS) (&s_warp_hist[0][0])[i] = 0;
  if (tid < CM_RADIX) s_cnt[tid] = 0;
  __syncthreads();

#define RS_TRACE(i) do { if (p.trace && p.trace_pass == (uint32_t)pass && tid == 0) p.trace[(size_t)tile * 8 + (i)] = (unsigned long long)(clock64() - tr0); } while (0)
  RS_TRACE(0);
  const bool odd = (pass & 1) != 0;
  const KeyT* __restrict__ in_keys = reinterpret_cast<const KeyT*>(odd ? p.keys_b : p.keys_a);
  KeyT* __restrict__ out_keys = reinterpret_cast<KeyT*>(odd ? p.keys_a : p.keys_b);
  const uint32_t* __restrict__ in_vals = odd ? p.vals_b : p.vals_a;
  uint32_t* __restrict__ out_vals = odd ? p.vals_a : p.vals_b;
  const uint32_t shift = (uint32_t)pass * CM_RADIX_BITS;

  const uint32_t tile_base = tile * TILE;
  const uint32_t n_here = min((uint32_t)TILE, M - tile_base);
  const uint32_t item0 = warp * WARP_ITEMS + lane;  // tile-local index of item 0 of this thread; item i = item0 + 32 i

  // ---- load keys (warp-striped, coalesced) -----------------------------------------------------------------------
  KeyT key[IPT];
#pragma unroll
  for (int i = 0; i < IPT; ++i) {
    const uint32_t li = item0 + 32 * i;
    key[i] = (li < n_here) ? in_keys[tile_base + li] : (KeyT)0;
  }

  // values ride along; issue their loads now so the latency hides behind the ranking
  uint32_t val[IPT];
#pragma unroll
  for (int i = 0; i < IPT; ++i) {
    const uint32_t li = item0 + 32 * i;
    val[i] = (li < n_here) ? in_vals[tile_base + li] : 0u;
  }

  RS_TRACE(1);
  // ---- stable rank of every key among the keys of its digit inside the warp -----------------------------------------
  // peers = lanes holding the same digit, found with one ballot per digit bit (constant time; __match_any_sync costs
  // one round per distinct value, i.e. up to 32 rounds on the low, uniformly distributed digits)
  uint32_t peers[IPT];
#pragma unroll
  for (int i = 0; i < IPT; ++i) {
    const uint32_t li = item0 + 32 * i;
    const bool valid = li < n_here;
    const uint32_t d = (uint32_t)(key[i] >> shift) & (CM_RADIX - 1);
    uint32_t pm = __ballot_sync(0xFFFFFFFFu, valid);
#pragma unroll
    for (int b = 0; b < CM_RADIX_BITS; ++b) {
      const bool bit = (d >> b) & 1u;
      const uint32_t m = __ballot_sync(0xFFFFFFFFu, bit);
      pm &= bit ? m : ~m;
    }
    peers[i] = valid ? pm : 0u;
  }
  // ---- early counts: the tile's digit histogram goes out before the (longer) ranking, so that by the time this tile
  // walks back over its predecessors they have all published theirs (the ranking time becomes slack for stragglers)
#pragma unroll
  for (int i = 0; i < IPT; ++i) {
    const uint32_t pm = peers[i];
    if (pm && (int)lane == __ffs(pm) - 1) atomicAdd(&s_cnt[(uint32_t)(key[i] >> shift) & (CM_RADIX - 1)], (uint32_t)__popc(pm));
  }
  __syncthreads();
  const uint32_t cnt = (tid < CM_RADIX) ? s_cnt[tid] : 0u;
  if (tid < CM_RADIX) st_relaxed_u64(p.lb_sort + (size_t)tile * CM_RADIX + tid, lb_pack(epoch, CM_LB_AGG, cnt));
  RS_TRACE(2);

  uint32_t rank[IPT];
#pragma unroll
  for (int i = 0; i < IPT; ++i) {
    const uint32_t d = (uint32_t)(key[i] >> shift) & (CM_RADIX - 1);
    const uint32_t pm = peers[i];
    const int leader = pm ? (__ffs(pm) - 1) : (int)lane;
    uint32_t prev = 0;
    if (pm && (int)lane == leader) {
      prev = s_warp_hist[warp][d];
      s_warp_hist[warp][d] = prev + (uint32_t)__popc(pm);
    }
    prev = __shfl_sync(0xFFFFFFFFu, prev, leader);
    rank[i] = prev + (uint32_t)__popc(pm & lanemask_lt());
    __syncwarp();
  }
  __syncthreads();

  // ---- per digit: prefix over warps, position in the sorted tile, global base ------------------------------------------
  if (tid < CM_RADIX) {
    uint32_t run = 0;
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) {
      const uint32_t t = s_warp_hist[w][tid];
      s_warp_hist[w][tid] = run;
      run += t;
    }
  }
  uint32_t tot;
  const uint32_t bin_start = block_excl_scan_256(cnt, s_scan, &tot);
  const uint32_t gbase = block_excl_scan_256(gcount, s_scan, &tot);
  if (tid < CM_RADIX) s_bin_start[tid] = bin_start;
  __syncthreads();
  RS_TRACE(3);

  // ---- keys and values into sorted-tile order in shared memory ---------------------------------------------------------
#pragma unroll
  for (int i = 0; i < IPT; ++i) {
    const uint32_t li = item0 + 32 * i;
    if (li < n_here) {
      const uint32_t d = (uint32_t)(key[i] >> shift) & (CM_RADIX - 1);
      const uint32_t pos = s_bin_start[d] + s_warp_hist[warp][d] + rank[i];
      s_keys[pos] = key[i];
      s_vals[pos] = val[i];
    }
  }
  RS_TRACE(4);
  // ---- one row from the scanners: the inclusive prefix of the previous tile ------------------------------------------------
  if (tid < CM_RADIX) {
    const uint32_t before =
        tile == 0 ? 0u : lb_wait_inclusive(p.lb_sort + (size_t)(tile - 1) * CM_RADIX + tid, epoch, &p.ctrl->error);
    s_scatter[tid] = gbase + before - bin_start;  // modulo 2^32
  }
  __syncthreads();
  RS_TRACE(5);

  // ---- scatter: consecutive threads write consecutive addresses inside each digit's run ------------------------------
#pragma unroll
  for (int j = 0; j < IPT; ++j) {
    const uint32_t pos = j * RS_THREADS + tid;
    if (pos < n_here) {
      const KeyT kk = s_keys[pos];
      const uint32_t d = (uint32_t)(kk >> shift) & (CM_RADIX - 1);
      const uint32_t dst = s_scatter[d] + pos;
      out_keys[dst] = kk;
      out_vals[dst] = s_vals[pos];
    }
  }
  RS_TRACE(6);
}

}  // namespace

uint32_t sort_tile_items(uint32_t key_bytes) {
  return key_bytes == 4 ? RS_THREADS * SortCfg<uint32_t>::IPT : RS_THREADS * SortCfg<unsigned long long>::IPT;
}

cudaError_t launch_sort_pass(const VoxelParams& p, int pass, cudaStream_t stream) {
  const uint32_t tile = sort_tile_items(p.key_bytes);
  const uint32_t tiles = (p.max_points + tile - 1) / tile;
  if (tiles == 0) return cudaSuccess;
  if (p.key_bytes == 4)
    k_onesweep_pass<uint32_t><<<tiles + RS_SCANNERS, RS_THREADS, 0, stream>>>(p, pass);
  else
    k_onesweep_pass<unsigned long long><<<tiles + RS_SCANNERS, RS_THREADS, 0, stream>>>(p, pass);
  return cudaGetLastError();
}

}  // namespace cm
