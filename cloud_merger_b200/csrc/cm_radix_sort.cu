// cm_radix_sort.cu -- in-house onesweep LSD radix sort of (voxel key, point index) pairs for sm_100a.
//
// Replaces the std::sort(index_vector) at the heart of pcl::VoxelGrid::applyFilter (PCL 1.8.1 voxel_grid.hpp, reached
// from the reference's voxelgrid(), pc_preprocessing_main.cpp:168-177). No CUB / Thrust.
//
// One launch per 8-bit digit. The digit histograms of all passes were produced up front by k_voxel_key_hist, so a pass
// is a single sweep: each CTA takes a tile, ranks its keys by digit (ballot-based match, stable), obtains for
// each of the 256 digits the global position of its first key of that digit (chained scan; see "scanner CTAs"), and
// scatters keys and values to their final place of this pass through shared memory so that global stores are
// digit-contiguous. The number of passes is decided on the device (SortInfo.num_passes, from the significant key bits); a
// pass beyond it returns immediately, and every kernel derives the ping-pong buffer it reads from the pass number.
//
// Record layout in HBM: 32-bit keys travel as 8-byte (key, value) records in keys_a / keys_b (one 8-byte load and one
// 8-byte store per item and pass, digit runs of 128 bytes on average); 64-bit keys travel as separate key and value
// arrays (keys_*, vals_*).
//
// The kernel is instruction-issue bound before it is HBM bound, so the per-item instruction count is what was designed:
//   load 1, digit 1, match 24 (R2P + 8 VOTE + 8 predicated NOT + 4 three-input OR), rank 8 (one LDS, one predicated STS
//   on a warp-private counter row, one POPC), shared-memory placement 5, scatter 8  ~= 50 per item (was ~150).
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
// The first RS_SCANNERS CTAs of a pass do not sort. Rows of 256 look-back words are indexed -1 .. n_tiles-1; worker
// tile t publishes its 256 digit counts in row t, and needs, per digit, the global position of its first key of that
// digit = (number of keys with a smaller digit in the whole input) + (keys of that digit in tiles < t). The scanners
// produce exactly that number: scanner h owns 32 consecutive digits (one per lane), its warp q owns 8 rows of every
// 64-row batch; it starts from the exclusive scan of the pass's global digit histogram (row -1), waits until the rows'
// counts are published, and overwrites every row with the running inclusive sum. A worker then reads one row -- row
// t-1 -- instead of walking back over every tile in flight (that walk, 256 digits x dozens of rows per tile, cost as much
// L2 bandwidth as the keys themselves). Scanners are CTAs 0..7 of the grid, i.e. resident before any worker; workers
// publish their counts before they wait, so the pair cannot deadlock.
constexpr int RS_SCANNERS = 8;
constexpr int RS_SCAN_DIGITS = CM_RADIX / RS_SCANNERS;  // 32: one digit per lane
constexpr int RS_SCAN_ROWS = 8;                         // rows per warp per batch
constexpr int RS_SCAN_BATCH = RS_SCAN_ROWS * RS_WARPS;  // 64 rows per round trip
static_assert(RS_SCAN_DIGITS == 32, "one digit per lane");

__device__ __forceinline__ void scanner_load(unsigned long long (&w)[RS_SCAN_ROWS], const unsigned long long* rows,
                                             uint32_t r0, uint32_t n_tiles, uint32_t d) {
#pragma unroll
  for (int k = 0; k < RS_SCAN_ROWS; ++k) w[k] = (r0 + k < n_tiles) ? ld_cg_u64(rows + (size_t)(r0 + k) * CM_RADIX + d) : 0ull;
}

__device__ __forceinline__ void scanner_cta(unsigned long long* rows /* row 0 */, uint32_t n_tiles, uint32_t epoch,
                                            const uint32_t* hist /* this pass */, uint32_t* err, uint32_t* s_scan,
                                            uint32_t* s_tot /* [2][RS_WARPS][32] */) {
  const uint32_t tid = threadIdx.x, lane = tid & 31u, q = tid >> 5;
  const uint32_t d = blockIdx.x * RS_SCAN_DIGITS + lane;  // digit
  unsigned long long w[RS_SCAN_ROWS], wn[RS_SCAN_ROWS];
  scanner_load(w, rows, q * RS_SCAN_ROWS, n_tiles, d);
  // exclusive scan of the global digit histogram: where digit d starts in the output of this pass
  uint32_t tot;
  const uint32_t gb = block_excl_scan_256(hist[tid], s_scan, &tot);
  s_tot[tid] = gb;
  __syncthreads();
  uint32_t run = s_tot[d];
  __syncthreads();
  if (q == 0) st_cg_u64(rows - CM_RADIX + d, lb_pack(epoch, CM_LB_INCL, run));  // row -1
  uint32_t buf = 0;
  for (uint32_t t0 = 0; t0 < n_tiles; t0 += RS_SCAN_BATCH, buf ^= 1u) {
    const uint32_t r0 = t0 + q * RS_SCAN_ROWS;
    scanner_load(wn, rows, r0 + RS_SCAN_BATCH, n_tiles, d);  // next batch in flight while this one is processed
    uint32_t v[RS_SCAN_ROWS];
    uint32_t sum = 0;
#pragma unroll
    for (int k = 0; k < RS_SCAN_ROWS; ++k) {
      uint32_t c = 0;
      if (r0 + k < n_tiles) {
        unsigned long long x = w[k];
        if (!lb_ready(x, epoch)) x = lb_wait(rows + (size_t)(r0 + k) * CM_RADIX + d, epoch, err);
        c = (uint32_t)x;
      }
      sum += c;
      v[k] = sum;  // inclusive within this warp's rows
    }
    uint32_t* tb = s_tot + buf * (RS_WARPS * 32);
    tb[q * 32 + lane] = sum;
    __syncthreads();
    uint32_t before = 0, total = 0;
#pragma unroll
    for (int u = 0; u < RS_WARPS; ++u) {
      const uint32_t t = tb[u * 32 + lane];
      if ((uint32_t)u < q) before += t;
      total += t;
    }
    const uint32_t base = run + before;
#pragma unroll
    for (int k = 0; k < RS_SCAN_ROWS; ++k)
      if (r0 + k < n_tiles) st_cg_u64(rows + (size_t)(r0 + k) * CM_RADIX + d, lb_pack(epoch, CM_LB_INCL, base + v[k]));
    run += total;
#pragma unroll
    for (int k = 0; k < RS_SCAN_ROWS; ++k) w[k] = wn[k];
  }
}

// Which lanes of the warp hold a different value in bits 0..7 of x: bit j of the result is set iff lane j differs.
// One ballot per bit; a lane whose bit is set complements the ballot, so the OR over bits marks the differing lanes.
__device__ __forceinline__ uint32_t warp_diff8(uint32_t x) {
  uint32_t t[8];
#pragma unroll
  for (int b = 0; b < 8; ++b) {
    asm volatile(
        "{\n .reg .pred p;\n .reg .b32 y;\n and.b32 y, %1, %2;\n setp.ne.u32 p, y, 0;\n"
        " vote.sync.ballot.b32 %0, p, 0xffffffff;\n @p not.b32 %0, %0;\n}"
        : "=r"(t[b])
        : "r"(x), "r"(1u << b));
  }
  uint32_t a, c;
  asm("lop3.b32 %0, %1, %2, %3, 0xFE;" : "=r"(a) : "r"(t[0]), "r"(t[1]), "r"(t[2]));
  asm("lop3.b32 %0, %1, %2, %3, 0xFE;" : "=r"(c) : "r"(t[3]), "r"(t[4]), "r"(t[5]));
  asm("lop3.b32 %0, %1, %2, %3, 0xFE;" : "=r"(a) : "r"(a), "r"(t[6]), "r"(t[7]));
  return a | c;
}

__device__ __forceinline__ uint32_t lanemask_gt() {
  uint32_t m;
  asm("mov.u32 %0, %%lanemask_gt;" : "=r"(m));
  return m;
}

template <typename KeyT>
__global__ void __launch_bounds__(RS_THREADS, 3) k_onesweep_pass(const VoxelParams p, const int pass) {
  constexpr int IPT = SortCfg<KeyT>::IPT;
  constexpr int TILE = RS_THREADS * IPT;
  constexpr int WARP_ITEMS = 32 * IPT;
  constexpr bool AOS = sizeof(KeyT) == 4;  // (key, value) records of 8 bytes

  __shared__ uint32_t s_hist[RS_WARPS * CM_RADIX];  // per-warp digit counters, then per-warp first positions
  __shared__ uint32_t s_scatter[CM_RADIX];          // global position of sorted-tile position 0 of digit d
  __shared__ uint32_t s_scan[9];
  __shared__ __align__(16) unsigned long long s_k[TILE];  // AOS: records (value << 32 | key); else keys
  __shared__ uint32_t s_v[AOS ? 1 : TILE];

  const long long tr0 = clock64();
  const SortInfo si = *p.info;
  if ((uint32_t)pass >= si.num_passes) return;
  const uint32_t M = si.n_keys;
  const uint32_t n_tiles = (M + TILE - 1) / TILE;

  const uint32_t tid = threadIdx.x, lane = tid & 31u;
  const uint32_t warp = __shfl_sync(0xFFFFFFFFu, tid >> 5, 0);  // provably warp-uniform
  // tile id = blockIdx.x: CTAs of a 1-D grid are dispatched in index order (what CUB's single-pass scan relies on too;
  // the look-back watchdog covers the case that this ever fails to hold)
  const uint32_t epoch = p.epoch + 1u + (uint32_t)pass;
  unsigned long long* const rows = p.lb_sort + CM_RADIX;  // row -1 lives in front
  if (blockIdx.x < RS_SCANNERS) {
    scanner_cta(rows, n_tiles, epoch, p.hist + pass * CM_RADIX, &p.ctrl->error, s_scan, s_hist);
    return;
  }
  const uint32_t tile = blockIdx.x - RS_SCANNERS;
  if (tile >= n_tiles) return;

#define RS_TRACE(i) do { if (p.trace && p.trace_pass == (uint32_t)pass && tid == 0) p.trace[(size_t)tile * 8 + (i)] = (unsigned long long)(clock64() - tr0); } while (0)
  const bool odd = (pass & 1) != 0;
  const uint32_t shift = (uint32_t)pass * CM_RADIX_BITS;
  const uint32_t tile_base = tile * TILE;
  const uint32_t n_here = min((uint32_t)TILE, M - tile_base);
  const bool full = n_here == (uint32_t)TILE;
  const uint32_t item0 = warp * WARP_ITEMS + lane;  // tile-local index of item 0 of this thread; item i = item0 + 32 i

  // ---- load keys and values (warp-striped, coalesced). Slots past the end of a partial (last) tile get the all-ones
  // key: digit 255 in every pass, and -- being the last items of the tile -- ranked after every real key of that digit,
  // so they land at sorted-tile positions >= n_here and are simply not written back.
  KeyT key[IPT];
  uint32_t val[IPT];
  if (AOS) {
    const uint2* __restrict__ in = reinterpret_cast<const uint2*>(odd ? p.keys_b : p.keys_a) + tile_base + item0;
    if (full) {
#pragma unroll
      for (int i = 0; i < IPT; ++i) {
        const uint2 r = in[32 * i];
        key[i] = (KeyT)r.x;
        val[i] = r.y;
      }
    } else {
#pragma unroll
      for (int i = 0; i < IPT; ++i) {
        uint2 r = make_uint2(0xFFFFFFFFu, 0u);
        if (item0 + 32 * i < n_here) r = in[32 * i];
        key[i] = (KeyT)r.x;
        val[i] = r.y;
      }
    }
  } else {
    const KeyT* __restrict__ in_keys = reinterpret_cast<const KeyT*>(odd ? p.keys_b : p.keys_a) + tile_base + item0;
    const uint32_t* __restrict__ in_vals = (odd ? p.vals_b : p.vals_a) + tile_base + item0;
    if (full) {
#pragma unroll
      for (int i = 0; i < IPT; ++i) key[i] = in_keys[32 * i];
#pragma unroll
      for (int i = 0; i < IPT; ++i) val[i] = in_vals[32 * i];
    } else {
#pragma unroll
      for (int i = 0; i < IPT; ++i) key[i] = (item0 + 32 * i < n_here) ? in_keys[32 * i] : ~(KeyT)0;
#pragma unroll
      for (int i = 0; i < IPT; ++i) val[i] = (item0 + 32 * i < n_here) ? in_vals[32 * i] : 0u;
    }
  }
  // the warp's own counter row
  uint32_t* const wh = s_hist + warp * CM_RADIX;
#pragma unroll
  for (int k = 0; k < CM_RADIX / 32; ++k) wh[lane + 32 * k] = 0;
  __syncwarp();
  RS_TRACE(0);

  // ---- stable rank of every key among the keys of its digit inside the warp ----------------------------------------
  // peers = lanes holding the same digit (ballot match, constant time; __match_any_sync costs one round per distinct
  // value). Every lane reads the digit's counter; the highest peer lane writes it back advanced by the peer count.
  // Shared-memory accesses of one warp are performed in program order, __syncwarp keeps the compiler from reordering.
  const uint32_t lt = lanemask_lt(), gt = lanemask_gt();
  uint32_t rank[IPT];
#pragma unroll
  for (int i = 0; i < IPT; ++i) {
    const uint32_t ks = (uint32_t)(key[i] >> shift);
    const uint32_t diff = warp_diff8(ks);
    const uint32_t below = (uint32_t)__popc(~diff & lt);
    volatile uint32_t* c = wh + (ks & (CM_RADIX - 1));
    const uint32_t r = *c + below;
    rank[i] = r;
    if ((~diff & gt) == 0u) *c = r + 1u;
    __syncwarp();
  }
  __syncthreads();
  RS_TRACE(1);

  // ---- per digit (one per thread): tile count -> published; prefix over digits and warps -> first positions ---------
  {
    uint32_t wc[RS_WARPS];
    uint32_t cnt = 0;
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) {
      wc[w] = s_hist[w * CM_RADIX + tid];
      cnt += wc[w];
    }
    const uint32_t real = (tid == CM_RADIX - 1) ? cnt - ((uint32_t)TILE - n_here) : cnt;
    st_relaxed_u64(rows + (size_t)tile * CM_RADIX + tid, lb_pack(epoch, CM_LB_AGG, real));
    RS_TRACE(2);
    uint32_t tot;
    const uint32_t bin_start = block_excl_scan_256(cnt, s_scan, &tot);
    uint32_t run = bin_start;
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) {
      s_hist[w * CM_RADIX + tid] = run;
      run += wc[w];
    }
    s_scatter[tid] = 0u - bin_start;  // completed below once the scanners have delivered the row
  }
  __syncthreads();
  RS_TRACE(3);

  // ---- keys and values into sorted-tile order in shared memory ---------------------------------------------------------
#pragma unroll
  for (int i = 0; i < IPT; ++i) {
    const uint32_t ks = (uint32_t)(key[i] >> shift);
    const uint32_t pos = wh[ks & (CM_RADIX - 1)] + rank[i];
    if (AOS) {
      s_k[pos] = ((unsigned long long)val[i] << 32) | (unsigned long long)key[i];
    } else {
      s_k[pos] = (unsigned long long)key[i];
      s_v[pos] = val[i];
    }
  }
  RS_TRACE(4);
  // ---- one row from the scanners: where this tile's keys of digit d start in the output -------------------------------
  {
    const uint32_t first = lb_wait_inclusive(rows + ((long long)tile - 1) * CM_RADIX + tid, epoch, &p.ctrl->error);
    s_scatter[tid] += first;  // modulo 2^32
  }
  __syncthreads();
  RS_TRACE(5);

  // ---- scatter: consecutive threads write consecutive addresses inside each digit's run ------------------------------
  // (two copies of the loop: the one for full tiles carries no bounds checks, so its shared-memory loads batch up)
  if (AOS) {
    unsigned long long* __restrict__ out = reinterpret_cast<unsigned long long*>(odd ? p.keys_a : p.keys_b);
    if (full) {
      unsigned long long r[IPT];
#pragma unroll
      for (int j = 0; j < IPT; ++j) r[j] = s_k[j * RS_THREADS + tid];
#pragma unroll
      for (int j = 0; j < IPT; ++j) {
        const uint32_t dg = ((uint32_t)r[j] >> shift) & (CM_RADIX - 1);
        out[s_scatter[dg] + (j * RS_THREADS + tid)] = r[j];
      }
    } else {
      for (uint32_t pos = tid; pos < n_here; pos += RS_THREADS) {
        const unsigned long long r = s_k[pos];
        const uint32_t dg = ((uint32_t)r >> shift) & (CM_RADIX - 1);
        out[s_scatter[dg] + pos] = r;
      }
    }
  } else {
    KeyT* __restrict__ out_keys = reinterpret_cast<KeyT*>(odd ? p.keys_a : p.keys_b);
    uint32_t* __restrict__ out_vals = odd ? p.vals_a : p.vals_b;
    if (full) {
#pragma unroll
      for (int j0 = 0; j0 < IPT; j0 += 4) {
        unsigned long long kk[4];
        uint32_t vv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          kk[j] = s_k[(j0 + j) * RS_THREADS + tid];
          vv[j] = s_v[(j0 + j) * RS_THREADS + tid];
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t dg = (uint32_t)(kk[j] >> shift) & (CM_RADIX - 1);
          const uint32_t dst = s_scatter[dg] + ((j0 + j) * RS_THREADS + tid);
          out_keys[dst] = (KeyT)kk[j];
          out_vals[dst] = vv[j];
        }
      }
    } else {
      for (uint32_t pos = tid; pos < n_here; pos += RS_THREADS) {
        const unsigned long long kk = s_k[pos];
        const uint32_t dg = (uint32_t)(kk >> shift) & (CM_RADIX - 1);
        const uint32_t dst = s_scatter[dg] + pos;
        out_keys[dst] = (KeyT)kk;
        out_vals[dst] = s_v[pos];
      }
    }
  }
  RS_TRACE(6);
}

// 32-bit keys leave the sort as 8-byte (key, value) records; callers that want two plain arrays get them from here.
__global__ void __launch_bounds__(256) k_split_records(const uint2* __restrict__ rec, uint32_t* __restrict__ keys,
                                                       uint32_t* __restrict__ vals, const uint32_t* n_ptr) {
  const uint32_t n = *n_ptr;
  for (uint32_t i = blockIdx.x * 256u + threadIdx.x; i < n; i += gridDim.x * 256u) {
    const uint2 r = rec[i];
    keys[i] = r.x;
    vals[i] = r.y;
  }
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

cudaError_t launch_split_records(const void* records, uint32_t* keys, uint32_t* vals, const uint32_t* n_ptr,
                                 uint32_t max_points, cudaStream_t stream) {
  if (max_points == 0) return cudaSuccess;
  const uint32_t blocks = std::min<uint32_t>((max_points + 255u) / 256u, 148u * 8u);
  k_split_records<<<blocks, 256, 0, stream>>>(reinterpret_cast<const uint2*>(records), keys, vals, n_ptr);
  return cudaGetLastError();
}

}  // namespace cm
