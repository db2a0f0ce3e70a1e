// cm_voxel.cu -- VoxelGrid front end (bounding box -> grid -> 64-bit voxel keys + digit histograms) and back end
// (run detection on the sorted keys -> centroid per voxel -> min-points filter -> stable compaction) for sm_100a.
//
// Replaces pcl::VoxelGrid<pcl::PointXYZI>::applyFilter as configured by the reference's voxelgrid()
// (pc_preprocessing_main.cpp:168-177, CloudFusionNode.h:276-289, PreprocessingNode.h:236-249):
//   first pass  (idx per point)             -> k_voxel_key_hist   (key = (frame << idx_bits) | idx, idx = i + j*div_x + k*div_x*div_y)
//   second pass (std::sort)                 -> cm_radix_sort.cu   (onesweep LSD radix sort of key / point-index pairs)
//   third+fourth pass (runs, CentroidPoint) -> k_voxel_centroid
// The voxel index arithmetic repeats PCL's float32 formulas exactly (floorf(x * inv_leaf) with a separately rounded
// multiply); the linearisation is carried in 64 bits so clouds beyond PCL's INT32 cell limit still work.
//
// Roofline: HBM for every kernel here. Algorithmic bytes: key_hist reads 16 B and writes key_bytes per point;
// centroid reads key_bytes + 4 per point, gathers 16 B per point and writes 16/32 + 4 + 8 B per voxel.
#include "cm_kernels.h"

namespace cm {

namespace {

constexpr int VX_THREADS = 256;
constexpr int KH_IPT = 8;
constexpr int KH_TILE = VX_THREADS * KH_IPT;  // 2048 points per key/hist tile
constexpr int CE_IPT = 4;
constexpr int CE_TILE = VX_THREADS * CE_IPT;  // 1024 sorted items per centroid tile

// largest f in [0, n_frames) with fstart[f] <= idx
__device__ __forceinline__ uint32_t find_frame(const uint32_t* __restrict__ fstart, uint32_t n_frames, uint32_t idx) {
  uint32_t lo = 0, hi = n_frames - 1;
  while (lo < hi) {
    const uint32_t mid = (lo + hi + 1) >> 1;
    if (__ldg(fstart + mid) <= idx) lo = mid; else hi = mid - 1;
  }
  return lo;
}

}  // namespace

// ---- pcl::getMinMax3D on a plain packed cloud (VoxelGrid-only entry) -------------------------------------------------
__global__ void __launch_bounds__(VX_THREADS) k_minmax(const float4* __restrict__ pts, uint32_t n, Ctrl* ctrl,
                                                       FrameAcc* acc, uint32_t* frame_surv_start) {
  __shared__ float s_mm[VX_THREADS / 32][6];
  __shared__ uint32_t s_inv[VX_THREADS / 32];
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  float mn0 = 3.402823466e+38f, mn1 = mn0, mn2 = mn0, mx0 = -mn0, mx1 = -mn0, mx2 = -mn0;
  uint32_t inv = 0, fin_cnt = 0;
  for (size_t i = (size_t)blockIdx.x * VX_THREADS + tid; i < n; i += (size_t)gridDim.x * VX_THREADS) {
    const float4 v = ldg_stream_f4(pts + i);
    if (finite_f32(v.x) && finite_f32(v.y) && finite_f32(v.z)) {
      mn0 = fminf(mn0, v.x); mx0 = fmaxf(mx0, v.x);
      mn1 = fminf(mn1, v.y); mx1 = fmaxf(mx1, v.y);
      mn2 = fminf(mn2, v.z); mx2 = fmaxf(mx2, v.z);
      ++fin_cnt;
    } else {
      ++inv;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn0 = fminf(mn0, __shfl_xor_sync(0xFFFFFFFFu, mn0, o)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xFFFFFFFFu, mx0, o));
    mn1 = fminf(mn1, __shfl_xor_sync(0xFFFFFFFFu, mn1, o)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xFFFFFFFFu, mx1, o));
    mn2 = fminf(mn2, __shfl_xor_sync(0xFFFFFFFFu, mn2, o)); mx2 = fmaxf(mx2, __shfl_xor_sync(0xFFFFFFFFu, mx2, o));
  }
  inv = warp_sum_u32(inv);
  fin_cnt = warp_sum_u32(fin_cnt);
  if (lane == 0) {
    s_mm[warp][0] = mn0; s_mm[warp][1] = mn1; s_mm[warp][2] = mn2;
    s_mm[warp][3] = mx0; s_mm[warp][4] = mx1; s_mm[warp][5] = mx2;
    s_inv[warp] = inv;
    if (fin_cnt == 0) s_mm[warp][0] = 3.402823466e+38f;  // keeps the fold below neutral
  }
  __syncthreads();
  if (tid == 0) {
    float a0 = s_mm[0][0], a1 = s_mm[0][1], a2 = s_mm[0][2], b0 = s_mm[0][3], b1 = s_mm[0][4], b2 = s_mm[0][5];
    uint32_t ti = s_inv[0];
    for (int w = 1; w < VX_THREADS / 32; ++w) {
      a0 = fminf(a0, s_mm[w][0]); a1 = fminf(a1, s_mm[w][1]); a2 = fminf(a2, s_mm[w][2]);
      b0 = fmaxf(b0, s_mm[w][3]); b1 = fmaxf(b1, s_mm[w][4]); b2 = fmaxf(b2, s_mm[w][5]);
      ti += s_inv[w];
    }
    if (b0 >= a0) {  // at least one finite point seen by this block
      atomicMax(&acc->nmin_enc[0], ~f32_order_enc(__float_as_uint(a0)));
      atomicMax(&acc->nmin_enc[1], ~f32_order_enc(__float_as_uint(a1)));
      atomicMax(&acc->nmin_enc[2], ~f32_order_enc(__float_as_uint(a2)));
      atomicMax(&acc->max_enc[0], f32_order_enc(__float_as_uint(b0)));
      atomicMax(&acc->max_enc[1], f32_order_enc(__float_as_uint(b1)));
      atomicMax(&acc->max_enc[2], f32_order_enc(__float_as_uint(b2)));
    }
    if (ti) {
      atomicAdd(&acc->n_invalid, ti);
      atomicOr(&ctrl->has_invalid, 1u);
    }
    if (blockIdx.x == 0) {
      frame_surv_start[0] = 0;
      frame_surv_start[1] = n;
    }
  }
}

// ---- grid per frame + sort plan (one block) -----------------------------------------------------------------------
// PCL 1.8.1: min_b = floor(min_p * inv), max_b = floor(max_p * inv), div_b = max_b - min_b + 1, and the guard
// dx*dy*dz > INT32_MAX with d = (int64)((max_p - min_p) * inv) + 1.
__global__ void __launch_bounds__(VX_THREADS) k_grid_setup(const VoxelParams p) {
  __shared__ uint32_t s_maxbits;
  const uint32_t tid = threadIdx.x;
  if (tid == 0) s_maxbits = 0;
  __syncthreads();
  for (uint32_t f = tid; f < p.n_frames; f += VX_THREADS) {
    const FrameAcc a = p.acc[f];
    GridDev g;
    g.pcl_overflow = 0;
    g.bits = 0;
    g.mul1 = 0;
    g.mul2 = 0;
    g.empty = (a.max_enc[0] == 0u && a.nmin_enc[0] == 0u) ? 1u : 0u;
    if (g.empty) {
      for (int k = 0; k < 3; ++k) { g.min_b[k] = 0; g.max_b[k] = 0; g.div_b[k] = 0; }
    } else {
      long long d[3];
      unsigned long long div[3];
      bool range_err = false;
      for (int k = 0; k < 3; ++k) {
        const float mn = __uint_as_float(f32_order_dec(~a.nmin_enc[k]));
        const float mx = __uint_as_float(f32_order_dec(a.max_enc[k]));
        const float inv = p.inv_leaf[k];
        d[k] = __float2ll_rz(__fmul_rn(__fsub_rn(mx, mn), inv)) + 1ll;
        const float fmn = floorf(__fmul_rn(mn, inv)), fmx = floorf(__fmul_rn(mx, inv));
        // coordinates are limited to +-2^30 cells so that the int32 grid origin is exact
        if (!(fabsf(fmn) < 1073741824.f) || !(fabsf(fmx) < 1073741824.f)) range_err = true;
        g.min_b[k] = (int)fmn;
        g.max_b[k] = (int)fmx;
        const long long dv = (long long)g.max_b[k] - (long long)g.min_b[k] + 1ll;
        g.div_b[k] = (int)dv;
        div[k] = (unsigned long long)dv;
        if (dv > (1ll << 21)) range_err = true;
      }
      const long long imax = 2147483647ll;
      bool ovf = d[0] > imax || d[1] > imax || d[2] > imax || d[0] < 0 || d[1] < 0 || d[2] < 0;
      if (!ovf) {
        const long long p01 = d[0] * d[1];
        ovf = p01 > imax || p01 * d[2] > imax;
      }
      g.pcl_overflow = ovf ? 1 : 0;
      if (range_err) {
        atomicExch(&p.ctrl->error, (uint32_t)CM_DEV_E_KEY_RANGE);
        div[0] = div[1] = div[2] = 1;
      }
      g.mul1 = div[0];
      g.mul2 = div[0] * div[1];
      const unsigned long long cells = div[0] * div[1] * div[2];
      g.bits = cells <= 1ull ? 0u : (uint32_t)(64 - __clzll((long long)(cells - 1ull)));
      atomicMax(&s_maxbits, g.bits);
    }
    p.grid[f] = g;
  }
  __syncthreads();
  if (tid == 0) {
    const uint32_t key_frames = p.n_frames + (p.ctrl->has_invalid ? 1u : 0u);
    const uint32_t frame_bits = key_frames <= 1u ? 0u : (uint32_t)(32 - __clz((int)(key_frames - 1u)));
    uint32_t idx_bits = s_maxbits;
    uint32_t total = frame_bits + idx_bits;
    if (total > p.key_bytes * 8u) {
      atomicExch(&p.ctrl->error, (uint32_t)CM_DEV_E_KEY_RANGE);
      total = p.key_bytes * 8u;
      idx_bits = total - frame_bits;
    }
    uint32_t passes = (total + CM_RADIX_BITS - 1) / CM_RADIX_BITS;
    if (passes == 0) passes = 1;
    if (passes > p.max_passes) {  // the host enqueued fewer pass launches than the data needs
      atomicExch(&p.ctrl->error, (uint32_t)CM_DEV_E_KEY_RANGE);
      passes = p.max_passes;
    }
    SortInfo si;
    si.num_passes = passes;
    si.total_bits = total;
    si.idx_bits = idx_bits;
    si.n_keys = p.frame_surv_start[p.n_frames];
    si.key_frames = key_frames;
    si.pad_[0] = si.pad_[1] = si.pad_[2] = 0;
    *p.info = si;
  }
}

// ---- voxel key per point + the digit histograms of every radix pass -----------------------------------------------------
template <typename KeyT>
__global__ void __launch_bounds__(VX_THREADS) k_voxel_key_hist(const VoxelParams p) {
  __shared__ uint32_t s_hist[CM_MAX_SORT_PASSES][CM_RADIX];
  const uint32_t tid = threadIdx.x, lane = tid & 31u;
  for (uint32_t i = tid; i < CM_MAX_SORT_PASSES * CM_RADIX; i += VX_THREADS) (&s_hist[0][0])[i] = 0;
  __syncthreads();

  const uint32_t F = p.n_frames;
  const uint32_t M = p.frame_surv_start[F];
  const SortInfo si = *p.info;
  const uint32_t n_pass = si.num_passes, idx_bits = si.idx_bits;
  const uint32_t n_tiles = (M + KH_TILE - 1) / KH_TILE;
  KeyT* __restrict__ keys = reinterpret_cast<KeyT*>(p.keys_a);
  const unsigned long long sentinel = (unsigned long long)F << idx_bits;
  const float inv0 = p.inv_leaf[0], inv1 = p.inv_leaf[1], inv2 = p.inv_leaf[2];

  for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const uint32_t base = tile * KH_TILE;
    const uint32_t last = min(base + KH_TILE, M) - 1;
    uint32_t f_lo = 0, f_hi = 0;
    if (F > 1) {
      f_lo = find_frame(p.frame_surv_start, F, base);
      f_hi = find_frame(p.frame_surv_start, F, last);
    }
    GridDev g = p.grid[f_lo];
    uint32_t f_cur = f_lo;
#pragma unroll
    for (int i = 0; i < KH_IPT; ++i) {
      const uint32_t idx = base + i * VX_THREADS + tid;
      const bool valid = idx < M;
      unsigned long long key = 0;
      if (valid) {
        const float4 v = ldg_stream_f4(p.pts + idx);
        uint32_t f = f_cur;
        while (f < f_hi && idx >= __ldg(p.frame_surv_start + f + 1)) ++f;
        if (f != f_cur) { g = p.grid[f]; f_cur = f; }
        if (finite_f32(v.x) && finite_f32(v.y) && finite_f32(v.z)) {
          // PCL: ijk = (int)(floor(x * inv) - (float)min_b); here in integers (identical below 2^24 cells)
          const long long i0 = (long long)(int)floorf(__fmul_rn(v.x, inv0)) - (long long)g.min_b[0];
          const long long i1 = (long long)(int)floorf(__fmul_rn(v.y, inv1)) - (long long)g.min_b[1];
          const long long i2 = (long long)(int)floorf(__fmul_rn(v.z, inv2)) - (long long)g.min_b[2];
          const unsigned long long cell =
              (unsigned long long)i0 + (unsigned long long)i1 * g.mul1 + (unsigned long long)i2 * g.mul2;
          key = ((unsigned long long)f << idx_bits) | cell;
        } else {
          key = sentinel;
        }
        keys[idx] = (KeyT)key;
      }
      // digit histograms; a warp whose valid lanes all share the digit adds once
      const uint32_t vmask = __ballot_sync(0xFFFFFFFFu, valid);
      if (vmask) {
        const int leader = __ffs(vmask) - 1;
        for (uint32_t ps = 0; ps < n_pass; ++ps) {
          const uint32_t d = (uint32_t)(key >> (ps * CM_RADIX_BITS)) & (CM_RADIX - 1);
          const uint32_t dl = __shfl_sync(0xFFFFFFFFu, d, leader);
          const bool uniform = __all_sync(0xFFFFFFFFu, !valid || d == dl);
          if (uniform) {
            if ((int)lane == leader) atomicAdd(&s_hist[ps][dl], (uint32_t)__popc(vmask));
          } else if (valid) {
            atomicAdd(&s_hist[ps][d], 1u);
          }
        }
      }
    }
  }
  __syncthreads();
  for (uint32_t i = tid; i < n_pass * CM_RADIX; i += VX_THREADS) {
    const uint32_t c = (&s_hist[0][0])[i];
    if (c) atomicAdd(p.hist + i, c);
  }
}

// ---- runs of equal key -> centroid, min-points filter, order-preserving compaction ----------------------------------
// The radix sort is stable and started from ascending point order, so inside a run the points are in ascending index
// order: the float accumulation below is sequential in that order (one valid order of PCL's CentroidPoint loop, whose
// own order is unspecified because std::sort is unstable). centroid = sum / (float)n with an IEEE division.
template <typename KeyT>
__global__ void __launch_bounds__(VX_THREADS) k_voxel_centroid(const VoxelParams p) {
  __shared__ uint32_t s_tile;
  __shared__ uint32_t s_scan[9];
  __shared__ uint32_t s_lb[2 * (VX_THREADS / 32) + 1];
  const uint32_t tid = threadIdx.x;
  const uint32_t F = p.n_frames;
  const uint32_t M = p.frame_surv_start[F];
  const uint32_t n_tiles = (M + CE_TILE - 1) / CE_TILE;
  if (tid == 0) s_tile = atomicAdd(&p.ctrl->tile_counter[9], 1u);
  __syncthreads();
  const uint32_t tile = s_tile;
  if (tile >= n_tiles) return;

  const SortInfo si = *p.info;
  const uint32_t idx_bits = si.idx_bits;
  const bool odd = (si.num_passes & 1u) != 0u;
  const KeyT* __restrict__ keys = reinterpret_cast<const KeyT*>(odd ? p.keys_b : p.keys_a);
  const uint32_t* __restrict__ vals = odd ? p.vals_b : p.vals_a;
  const unsigned long long limit = (si.key_frames > F) ? ((unsigned long long)F << idx_bits) : ~0ull;
  const unsigned long long idx_mask = idx_bits >= 64 ? ~0ull : ((1ull << idx_bits) - 1ull);
  const uint32_t m_req = p.min_points > 1u ? p.min_points : 1u;

  const uint32_t tile_base = tile * CE_TILE;
  const uint32_t base = tile_base + tid * CE_IPT;
  KeyT k[CE_IPT];
#pragma unroll
  for (int j = 0; j < CE_IPT; ++j) k[j] = (base + j < M) ? keys[base + j] : (KeyT)0;
  KeyT prev = (KeyT)0;
  if (base > 0 && base < M) prev = keys[base - 1];

  uint32_t passbits = 0, cnt = 0;
#pragma unroll
  for (int j = 0; j < CE_IPT; ++j) {
    const uint32_t i = base + j;
    if (i < M) {
      const bool head = (i == 0) || (k[j] != (j == 0 ? prev : k[j - 1]));
      if (head) {
        bool ok = (unsigned long long)k[j] < limit;
        if (ok && m_req > 1u) {
          const unsigned long long i2 = (unsigned long long)i + m_req - 1ull;
          ok = i2 < (unsigned long long)M && keys[i2] == k[j];
        }
        if (ok) { passbits |= 1u << j; ++cnt; }
      }
    }
  }

  uint32_t total;
  const uint32_t excl_thread = block_excl_scan_256(cnt, s_scan, &total);
  const uint32_t tile_excl = lb_exclusive_block<VX_THREADS / 32>(p.lb_cent, tile, total, p.epoch + 9u, &p.ctrl->error, s_lb);
  if (tid == 0 && tile == n_tiles - 1) p.ctrl->total_voxels = tile_excl + total;

  // per-frame voxel counts: one atomic per tile unless the tile straddles frames
  bool per_head_count = false;
  if (F == 1) {
    if (tid == 0 && total) atomicAdd(&p.acc[0].voxel_count, total);
  } else {
    const uint32_t tile_last = min(tile_base + CE_TILE, M) - 1;
    const unsigned long long f_first = (unsigned long long)keys[tile_base] >> idx_bits;
    const unsigned long long f_last = (unsigned long long)keys[tile_last] >> idx_bits;
    if (f_first == f_last) {
      if (tid == 0 && total && f_first < F) atomicAdd(&p.acc[f_first].voxel_count, total);
    } else {
      per_head_count = true;
    }
  }

  uint32_t slot = tile_excl + excl_thread;
#pragma unroll
  for (int j = 0; j < CE_IPT; ++j) {
    if (!(passbits & (1u << j))) continue;
    const KeyT key = k[j];
    uint32_t q = base + j;
    float sx = 0.f, sy = 0.f, sz = 0.f, sw = 0.f;
    uint32_t n = 0;
    do {
      const uint32_t v = __ldg(vals + q);
      const float4 pt = __ldg(p.pts + v);
      sx = __fadd_rn(sx, pt.x); sy = __fadd_rn(sy, pt.y); sz = __fadd_rn(sz, pt.z); sw = __fadd_rn(sw, pt.w);
      ++n; ++q;
    } while (q < M && keys[q] == key);
    const float nf = (float)n;
    const float cx = __fdiv_rn(sx, nf), cy = __fdiv_rn(sy, nf), cz = __fdiv_rn(sz, nf);
    const float ci = p.downsample_all ? __fdiv_rn(sw, nf) : 0.f;
    if (p.out_step == 32) {
      float4* o = reinterpret_cast<float4*>(reinterpret_cast<uint8_t*>(p.out_xyzi) + (size_t)slot * 32);
      o[0] = make_float4(cx, cy, cz, 1.0f);
      o[1] = make_float4(ci, 0.f, 0.f, 0.f);
    } else {
      reinterpret_cast<float4*>(p.out_xyzi)[slot] = make_float4(cx, cy, cz, ci);
    }
    p.out_count[slot] = n;
    p.out_idx[slot] = (unsigned long long)key & idx_mask;
    if (per_head_count) {
      const unsigned long long f = (unsigned long long)key >> idx_bits;
      if (f < F) atomicAdd(&p.acc[f].voxel_count, 1u);
    }
    ++slot;
  }
}

// ---- launchers ----------------------------------------------------------------------------------------------------
static inline uint32_t persistent_grid(uint32_t n_tiles) {
  const uint32_t cap = 148u * 8u;
  return n_tiles < 1u ? 1u : (n_tiles < cap ? n_tiles : cap);
}

cudaError_t launch_minmax(const float4* pts, uint32_t n, Ctrl* ctrl, FrameAcc* acc, uint32_t* frame_surv_start,
                          cudaStream_t stream) {
  const uint32_t tiles = (n + KH_TILE - 1) / KH_TILE;
  k_minmax<<<persistent_grid(tiles), VX_THREADS, 0, stream>>>(pts, n, ctrl, acc, frame_surv_start);
  return cudaGetLastError();
}

cudaError_t launch_grid_setup(const VoxelParams& p, cudaStream_t stream) {
  k_grid_setup<<<1, VX_THREADS, 0, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_key_hist(const VoxelParams& p, cudaStream_t stream) {
  const uint32_t tiles = (p.max_points + KH_TILE - 1) / KH_TILE;
  if (p.key_bytes == 4)
    k_voxel_key_hist<uint32_t><<<persistent_grid(tiles), VX_THREADS, 0, stream>>>(p);
  else
    k_voxel_key_hist<unsigned long long><<<persistent_grid(tiles), VX_THREADS, 0, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_centroid(const VoxelParams& p, cudaStream_t stream) {
  const uint32_t tiles = (p.max_points + CE_TILE - 1) / CE_TILE;
  if (tiles == 0) return cudaSuccess;
  if (p.key_bytes == 4)
    k_voxel_centroid<uint32_t><<<tiles, VX_THREADS, 0, stream>>>(p);
  else
    k_voxel_centroid<unsigned long long><<<tiles, VX_THREADS, 0, stream>>>(p);
  return cudaGetLastError();
}

uint32_t centroid_tile_items() { return CE_TILE; }

}  // namespace cm
