// cm_common.cuh -- device-side structures and helpers shared by the sm_100a kernels of the merge hot path.
//
// Conventions
//  * one "run" = one batch of F frames x S sensors = n_seg segments, processed by a fixed sequence of launches;
//  * every kernel that needs an order-preserving prefix across CTAs uses a single-pass decoupled look-back over
//    tiles whose ids are handed out by an atomic counter (forward progress does not depend on CTA dispatch order);
//  * look-back words carry an epoch, so the state arrays never need clearing between runs;
//  * all spin loops have a watchdog: a stuck wait raises CM_E_INTERNAL in the control block instead of hanging the GPU.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#define CM_RADIX_BITS 8
#define CM_RADIX 256
#define CM_MAX_SORT_PASSES 8

#define CM_DEV_OK 0
#define CM_DEV_E_KEY_RANGE 5
#define CM_DEV_E_INTERNAL 6

namespace cm {

// ---- one sensor cloud of one frame, as the kernels see it ------------------------------------------------------
enum SegMode : uint32_t {
  SEG_PACKED16 = 0,  // point_step 16, x y z i at 0 4 8 12: one 16-byte load per point
  SEG_PCL32 = 1,     // pcl::PointXYZI / Velodyne-Melodic: point_step 32, xyz at 0 4 8, intensity at 16
  SEG_ALIGNED4 = 2,  // point_step % 4 == 0 and offsets % 4 == 0: four 4-byte loads
  SEG_STAGED = 3,    // anything else up to CM_MAX_STAGED_STEP: tile bytes staged to shared memory by cp.async.bulk
  SEG_BYTES = 4      // last resort: byte loads from global memory
};
#define CM_MAX_STAGED_STEP 96

struct SegDev {
  const uint8_t* data;
  uint32_t n_points;
  uint32_t tile_begin;      // first K1 tile of this segment (every segment owns >= 1 tile)
  uint32_t src_base;        // index of this segment's point 0 in its frame's un-cropped concatenation
  uint32_t frame;
  int32_t point_step;
  int32_t off_x, off_y, off_z, off_i;
  uint32_t is_dense;
  uint32_t sensor;
  uint32_t first_of_frame;
  uint32_t mode;
  uint32_t pad_;
};

struct PassDev {
  int32_t axis;
  float lo, hi;
  int32_t negative;
};

struct CropDev {
  int32_t n_pass;
  PassDev pass[8];
};

// ---- per-run control block; zeroed by one memset at the start of every run --------------------------------------
struct FrameAcc {
  uint32_t max_enc[3];   // atomicMax of enc(v)         -> max_p
  uint32_t nmin_enc[3];  // atomicMax of ~enc(v)        -> min_p
  uint32_t n_invalid;    // survivors with a non-finite coordinate (only possible when no crop pass is configured)
  uint32_t voxel_count;  // voxels emitted for this frame
};

struct Ctrl {
  uint32_t tile_counter[12];  // [0] transform_crop, [1..8] sort passes, [9] centroid, [10] minmax
  uint32_t error;             // CM_DEV_E_*
  uint32_t total_voxels;
  uint32_t has_invalid;
  uint32_t pad_;
};

// ---- written by k_grid_setup -----------------------------------------------------------------------------------
struct GridDev {
  int32_t min_b[3];
  int32_t max_b[3];
  int32_t div_b[3];
  int32_t pcl_overflow;
  uint32_t bits;
  uint32_t empty;
  unsigned long long mul1, mul2;  // div0, div0*div1
};

struct SortInfo {
  uint32_t num_passes;
  uint32_t total_bits;
  uint32_t idx_bits;
  uint32_t n_keys;      // = survivors
  uint32_t key_frames;  // F (+1 when a sentinel frame is needed for invalid points)
  uint32_t pad_[3];
};

// ---- float <-> order-preserving uint -----------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t f32_order_enc(uint32_t bits) {
  return bits ^ ((bits >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}
__host__ __device__ __forceinline__ uint32_t f32_order_dec(uint32_t enc) {
  return enc ^ ((enc >> 31) ? 0x80000000u : 0xFFFFFFFFu);
}

#ifdef __CUDACC__

__device__ __forceinline__ bool finite_f32(float v) { return (__float_as_uint(v) & 0x7F800000u) != 0x7F800000u; }

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ uint32_t lanemask_lt() {
  uint32_t m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

__device__ __forceinline__ uint32_t warp_sum_u32(uint32_t v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
  return v;
}
__device__ __forceinline__ uint32_t warp_incl_scan_u32(uint32_t v) {
  const uint32_t lane = lane_id();
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xFFFFFFFFu, v, o);
    if (lane >= (uint32_t)o) v += t;
  }
  return v;
}

// streaming 16-byte accesses that do not pollute L1
__device__ __forceinline__ float4 ldg_stream_f4(const void* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ float ldg_stream_f1(const void* p) {
  float r;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
  return r;
}

// ---- decoupled look-back --------------------------------------------------------------------------------------
// word = ((epoch * 4 + flag) << 32) | value ; flag 1 = tile aggregate published, 2 = inclusive prefix published.
#define CM_LB_AGG 1u
#define CM_LB_INCL 2u
#define CM_SPIN_LIMIT (1u << 22)

__device__ __forceinline__ unsigned long long lb_pack(uint32_t epoch, uint32_t flag, uint32_t value) {
  return ((unsigned long long)(epoch * 4u + flag) << 32) | (unsigned long long)value;
}

// Spin until the word belongs to this epoch and carries a flag. Returns the word; on watchdog expiry raises the
// device error and returns an "inclusive 0" word so that every waiter drains.
__device__ __forceinline__ unsigned long long lb_wait(volatile unsigned long long* p, uint32_t epoch, uint32_t* err) {
  uint32_t spins = 0;
  while (true) {
    unsigned long long w = *p;
    uint32_t hi = (uint32_t)(w >> 32);
    if ((hi >> 2) == epoch && (hi & 3u) != 0u) return w;
    if (++spins > CM_SPIN_LIMIT) {
      atomicExch(err, (uint32_t)CM_DEV_E_INTERNAL);
      return lb_pack(epoch, CM_LB_INCL, 0u);
    }
    __nanosleep(40);
  }
}

// Exclusive prefix of `agg` over all tiles before `tile`. Must be called by one full warp (all 32 lanes);
// every lane returns the prefix.
__device__ __forceinline__ uint32_t lb_exclusive_warp(volatile unsigned long long* st, uint32_t tile, uint32_t agg,
                                                      uint32_t epoch, uint32_t* err) {
  const uint32_t lane = lane_id();
  if (tile == 0) {
    if (lane == 0) st[0] = lb_pack(epoch, CM_LB_INCL, agg);
    return 0u;
  }
  if (lane == 0) st[tile] = lb_pack(epoch, CM_LB_AGG, agg);
  uint32_t excl = 0;
  long long base = (long long)tile - 1;
  while (true) {
    const long long idx = base - (long long)lane;
    uint32_t flag = CM_LB_INCL, val = 0;
    if (idx >= 0) {
      const unsigned long long w = lb_wait(st + idx, epoch, err);
      flag = ((uint32_t)(w >> 32)) & 3u;
      val = (uint32_t)w;
    }
    const uint32_t incl = __ballot_sync(0xFFFFFFFFu, flag == CM_LB_INCL);
    const int first = incl ? (__ffs(incl) - 1) : 32;
    excl += warp_sum_u32(((int)lane <= first) ? val : 0u);
    if (incl) break;
    base -= 32;
  }
  if (lane == 0) st[tile] = lb_pack(epoch, CM_LB_INCL, excl + agg);
  return excl;
}

// Per-thread variant used by the radix pass: thread d walks back over tiles for its own digit d.
__device__ __forceinline__ uint32_t lb_exclusive_digit(volatile unsigned long long* st, uint32_t tile, uint32_t d,
                                                       uint32_t agg, uint32_t epoch, uint32_t* err) {
  volatile unsigned long long* mine = st + (size_t)tile * CM_RADIX + d;
  if (tile == 0) {
    *mine = lb_pack(epoch, CM_LB_INCL, agg);
    return 0u;
  }
  *mine = lb_pack(epoch, CM_LB_AGG, agg);
  uint32_t excl = 0;
  for (long long j = (long long)tile - 1; j >= 0; --j) {
    const unsigned long long w = lb_wait(st + (size_t)j * CM_RADIX + d, epoch, err);
    excl += (uint32_t)w;
    if ((((uint32_t)(w >> 32)) & 3u) == CM_LB_INCL) break;
  }
  *mine = lb_pack(epoch, CM_LB_INCL, excl + agg);
  return excl;
}

// Exclusive scan over 256 values held one per thread by threads 0..255 of a block with >= 256 threads.
// `scratch` is >= 9 uint32 of shared memory. Every thread of the block must call it; returns the exclusive prefix
// (for threads >= 256 the return value is meaningless). *total receives the sum (valid in all threads).
__device__ __forceinline__ uint32_t block_excl_scan_256(uint32_t v, uint32_t* scratch, uint32_t* total) {
  const uint32_t tid = threadIdx.x, lane = tid & 31u, w = tid >> 5;
  if (tid >= 256) v = 0;
  const uint32_t incl = warp_incl_scan_u32(v);
  if (lane == 31 && w < 8) scratch[w] = incl;
  __syncthreads();
  if (tid == 0) {
    uint32_t run = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      uint32_t t = scratch[i];
      scratch[i] = run;
      run += t;
    }
    scratch[8] = run;
  }
  __syncthreads();
  const uint32_t res = incl - v + (w < 8 ? scratch[w] : 0u);
  *total = scratch[8];
  __syncthreads();
  return res;
}

#endif  // __CUDACC__

}  // namespace cm
