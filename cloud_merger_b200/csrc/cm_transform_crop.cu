// cm_transform_crop.cu -- K1: fused PointCloud2 unpack + rigid transform + PassThrough crop + stable stream compaction
// (+ per-frame bounding box of the survivors) for sm_100a.
//
// Replaces, in one pass over the raw bytes (reference paths relative to timspilak/cloud_merger):
//   * the pcl_ros subscriber deserialisation            pc_preprocessing_main.cpp:520-525
//   * pcl_ros::transformPointCloud                      pc_preprocessing_main.cpp:322,348,373,399,426,465
//   * the pcl::PassThrough chains getROI/getCloudPart   pc_preprocessing_main.cpp:20-59 (and CloudFusionNode.h:145-216)
//   * the operator+= concatenation in fusePointclouds   pc_preprocessing_main.cpp:137-149
//   * pcl::getMinMax3D, the first step of VoxelGrid     (PCL 1.8.1 voxel_grid.hpp)
//
// Roofline: HBM. Algorithmic bytes per launch = sum(n_points * point_step) + survivors * (16 + 4).
// Arithmetic is kept bit-identical to the PCL 1.8.1 CPU build: x' = ((m00*x + m01*y) + m02*z) + m03 with every
// multiply and add rounded separately (__fmul_rn/__fadd_rn are never contracted into FMA).
#include "cm_kernels.h"

namespace cm {

namespace {

constexpr int K1_THREADS = 256;
constexpr int K1_IPT = 8;
constexpr int K1_TILE = K1_THREADS * K1_IPT;  // 2048 points
constexpr int K1_WARPS = K1_THREADS / 32;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// 1-D TMA bulk copy global -> shared, completion signalled on the mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}

__device__ __forceinline__ float lds_f32_unaligned(const uint8_t* s, uint32_t byte_off) {
  const uint32_t a = byte_off & ~3u, sh = (byte_off & 3u) * 8u;
  const uint32_t w0 = *reinterpret_cast<const uint32_t*>(s + a);
  const uint32_t w1 = *reinterpret_cast<const uint32_t*>(s + a + 4);
  return __uint_as_float(__funnelshift_r(w0, w1, sh));
}
__device__ __forceinline__ float ldg_f32_bytes(const uint8_t* p) {
  const uint32_t v = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
  return __uint_as_float(v);
}

// One pcl::PassThrough stage on an already finite point (PCL 1.8.1 applyFilterIndices).
__device__ __forceinline__ bool pass_keeps(const PassDev& ps, float x, float y, float z, float it) {
  const float v = ps.axis == 0 ? x : (ps.axis == 1 ? y : (ps.axis == 2 ? z : it));
  if (!finite_f32(v)) return false;
  if (!ps.negative) return !(v < ps.lo || v > ps.hi);
  return !(v >= ps.lo && v <= ps.hi);
}

}  // namespace

__global__ void __launch_bounds__(K1_THREADS) k_transform_crop(const K1Params p) {
  extern __shared__ __align__(16) uint8_t stage[];
  __shared__ uint32_t s_tile, s_seg, s_tile_excl;
  __shared__ uint32_t s_warp_tot[K1_WARPS], s_warp_inv[K1_WARPS];
  __shared__ float s_mm[K1_WARPS][6];
  __shared__ __align__(8) unsigned long long s_bar;

  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;

  if (tid == 0) s_tile = atomicAdd(&p.ctrl->tile_counter[0], 1u);
  __syncthreads();
  const uint32_t tile = s_tile;

  // which segment owns this tile (segments are few: one parallel probe)
  for (uint32_t s = tid; s < p.n_seg; s += K1_THREADS) {
    const uint32_t tb = p.segs[s].tile_begin;
    const uint32_t te = (s + 1 < p.n_seg) ? p.segs[s + 1].tile_begin : p.n_tiles;
    if (tb <= tile && tile < te) s_seg = s;
  }
  __syncthreads();
  const uint32_t seg_id = s_seg;
  const SegDev sg = p.segs[seg_id];
  const uint32_t pt0 = (tile - sg.tile_begin) * K1_TILE;
  const uint32_t n_here = sg.n_points > pt0 ? min((uint32_t)K1_TILE, sg.n_points - pt0) : 0u;

  float m[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) m[k] = __ldg(p.mats + sg.sensor * 12 + k);

  float x[K1_IPT], y[K1_IPT], z[K1_IPT], it[K1_IPT];
  const uint32_t li0 = warp * (32 * K1_IPT) + lane;  // local index of item 0; item i is li0 + 32*i

  // ---- unpack -------------------------------------------------------------------------------------------------
  if (sg.mode == SEG_PACKED16) {
    const uint8_t* base = sg.data + (size_t)pt0 * 16;
#pragma unroll
    for (int i = 0; i < K1_IPT; ++i) {
      const uint32_t li = li0 + 32 * i;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (li < n_here) v = ldg_stream_f4(base + (size_t)li * 16);
      x[i] = v.x; y[i] = v.y; z[i] = v.z; it[i] = v.w;
    }
  } else if (sg.mode == SEG_PCL32) {
    const uint8_t* base = sg.data + (size_t)pt0 * 32;
#pragma unroll
    for (int i = 0; i < K1_IPT; ++i) {
      const uint32_t li = li0 + 32 * i;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      float w = 0.f;
      if (li < n_here) {
        v = ldg_stream_f4(base + (size_t)li * 32);
        w = ldg_stream_f1(base + (size_t)li * 32 + 16);
      }
      x[i] = v.x; y[i] = v.y; z[i] = v.z; it[i] = w;
    }
  } else if (sg.mode == SEG_ALIGNED4) {
    const uint8_t* base = sg.data + (size_t)pt0 * sg.point_step;
#pragma unroll
    for (int i = 0; i < K1_IPT; ++i) {
      const uint32_t li = li0 + 32 * i;
      x[i] = y[i] = z[i] = it[i] = 0.f;
      if (li < n_here) {
        const uint8_t* r = base + (size_t)li * sg.point_step;
        x[i] = ldg_stream_f1(r + sg.off_x);
        y[i] = ldg_stream_f1(r + sg.off_y);
        z[i] = ldg_stream_f1(r + sg.off_z);
        if (sg.off_i >= 0) it[i] = ldg_stream_f1(r + sg.off_i);
      }
    }
  } else if (sg.mode == SEG_STAGED) {
    // raw tile bytes -> shared memory with one TMA bulk copy (16-byte multiple) + a < 16-byte tail by plain loads
    const uint32_t bytes = n_here * (uint32_t)sg.point_step;
    const uint32_t bulk = bytes & ~15u;
    const uint8_t* src = sg.data + (size_t)pt0 * sg.point_step;
    if (tid == 0) mbar_init(&s_bar, 1);
    __syncthreads();
    if (tid == 0 && bulk) {
      mbar_arrive_expect_tx(&s_bar, bulk);
      bulk_g2s(stage, src, bulk, &s_bar);
    }
    for (uint32_t b = bulk + tid; b < bytes; b += K1_THREADS) stage[b] = src[b];
    if (bulk) {
      uint32_t spins = 0;
      while (!mbar_try_wait(&s_bar, 0)) {
        if (++spins > CM_SPIN_LIMIT) {
          atomicExch(&p.ctrl->error, (uint32_t)CM_DEV_E_INTERNAL);
          break;
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < K1_IPT; ++i) {
      const uint32_t li = li0 + 32 * i;
      x[i] = y[i] = z[i] = it[i] = 0.f;
      if (li < n_here) {
        const uint32_t r = li * (uint32_t)sg.point_step;
        x[i] = lds_f32_unaligned(stage, r + sg.off_x);
        y[i] = lds_f32_unaligned(stage, r + sg.off_y);
        z[i] = lds_f32_unaligned(stage, r + sg.off_z);
        if (sg.off_i >= 0) it[i] = lds_f32_unaligned(stage, r + sg.off_i);
      }
    }
  } else {
    const uint8_t* base = sg.data + (size_t)pt0 * sg.point_step;
#pragma unroll
    for (int i = 0; i < K1_IPT; ++i) {
      const uint32_t li = li0 + 32 * i;
      x[i] = y[i] = z[i] = it[i] = 0.f;
      if (li < n_here) {
        const uint8_t* r = base + (size_t)li * sg.point_step;
        x[i] = ldg_f32_bytes(r + sg.off_x);
        y[i] = ldg_f32_bytes(r + sg.off_y);
        z[i] = ldg_f32_bytes(r + sg.off_z);
        if (sg.off_i >= 0) it[i] = ldg_f32_bytes(r + sg.off_i);
      }
    }
  }

  // ---- transform + crop predicate + in-warp ranks ------------------------------------------------------------------
  uint32_t keep_bits = 0;            // bit i: item i survives
  uint32_t rank_in_warp[K1_IPT];     // exclusive rank of item i among the warp's survivors
  uint32_t warp_run = 0, inv_run = 0;
  float mn0 = 3.402823466e+38f, mn1 = mn0, mn2 = mn0, mx0 = -mn0, mx1 = -mn0, mx2 = -mn0;
  const int n_pass = p.crop.n_pass;
#pragma unroll
  for (int i = 0; i < K1_IPT; ++i) {
    const uint32_t li = li0 + 32 * i;
    const bool in_range = li < n_here;
    const bool fin_in = finite_f32(x[i]) && finite_f32(y[i]) && finite_f32(z[i]);
    if (sg.is_dense || fin_in) {
      const float a = x[i], b = y[i], c = z[i];
      x[i] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m[0], a), __fmul_rn(m[1], b)), __fmul_rn(m[2], c)), m[3]);
      y[i] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m[4], a), __fmul_rn(m[5], b)), __fmul_rn(m[6], c)), m[7]);
      z[i] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m[8], a), __fmul_rn(m[9], b)), __fmul_rn(m[10], c)), m[11]);
    }
    const bool fin = finite_f32(x[i]) && finite_f32(y[i]) && finite_f32(z[i]);
    bool keep = in_range;
    if (n_pass > 0) {
      keep = keep && fin;
      for (int k = 0; k < n_pass; ++k) keep = keep && pass_keeps(p.crop.pass[k], x[i], y[i], z[i], it[i]);
    }
    if (keep && fin) {
      mn0 = fminf(mn0, x[i]); mx0 = fmaxf(mx0, x[i]);
      mn1 = fminf(mn1, y[i]); mx1 = fmaxf(mx1, y[i]);
      mn2 = fminf(mn2, z[i]); mx2 = fmaxf(mx2, z[i]);
    }
    const uint32_t ballot = __ballot_sync(0xFFFFFFFFu, keep);
    const uint32_t inv_ballot = __ballot_sync(0xFFFFFFFFu, keep && !fin);
    rank_in_warp[i] = warp_run + __popc(ballot & lanemask_lt());
    warp_run += __popc(ballot);
    inv_run += __popc(inv_ballot);
    keep_bits |= (keep ? 1u : 0u) << i;
  }

  // ---- bounding box of the survivors (pcl::getMinMax3D) ----------------------------------------------------------------
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn0 = fminf(mn0, __shfl_xor_sync(0xFFFFFFFFu, mn0, o)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xFFFFFFFFu, mx0, o));
    mn1 = fminf(mn1, __shfl_xor_sync(0xFFFFFFFFu, mn1, o)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xFFFFFFFFu, mx1, o));
    mn2 = fminf(mn2, __shfl_xor_sync(0xFFFFFFFFu, mn2, o)); mx2 = fmaxf(mx2, __shfl_xor_sync(0xFFFFFFFFu, mx2, o));
  }
  if (lane == 0) {
    s_warp_tot[warp] = warp_run;
    s_warp_inv[warp] = inv_run;
    s_mm[warp][0] = mn0; s_mm[warp][1] = mn1; s_mm[warp][2] = mn2;
    s_mm[warp][3] = mx0; s_mm[warp][4] = mx1; s_mm[warp][5] = mx2;
  }
  __syncthreads();

  // ---- tile prefix by decoupled look-back (warp 0) ---------------------------------------------------------------------
  if (warp == 0) {
    uint32_t tot = 0, inv = 0;
#pragma unroll
    for (int w = 0; w < K1_WARPS; ++w) { tot += s_warp_tot[w]; inv += s_warp_inv[w]; }
    const uint32_t excl = lb_exclusive_warp(p.lb, tile, tot, p.epoch, &p.ctrl->error);
    if (lane == 0) {
      s_tile_excl = excl;
      if (tile == sg.tile_begin) {
        p.seg_surv_start[seg_id] = excl;
        if (sg.first_of_frame) p.frame_surv_start[sg.frame] = excl;
      }
      if (tile == p.n_tiles - 1) p.frame_surv_start[p.n_frames] = excl + tot;
      if (inv) {
        atomicAdd(&p.acc[sg.frame].n_invalid, inv);
        atomicOr(&p.ctrl->has_invalid, 1u);
      }
      if (tot > inv) {  // at least one finite survivor: fold the tile's box into the frame's
        float a0 = s_mm[0][0], a1 = s_mm[0][1], a2 = s_mm[0][2], b0 = s_mm[0][3], b1 = s_mm[0][4], b2 = s_mm[0][5];
#pragma unroll
        for (int w = 1; w < K1_WARPS; ++w) {
          a0 = fminf(a0, s_mm[w][0]); a1 = fminf(a1, s_mm[w][1]); a2 = fminf(a2, s_mm[w][2]);
          b0 = fmaxf(b0, s_mm[w][3]); b1 = fmaxf(b1, s_mm[w][4]); b2 = fmaxf(b2, s_mm[w][5]);
        }
        FrameAcc* fa = p.acc + sg.frame;
        atomicMax(&fa->nmin_enc[0], ~f32_order_enc(__float_as_uint(a0)));
        atomicMax(&fa->nmin_enc[1], ~f32_order_enc(__float_as_uint(a1)));
        atomicMax(&fa->nmin_enc[2], ~f32_order_enc(__float_as_uint(a2)));
        atomicMax(&fa->max_enc[0], f32_order_enc(__float_as_uint(b0)));
        atomicMax(&fa->max_enc[1], f32_order_enc(__float_as_uint(b1)));
        atomicMax(&fa->max_enc[2], f32_order_enc(__float_as_uint(b2)));
      }
    }
  }
  __syncthreads();

  // ---- write survivors at their final, order-preserving position ----------------------------------------------------
  uint32_t warp_excl = s_tile_excl;
  for (uint32_t w = 0; w < warp; ++w) warp_excl += s_warp_tot[w];
  const uint32_t src0 = sg.src_base + pt0;
#pragma unroll
  for (int i = 0; i < K1_IPT; ++i) {
    if (keep_bits & (1u << i)) {
      const uint32_t pos = warp_excl + rank_in_warp[i];
      p.surv_xyzi[pos] = make_float4(x[i], y[i], z[i], it[i]);
      if (p.surv_src) p.surv_src[pos] = src0 + li0 + 32 * i;
    }
  }
}

uint32_t k1_tile_points() { return K1_TILE; }

uint32_t k1_max_staged_smem() { return (uint32_t)K1_TILE * CM_MAX_STAGED_STEP + 16u; }

cudaError_t configure_device_kernels() {
  return cudaFuncSetAttribute(k_transform_crop, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k1_max_staged_smem());
}

cudaError_t launch_transform_crop(const K1Params& p, uint32_t staged_smem_bytes, cudaStream_t stream) {
  if (p.n_tiles == 0) return cudaSuccess;
  k_transform_crop<<<p.n_tiles, K1_THREADS, staged_smem_bytes, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace cm
