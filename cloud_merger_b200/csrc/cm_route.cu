// cm_route.cu -- device side of the single-giant-cloud mode (BASELINE config 4): one cloud block-distributed over the GPUs of
// a box, VoxelGrid with ONE all-to-all that moves every point to the rank owning its voxel-key range.
//
// The reference has no counterpart (one process, one cloud): what is reproduced is the voxel index of PCL 1.8.1
// VoxelGrid::applyFilter (idx = i + j*div_x + k*div_x*div_y with ijk = floor(p * inv_leaf) - min_b, float32 multiply) on the
// GLOBAL grid, because a voxel must not straddle ranks and the concatenation of the rank outputs must be the PCL order.
//   k_route_hist : coarse histogram of the voxel indices (bins of equal key width) -> all-reduced by the host layer to
//                  pick balanced splitters
//   k_route_mask : destination rank of every point = number of splitters <= its voxel index, as a one-hot mask for the
//                  zone-slicing count / scan / scatter kernels (cm_zones.cu), which group the points by destination in
//                  source order -- the send buffer of the all-to-all.
// Roofline: HBM, 16 B read per point and kernel (+ 2 B mask written).
#include "cm_kernels.h"

namespace cm {

namespace {

constexpr int RT_THREADS = 256;

__device__ __forceinline__ bool route_key(const RouteGrid& g, const float4& v, unsigned long long* key) {
  if (!(finite_f32(v.x) && finite_f32(v.y) && finite_f32(v.z))) return false;
  const long long i0 = (long long)__float2int_rd(__fmul_rn(v.x, g.inv[0])) - g.min_b[0];
  const long long i1 = (long long)__float2int_rd(__fmul_rn(v.y, g.inv[1])) - g.min_b[1];
  const long long i2 = (long long)__float2int_rd(__fmul_rn(v.z, g.inv[2])) - g.min_b[2];
  *key = (unsigned long long)(i0 + i1 * g.div0 + i2 * g.div01);
  return true;
}

// Up to RT_SMEM_BINS bins the histogram is privatised per CTA in shared memory (hardware-aggregated increments) and
// flushed once: a map cloud puts most of its points on a few planes, and 50 M global atomics on those hot bins took
// 4.7 ms; above that size the counters are updated in global memory directly.
constexpr uint32_t RT_SMEM_BINS = 16384;
constexpr int RT_HIST_THREADS = 512;

__global__ void __launch_bounds__(RT_HIST_THREADS) k_route_hist(const float4* __restrict__ pts, uint32_t n, const RouteGrid g,
                                                                unsigned long long width, uint32_t bins,
                                                                unsigned long long* __restrict__ hist, int use_smem) {
  extern __shared__ uint32_t s_bins[];
  if (use_smem) {
    for (uint32_t b = threadIdx.x; b < bins; b += RT_HIST_THREADS) s_bins[b] = 0;
    __syncthreads();
  }
  for (uint32_t i = blockIdx.x * RT_HIST_THREADS + threadIdx.x; i < n; i += gridDim.x * RT_HIST_THREADS) {
    unsigned long long key;
    if (route_key(g, ldg_stream_f4(pts + i), &key)) {
      unsigned long long b = key / width;
      if (b >= bins) b = bins - 1;
      if (use_smem) atomicAdd(&s_bins[(uint32_t)b], 1u);
      else atomicAdd(hist + b, 1ull);
    }
  }
  if (use_smem) {
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < bins; b += RT_HIST_THREADS) {
      const uint32_t c = s_bins[b];
      if (c) atomicAdd(hist + b, (unsigned long long)c);
    }
  }
}

__global__ void __launch_bounds__(RT_THREADS) k_route_mask(const float4* __restrict__ pts, uint32_t n, const RouteGrid g,
                                                           const RouteSplit sp, unsigned short* __restrict__ mask) {
  for (uint32_t i = blockIdx.x * RT_THREADS + threadIdx.x; i < n; i += gridDim.x * RT_THREADS) {
    unsigned long long key;
    uint32_t dest = sp.invalid_part;  // non-finite points stay where they are (VoxelGrid skips them)
    if (route_key(g, ldg_stream_f4(pts + i), &key)) {
      dest = 0;
      for (int k = 0; k < sp.n_parts - 1; ++k) dest += (key >= sp.splitter[k]) ? 1u : 0u;
    }
    mask[i] = (unsigned short)(1u << dest);
  }
}

}  // namespace

cudaError_t launch_route_hist(const float4* pts, uint32_t n, const RouteGrid& g, unsigned long long width, uint32_t bins,
                              unsigned long long* hist, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  const int use_smem = bins <= RT_SMEM_BINS ? 1 : 0;
  const size_t smem = use_smem ? (size_t)bins * sizeof(uint32_t) : 0;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(k_route_hist, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(RT_SMEM_BINS * sizeof(uint32_t)));
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const uint32_t blocks = std::min<uint32_t>((n + RT_HIST_THREADS - 1) / RT_HIST_THREADS, 148u * 2u);
  k_route_hist<<<blocks, RT_HIST_THREADS, smem, stream>>>(pts, n, g, width, bins, hist, use_smem);
  return cudaGetLastError();
}

cudaError_t launch_route_mask(const float4* pts, uint32_t n, const RouteGrid& g, const RouteSplit& sp, unsigned short* mask,
                              cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  const uint32_t blocks = std::min<uint32_t>((n + RT_THREADS - 1) / RT_THREADS, 148u * 8u);
  k_route_mask<<<blocks, RT_THREADS, 0, stream>>>(pts, n, g, sp, mask);
  return cudaGetLastError();
}

}  // namespace cm
