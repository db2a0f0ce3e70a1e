// cm_plane.cu -- RANSAC ground plane (pcl::SACSegmentation, SACMODEL_PLANE + SAC_RANSAC) for sm_100a.
//
// Replaces the plane search of removeGround() in the reference (pcl_preprocessing/src/pc_preprocessing_main.cpp:95-117;
// parameters Parameter.h:38-42: 1000 iterations, threshold 0.3 m, probability 0.99, optimize on). PCL 1.8.1 evaluates one
// hypothesis after the other: draw three indices, fit the plane through them, count the points within the threshold, lower
// the iteration bound k from the best count so far, stop when iterations >= k. The three-index draws come from a
// fixed-seed Mersenne twister and do not depend on the data, so the host generates the draw stream ahead, and
//
//   k_plane_score   fits and scores a whole batch of draws at once: blockIdx.y picks 128 draws (one per thread), blockIdx.x
//                   a chunk of 512 points staged in shared memory and broadcast to the 128 threads; per draw the kernel
//                   writes the sample test (isSampleGood), the model and the inlier count. The host then walks the batch
//                   in draw order with PCL's stopping rule -- the result is the model PCL's sequential loop ends on.
//   k_plane_select  selectWithinDistance: the inlier / rest flags of every point for one model (bit 0 / bit 1 of the
//                   zone-slicing mask; its scan + scatter kernels produce ground and no-ground clouds in input order, as the
//                   two pcl::ExtractIndices passes of the reference do).
//   k_plane_moments optimizeModelCoefficients' running sums (xx, xy, xz, yy, yz, zz, x, y, z over the inliers) in PCL's
//                   order: float accumulators, one add per inlier in index order. Nine lanes of one warp carry the nine
//                   serial chains while eight producer warps stage the products of the next 256 inliers (double buffer).
//
// All float arithmetic is single IEEE operations in PCL's / Eigen's order (no FMA); the order of Eigen's 4-wide packet
// reductions depends on the instruction set PCL was built for and is a parameter (PlaneParams::sum_order).
//
// Roofline: k_plane_score is FP32-issue bound (8 instructions per point x draw); everything here is microseconds per zone --
// the cost of the path is the host round trips of the stopping rule, not the kernels.
#include "cm_kernels.h"

namespace cm {

namespace {

constexpr int PL_THREADS = 128;  // draws per CTA
constexpr int PL_CHUNK = 512;    // points per CTA

template <int ORDER>
__device__ __forceinline__ float sum4(float l0, float l1, float l2, float l3) {
  if constexpr (ORDER == 0) return __fadd_rn(__fadd_rn(l0, l2), __fadd_rn(l1, l3));
  else if constexpr (ORDER == 1) return __fadd_rn(__fadd_rn(l0, l1), __fadd_rn(l2, l3));
  else return __fadd_rn(__fadd_rn(__fadd_rn(l0, l1), l2), l3);
}

template <int ORDER>
__device__ __forceinline__ float plane_abs_dist(const float4 c, const float4 q) {
  return fabsf(sum4<ORDER>(__fmul_rn(c.x, q.x), __fmul_rn(c.y, q.y), __fmul_rn(c.z, q.z), c.w));
}

template <int ORDER>
__global__ void __launch_bounds__(PL_THREADS) k_plane_score(const PlaneParams p) {
  __shared__ float4 s_pts[PL_CHUNK];
  const uint32_t base = blockIdx.x * PL_CHUNK;
  const float qnan = __int_as_float(0x7fc00000);
  for (int i = threadIdx.x; i < PL_CHUNK; i += PL_THREADS) {
    const uint32_t idx = base + i;
    s_pts[i] = idx < p.n_points ? __ldg(p.pts + idx) : make_float4(qnan, qnan, qnan, qnan);  // a NaN distance is never inside
  }
  const uint32_t h = blockIdx.y * PL_THREADS + threadIdx.x;
  bool good = false;
  float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
  if (h < p.n_draws) {
    const float4 p0 = __ldg(p.pts + p.samples[3 * h]), p1 = __ldg(p.pts + p.samples[3 * h + 1]),
                 p2 = __ldg(p.pts + p.samples[3 * h + 2]);
    const float ux = __fsub_rn(p1.x, p0.x), uy = __fsub_rn(p1.y, p0.y), uz = __fsub_rn(p1.z, p0.z);
    const float vx = __fsub_rn(p2.x, p0.x), vy = __fsub_rn(p2.y, p0.y), vz = __fsub_rn(p2.z, p0.z);
    // isSampleGood / the collinearity test of computeModelCoefficients: the quotients of the two edge vectors
    const float qx = __fdiv_rn(ux, vx), qy = __fdiv_rn(uy, vy), qz = __fdiv_rn(uz, vz);
    good = (qx != qy) || (qz != qy);
    c.x = __fsub_rn(__fmul_rn(uy, vz), __fmul_rn(uz, vy));
    c.y = __fsub_rn(__fmul_rn(uz, vx), __fmul_rn(ux, vz));
    c.z = __fsub_rn(__fmul_rn(ux, vy), __fmul_rn(uy, vx));
    c.w = 0.f;
    // Eigen 3.3 normalize(): only when the squared norm is positive
    const float z = sum4<ORDER>(__fmul_rn(c.x, c.x), __fmul_rn(c.y, c.y), __fmul_rn(c.z, c.z), __fmul_rn(c.w, c.w));
    if (z > 0.f) {
      const float nrm = __fsqrt_rn(z);
      c.x = __fdiv_rn(c.x, nrm); c.y = __fdiv_rn(c.y, nrm); c.z = __fdiv_rn(c.z, nrm); c.w = __fdiv_rn(c.w, nrm);
    }
    c.w = __fmul_rn(-1.0f, sum4<ORDER>(__fmul_rn(c.x, p0.x), __fmul_rn(c.y, p0.y), __fmul_rn(c.z, p0.z), __fmul_rn(c.w, 1.0f)));
    if (blockIdx.x == 0) {
      p.models[h] = c;
      p.good[h] = good ? 1 : 0;
    }
  }
  __syncthreads();
  if (!good) return;
  int cnt = 0;
#pragma unroll 8
  for (int i = 0; i < PL_CHUNK; ++i) cnt += plane_abs_dist<ORDER>(c, s_pts[i]) < p.threshold ? 1 : 0;
  if (cnt) atomicAdd(p.counts + h, cnt);
}

template <int ORDER>
__global__ void __launch_bounds__(256) k_plane_select(const float4* __restrict__ pts, uint32_t n, float4 c, float threshold,
                                                      unsigned short* __restrict__ mask) {
  const uint32_t i = blockIdx.x * 256u + threadIdx.x;
  if (i >= n) return;
  const bool in = plane_abs_dist<ORDER>(c, __ldg(pts + i)) < threshold;
  mask[i] = in ? 1 : 2;
}

__global__ void __launch_bounds__(256) k_plane_mask_all(uint32_t n, unsigned short v, unsigned short* __restrict__ mask) {
  const uint32_t i = blockIdx.x * 256u + threadIdx.x;
  if (i < n) mask[i] = v;
}

constexpr int PM_TILE = 256;
constexpr int PM_THREADS = 32 + PM_TILE;  // warp 0 accumulates, the others stage

__global__ void __launch_bounds__(PM_THREADS) k_plane_moments(const float4* __restrict__ inliers,
                                                             const uint32_t* __restrict__ n_inliers_dev,
                                                             float* __restrict__ out /* [9] sums + [1] count bits */) {
  __shared__ float s_term[2][9][PM_TILE + 1];  // +1: the nine chains read nine different banks
  const uint32_t n = *n_inliers_dev;
  const uint32_t n_tiles = (n + PM_TILE - 1) / PM_TILE;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float acc = 0.f;
  for (uint32_t t = 0; t <= n_tiles; ++t) {
    if (warp > 0) {
      if (t < n_tiles) {
        const int j = threadIdx.x - 32;
        const uint32_t idx = t * PM_TILE + j;
        if (idx < n) {
          const float4 q = __ldg(inliers + idx);
          float(*term)[PM_TILE + 1] = s_term[t & 1];
          term[0][j] = __fmul_rn(q.x, q.x);
          term[1][j] = __fmul_rn(q.x, q.y);
          term[2][j] = __fmul_rn(q.x, q.z);
          term[3][j] = __fmul_rn(q.y, q.y);
          term[4][j] = __fmul_rn(q.y, q.z);
          term[5][j] = __fmul_rn(q.z, q.z);
          term[6][j] = q.x;
          term[7][j] = q.y;
          term[8][j] = q.z;
        }
      }
    } else if (t > 0 && lane < 9) {
      const uint32_t first = (t - 1) * PM_TILE;
      const int cnt = (int)min((uint32_t)PM_TILE, n - first);
      const float* term = s_term[(t - 1) & 1][lane];
      int j = 0;
      for (; j + 8 <= cnt; j += 8) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = term[j + u];
#pragma unroll
        for (int u = 0; u < 8; ++u) acc = __fadd_rn(acc, v[u]);
      }
      for (; j < cnt; ++j) acc = __fadd_rn(acc, term[j]);
    }
    __syncthreads();
  }
  if (warp == 0 && lane < 9) out[lane] = acc;
  if (threadIdx.x == 0) out[9] = __uint_as_float(n);
}

}  // namespace

cudaError_t launch_plane_score(const PlaneParams& p, cudaStream_t stream) {
  if (p.n_draws == 0 || p.n_points == 0) return cudaSuccess;
  const dim3 grid((p.n_points + PL_CHUNK - 1) / PL_CHUNK, (p.n_draws + PL_THREADS - 1) / PL_THREADS);
  if (p.sum_order == 0) k_plane_score<0><<<grid, PL_THREADS, 0, stream>>>(p);
  else if (p.sum_order == 1) k_plane_score<1><<<grid, PL_THREADS, 0, stream>>>(p);
  else k_plane_score<2><<<grid, PL_THREADS, 0, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_plane_select(const float4* pts, uint32_t n, const float* coeff, float threshold, uint32_t sum_order,
                                bool found, unsigned short* mask, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  const uint32_t blocks = (n + 255u) / 256u;
  if (!found) {  // segment() failed: no inliers, every point goes to the rest
    k_plane_mask_all<<<blocks, 256, 0, stream>>>(n, 2, mask);
    return cudaGetLastError();
  }
  const float4 c = make_float4(coeff[0], coeff[1], coeff[2], coeff[3]);
  if (sum_order == 0) k_plane_select<0><<<blocks, 256, 0, stream>>>(pts, n, c, threshold, mask);
  else if (sum_order == 1) k_plane_select<1><<<blocks, 256, 0, stream>>>(pts, n, c, threshold, mask);
  else k_plane_select<2><<<blocks, 256, 0, stream>>>(pts, n, c, threshold, mask);
  return cudaGetLastError();
}

cudaError_t launch_plane_moments(const float4* inliers, const uint32_t* n_inliers_dev, float* out, cudaStream_t stream) {
  k_plane_moments<<<1, PM_THREADS, 0, stream>>>(inliers, n_inliers_dev, out);
  return cudaGetLastError();
}

}  // namespace cm
