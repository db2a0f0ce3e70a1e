// micro-benchmark: cycles per dependent FADD for one warp (9 active lanes), operands from registers / shared memory
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(float* out, long long* cyc, int reps, int active_lanes) {
  __shared__ __align__(16) float s[9][260];
  for (int i = threadIdx.x; i < 9 * 260; i += blockDim.x) (&s[0][0])[i] = 1.0f + i * 1e-6f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  if (threadIdx.x >= 32 || lane >= active_lanes) return;
  float acc = 0.f;
  // (a) register chain
  float v[16];
  for (int u = 0; u < 16; ++u) v[u] = s[lane % 9][u];
  long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
#pragma unroll
    for (int u = 0; u < 16; ++u) acc = __fadd_rn(acc, v[u]);
  }
  long long t1 = clock64();
  // (b) LDS.128 ping-pong like the product kernel
  const float* row = s[lane % 9];
  float va[16], vb[16];
  auto ld16 = [&](float (&w)[16], const float* p) {
#pragma unroll
    for (int u = 0; u < 4; ++u) { float4 x = *reinterpret_cast<const float4*>(p + 4 * u); w[4*u]=x.x; w[4*u+1]=x.y; w[4*u+2]=x.z; w[4*u+3]=x.w; }
  };
  ld16(va, row);
  long long t2 = clock64();
#pragma unroll 1
  for (int g = 0; g < reps; g += 2) {
    ld16(vb, row + ((g + 1) & 15) * 16);
#pragma unroll
    for (int u = 0; u < 16; ++u) acc = __fadd_rn(acc, va[u]);
    ld16(va, row + ((g + 2) & 15) * 16);
#pragma unroll
    for (int u = 0; u < 16; ++u) acc = __fadd_rn(acc, vb[u]);
  }
  long long t3 = clock64();
  if (lane == 0) { cyc[0] = t1 - t0; cyc[1] = t3 - t2; }
  out[lane] = acc;
}
int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 128); cudaMalloc(&cyc, 16);
  for (int threads : {32, 288}) for (int lanes : {9, 32}) {
    const int reps = 4096;
    k<<<1, threads>>>(out, cyc, reps, lanes);
    k<<<1, threads>>>(out, cyc, reps, lanes);
    long long h[2]; cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost);
    printf("threads %d lanes %d: register chain %.2f cycles/add, LDS.128 ping-pong %.2f cycles/add\n", threads, lanes,
           (double)h[0] / (reps * 16.0), (double)h[1] / (reps * 16.0));
  }
  return 0;
}
