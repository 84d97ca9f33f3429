// Microbenchmark: tcgen05.ld (32x32b.x32) read bandwidth of tensor memory per SM, as a function of the number of warps.
// nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -I autoregressive_diffusion_b200/csrc -o build/tmem_bw tools/probes/tmem_bw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace ob;

__device__ __forceinline__ void ld_16_256b(uint32_t taddr, float (&v)[32]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) : "r"(taddr));
}
__device__ __forceinline__ void ld_16_128b(uint32_t taddr, float (&v)[32]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.ld.sync.aligned.16x128b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) : "r"(taddr));
}
__device__ __forceinline__ void ld_16_64b(uint32_t taddr, float (&v)[32]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.ld.sync.aligned.16x64b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) : "r"(taddr));
}
template <int SHAPE>
__global__ void __launch_bounds__(512, 1) tmem_bw_kernel(int iters, int warps_active, long long* clocks, float* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(smem_u32(&slot), 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  const uint32_t lane_off = static_cast<uint32_t>((warp & 3) * 32) << 16;
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  if (warp < warps_active) {
    for (int i = 0; i < iters; ++i) {
      float v[32];
      const uint32_t a = tmem + lane_off + ((i * 64 + (warp >> 2) * 64) & 255);
      if (SHAPE == 0) {
        float v2[32];
        tmem_ld32(a, v);
        tmem_ld32(a + 32, v2);
        tmem_ld_wait();
        acc += v2[3] + v[7];
      }
      else if (SHAPE == 1) ld_16_256b(a, v);
      else if (SHAPE == 2) ld_16_128b(a, v);
      else ld_16_64b(a, v);
      tmem_ld_wait();
      acc += v[5] + v[31];
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) clocks[blockIdx.x] = t1 - t0;
  if (acc == 123.456f) sink[0] = acc;
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
  long long* d_clk; float* d_sink;
  cudaMalloc(&d_clk, 148 * sizeof(long long)); cudaMalloc(&d_sink, 4);
  const int iters = 4096;
  for (int shape = 0; shape < 4; ++shape)
  for (int w : {1, 4, 16}) {
    auto k = shape == 0 ? tmem_bw_kernel<0> : shape == 1 ? tmem_bw_kernel<1> : shape == 2 ? tmem_bw_kernel<2> : tmem_bw_kernel<3>;
    k<<<148, 512>>>(iters, w, d_clk, d_sink);
    k<<<148, 512>>>(iters, w, d_clk, d_sink);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    long long h[148]; cudaMemcpy(h, d_clk, sizeof(h), cudaMemcpyDeviceToHost);
    const double bytes = double(w) * iters * 32 * 32 * 4 * (shape == 0 ? 2 : 1);
    printf("shape=%d warps=%2d  clocks=%lld  bytes/clk/SM=%.1f  clk per x32 load per warp=%.1f\n", shape, w, h[0], bytes / h[0], double(h[0]) / iters);
  }
  return 0;
}
