// Micro-benchmark: tcgen05.ld (TMEM -> registers) throughput per SM, as a function of the number of
// reading warps and the load width.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
//   -I librecommender_b200/csrc tools/ubench/ldtm.cu -o tools/ubench/ldtm
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx_sm100.cuh"
using namespace b200;

template <int WIDTH>
__global__ void ldtm_kernel(int iters, unsigned long long* cycles, uint32_t* sink) {
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { ptx::tmem_alloc(&tmem_base_s, 512); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t base = tmem_base_s + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const unsigned long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const uint32_t col = (uint32_t)(((it * 4 + (warp >> 2)) * WIDTH) & 511) & ~(uint32_t)(WIDTH - 1);
    if (WIDTH == 64) {
      uint32_t r[64];
      ptx::tmem_ld_32x32b_x64(base + col, r);
      ptx::tmem_ld_wait_regs64(r);
      acc ^= r[0] ^ r[63];
    } else {
      uint32_t r[32];
      ptx::tmem_ld_32x32b_x32(base + col, r);
      ptx::tmem_ld_wait_regs(r);
      acc ^= r[0] ^ r[31];
    }
  }
  __syncthreads();
  const unsigned long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  __syncthreads();
  if (warp == 0) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem_base_s, 512); }
}

int main() {
  unsigned long long* d_cyc; uint32_t* d_sink;
  cudaMalloc(&d_cyc, 148 * 8); cudaMalloc(&d_sink, 148 * 1024 * 4);
  const int iters = 4000;
  for (int width : {32, 64}) {
    for (int warps : {4, 8, 16}) {
      for (int rep = 0; rep < 2; ++rep) {
        if (width == 64) ldtm_kernel<64><<<148, warps * 32>>>(iters, d_cyc, d_sink);
        else ldtm_kernel<32><<<148, warps * 32>>>(iters, d_cyc, d_sink);
        cudaDeviceSynchronize();
      }
      unsigned long long h[148];
      cudaMemcpy(h, d_cyc, sizeof(h), cudaMemcpyDeviceToHost);
      double avg = 0; for (int i = 0; i < 148; ++i) avg += (double)h[i]; avg /= 148;
      const double bytes = (double)iters * warps * 32 * width * 4;
      printf("{\"ubench\": \"tcgen05.ld.32x32b.x%d\", \"warps\": %d, \"cycles\": %.0f, \"bytes_per_clk_per_sm\": %.1f, \"err\": \"%s\"}\n",
             width, warps, avg, bytes / avg, cudaGetErrorString(cudaGetLastError()));
    }
  }
  return 0;
}
