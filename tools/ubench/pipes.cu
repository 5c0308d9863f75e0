// Which issue pipe do the max instructions use, and at what rate?  (sm_100a micro-benchmark)
// Each kernel runs ITER iterations of 16 INDEPENDENT dependency chains per thread, on `warps` warps of
// ONE SM sub-partition group (block = 128 * wps threads -> wps warps per scheduler), and reports
// warp-instructions per cycle per scheduler.  Mixed kernels interleave two instruction kinds: if the
// mixed rate is about the SUM of the two single rates the kinds issue to different pipes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/pipes tools/ubench/pipes.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ITER = 2048;
constexpr int CH = 16;

template <int KIND>
__global__ void k(uint32_t* out, unsigned long long* cyc, uint32_t seed) {
  uint32_t a[CH], b[CH];
#pragma unroll
  for (int i = 0; i < CH; ++i) { a[i] = seed + i * 7919u + threadIdx.x; b[i] = seed * 3u + i; }
  __syncthreads();
  const unsigned long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      if (KIND == 0) {          // fp32 3-input max (FMNMX3)
        asm volatile("max.f32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(b[(i + 1) % CH]));
      } else if (KIND == 1) {   // packed half2 max (HMNMX2)
        asm volatile("max.f16x2 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i]));
      } else if (KIND == 2) {   // mixed: even chains fp32 max, odd chains half2 max
        if (i & 1) asm volatile("max.f16x2 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i]));
        else asm volatile("max.f32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(b[(i + 1) % CH]));
      } else if (KIND == 3) {   // fp32 2-input max (FMNMX)
        asm volatile("max.f32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i]));
      } else if (KIND == 4) {   // FFMA (fma pipe reference)
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(b[(i + 1) % CH]));
      } else if (KIND == 5) {   // mixed: fp32 max + FFMA
        if (i & 1) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(b[(i + 1) % CH]));
        else asm volatile("max.f32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(b[(i + 1) % CH]));
      } else if (KIND == 6) {   // mixed: half2 max + FFMA
        if (i & 1) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(b[(i + 1) % CH]));
        else asm volatile("max.f16x2 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i]));
      } else if (KIND == 7) {   // signed int 3-input max (VIMNMX3)
        asm volatile("max.s32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i]));
      } else if (KIND == 8) {   // packed s16x2 max
        asm volatile("max.s16x2 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i]));
      } else if (KIND == 9) {   // mixed: fp32 max + s16x2 max
        if (i & 1) asm volatile("max.s16x2 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i]));
        else asm volatile("max.f32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(b[(i + 1) % CH]));
      } else if (KIND == 10) {  // half2 compare-and-set (HSET2 / HSETP2)
        asm volatile("set.ge.u32.f16x2 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i]));
      }
    }
  }
  const unsigned long long t1 = clock64();
  uint32_t x = 0;
#pragma unroll
  for (int i = 0; i < CH; ++i) x ^= a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int KIND>
void run(const char* name) {
  uint32_t* out; unsigned long long* cyc;
  cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 1024);
  for (int wps = 1; wps <= 8; wps *= 2) {
    const int threads = 128 * wps;   // wps warps on each of the 4 schedulers
    if (threads > 1024) break;
    k<KIND><<<1, threads>>>(out, cyc, 12345u);
    k<KIND><<<1, threads>>>(out, cyc, 12345u);
    cudaDeviceSynchronize();
    unsigned long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double inst_per_sched = (double)ITER * CH * wps;
    printf("{\"kind\": \"%s\", \"warps_per_scheduler\": %d, \"cycles\": %llu, \"ipc_per_scheduler\": %.3f}\n", name, wps, c,
           inst_per_sched / (double)c);
  }
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<0>("FMNMX3 (max.f32 3-input)");
  run<3>("FMNMX (max.f32 2-input)");
  run<1>("HMNMX2 (max.f16x2)");
  run<2>("mixed FMNMX3 + HMNMX2");
  run<4>("FFMA");
  run<5>("mixed FMNMX3 + FFMA");
  run<6>("mixed HMNMX2 + FFMA");
  run<7>("IMNMX (max.s32)");
  run<8>("max.s16x2");
  run<9>("mixed FMNMX3 + max.s16x2");
  run<10>("set.ge.f16x2");
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { printf("cuda error %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
