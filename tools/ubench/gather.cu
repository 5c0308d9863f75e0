// What does HBM3e give for RANDOM row gathers?  (sm_100a micro-benchmark: the ceiling K1 is judged against)
// Bare gather: `lanes` consecutive lanes read one ROW_B-byte row (16 B each) picked by a random id, UNROLL
// independent rows in flight per lane group, result folded into one register and written once per thread
// (so the only HBM traffic is the gathered rows + 8 B/id).  Prints GB/s of row bytes for several row
// sizes and table sizes (inside / outside the 126 MB L2), plus a streaming read of the same volume.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/gather tools/ubench/gather.cu
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

template <int LPR, int UNROLL>   // LPR lanes per row (row bytes = 16 * LPR)
__global__ void gather_k(const float4* __restrict__ table, const int64_t* __restrict__ ids, int64_t n_ids,
                         float* __restrict__ sink) {
  const int64_t groups = (int64_t)gridDim.x * blockDim.x / LPR;
  const int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LPR;
  const int l = threadIdx.x % LPR;
  float acc = 0.f;
  for (int64_t base = g * UNROLL; base < n_ids; base += groups * UNROLL) {
    float4 v[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const int64_t i = base + u;
      const int64_t id = i < n_ids ? __ldg(ids + i) : 0;
      v[u] = __ldg(table + id * LPR + l);
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
  }
  if (acc == 12345.678f) sink[0] = acc;
}

__global__ void stream_k(const float4* __restrict__ p, int64_t n, float* sink) {
  float acc = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float4 v = __ldg(p + i);
    acc += v.x + v.y + v.z + v.w;
  }
  if (acc == 12345.678f) sink[0] = acc;
}

template <int LPR, int UNROLL>
static float run(const float4* table, const int64_t* ids, int64_t n_ids, float* sink, int ctas_per_sm) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int grid = 148 * ctas_per_sm;
  for (int w = 0; w < 2; ++w) gather_k<LPR, UNROLL><<<grid, 256>>>(table, ids, n_ids, sink);
  cudaEventRecord(e0);
  const int it = 5;
  for (int w = 0; w < it; ++w) gather_k<LPR, UNROLL><<<grid, 256>>>(table, ids, n_ids, sink);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms / it;
}

int main() {
  float* sink;
  cudaMalloc(&sink, 4);
  const int64_t n_ids = 64ll << 20;   // 64 Mi row reads per launch
  int64_t* ids;
  cudaMalloc(&ids, n_ids * 8);
  std::vector<int64_t> h(n_ids);
  for (int row_b : {64, 128, 256}) {
    for (double table_mb : {32.0, 152.0, 1300.0, 8000.0}) {
      const int64_t rows = (int64_t)(table_mb * 1e6 / row_b);
      uint64_t s = 88172645463325252ull;
      for (int64_t i = 0; i < n_ids; ++i) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        h[i] = (int64_t)(s % (uint64_t)rows);
      }
      cudaMemcpy(ids, h.data(), n_ids * 8, cudaMemcpyHostToDevice);
      float4* table;
      if (cudaMalloc(&table, rows * row_b) != cudaSuccess) { printf("{\"skip\": %f}\n", table_mb); continue; }
      cudaMemset(table, 0, rows * row_b);
      float best = 1e30f;
      int best_cfg = 0;
      for (int cps : {4, 8}) {
        float ms[3];
        if (row_b == 64) { ms[0] = run<4, 4>(table, ids, n_ids, sink, cps); ms[1] = run<4, 8>(table, ids, n_ids, sink, cps); ms[2] = run<4, 16>(table, ids, n_ids, sink, cps); }
        else if (row_b == 128) { ms[0] = run<8, 4>(table, ids, n_ids, sink, cps); ms[1] = run<8, 8>(table, ids, n_ids, sink, cps); ms[2] = run<8, 16>(table, ids, n_ids, sink, cps); }
        else { ms[0] = run<16, 4>(table, ids, n_ids, sink, cps); ms[1] = run<16, 8>(table, ids, n_ids, sink, cps); ms[2] = run<16, 16>(table, ids, n_ids, sink, cps); }
        for (int k = 0; k < 3; ++k) if (ms[k] < best) { best = ms[k]; best_cfg = cps * 100 + (4 << k); }
      }
      printf("{\"ubench\": \"random row gather\", \"row_bytes\": %d, \"table_mb\": %.0f, \"rows_read\": %lld, \"ms\": %.4f, "
             "\"row_gbs\": %.1f, \"row_plus_id_gbs\": %.1f, \"best_ctas_per_sm_x100_plus_unroll\": %d}\n",
             row_b, table_mb, (long long)n_ids, best, n_ids * (double)row_b / best / 1e6,
             n_ids * (double)(row_b + 8) / best / 1e6, best_cfg);
      fflush(stdout);
      cudaFree(table);
    }
  }
  {  // streaming read reference, 4 GB
    const int64_t n = (4ll << 30) / 16;
    float4* p;
    cudaMalloc(&p, n * 16);
    cudaMemset(p, 0, n * 16);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    stream_k<<<148 * 8, 256>>>(p, n, sink);
    cudaEventRecord(e0);
    for (int w = 0; w < 5; ++w) stream_k<<<148 * 8, 256>>>(p, n, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("{\"ubench\": \"streaming read\", \"gb\": 4.29, \"ms\": %.4f, \"gbs\": %.1f}\n", ms / 5, n * 16.0 / (ms / 5) / 1e6);
  }
  return 0;
}
