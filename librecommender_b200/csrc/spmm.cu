// K6 — CSR SpMM for LightGCN propagation, with the layer-mean accumulation fused.
//
// Replaces torch.sparse.mm(L, E) + stack/mean of
// libreco/algorithms/torch_modules/lightgcn_module.py:74-88 (embedding_propagation).
//
//   out[r, :] = sum_j val[j] * E[col[j], :]      j in [indptr[r], indptr[r+1])   (fma, CSR order)
//   acc[r, :] (+)= out[r, :]                      running sum of the layers (optional)
//   acc[r, :] /= final_div                        on the last layer (optional)
//
// HBM-bound gather: per nnz 4 B (col) + 4 B (val) + 4*d B (the gathered row).  One sub-warp of
// `lpr` lanes per row (lpr = min(32, pow2 >= d)), lanes stride the embedding width so that every
// gathered row is read with full 128-B requests; indices/values are loaded cooperatively and
// broadcast with shuffles; 4 gathers in flight per sub-warp.  Rows longer than LONG_ROW nnz are
// skipped by the main kernel and processed as fixed-size chunks (one warp per chunk -> partial
// sums -> ordered reduction), so that a Zipf-popular item with millions of edges cannot serialise
// a warp.  Everything is deterministic (no float atomics).
#include "common.cuh"
#include "../../include/b200reco.h"

namespace b200 {
namespace spmm {

constexpr int MAX_T = 8;          // d <= 256
constexpr int LONG_ROW = 1024;    // rows above this go to the chunked path
constexpr int CHUNK = 1024;       // nnz per chunk of a long row

struct Args {
  const int64_t* indptr;
  const int32_t* col;
  const float* val;
  int64_t n_rows;
  const float* E; int64_t ld_e;
  float* out; int64_t ld_out;
  float* acc; int64_t ld_acc;
  int acc_init;       // 1: acc = E_in_row + out (first layer: E^0 enters the sum), 0: acc += out
  float final_div;    // > 0: acc /= final_div after the add
  int d, lpr, T;
};

__device__ __forceinline__ void row_epilogue(const Args& a, int64_t r, int li, const float (&sum)[MAX_T]) {
#pragma unroll
  for (int t = 0; t < MAX_T; ++t) {
    const int c = li + t * a.lpr;
    if (t < a.T && c < a.d) {
      if (a.out) a.out[r * a.ld_out + c] = sum[t];
      if (a.acc) {
        float base = a.acc_init ? __ldg(a.E + r * a.ld_e + c) : a.acc[r * a.ld_acc + c];
        float v = base + sum[t];
        if (a.final_div > 0.f) v = v / a.final_div;
        a.acc[r * a.ld_acc + c] = v;
      }
    }
  }
}

// accumulate nnz [beg, end) of one row into sum[] (sub-warp of lpr lanes, lane index li)
__device__ __forceinline__ void accumulate_range(const Args& a, int64_t beg, int64_t end, int li,
                                                 uint32_t gmask, int gbase, float (&sum)[MAX_T]) {
  for (int64_t j0 = beg; j0 < end; j0 += a.lpr) {
    const int64_t j = j0 + li;
    int32_t c = 0;
    float v = 0.f;
    if (j < end) { c = __ldg(a.col + j); v = __ldg(a.val + j); }
    const int cnt = (int)min((int64_t)a.lpr, end - j0);
    for (int q0 = 0; q0 < cnt; q0 += 4) {
      int32_t cc[4];
      float vv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int src = gbase + min(q0 + u, cnt - 1);
        cc[u] = __shfl_sync(gmask, c, src);
        vv[u] = __shfl_sync(gmask, v, src);
        if (q0 + u >= cnt) vv[u] = 0.f;
      }
      float e[4][MAX_T];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int t = 0; t < MAX_T; ++t) {
          const int cidx = li + t * a.lpr;
          e[u][t] = (t < a.T && cidx < a.d) ? __ldg(a.E + (int64_t)cc[u] * a.ld_e + cidx) : 0.f;
        }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int t = 0; t < MAX_T; ++t) sum[t] = fmaf(vv[u], e[u][t], sum[t]);
    }
  }
}

__global__ void __launch_bounds__(256)
spmm_rows_kernel(const Args a) {
  const int lane = threadIdx.x & 31;
  const int rows_per_warp = 32 / a.lpr;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int g = lane / a.lpr, li = lane % a.lpr;
  const int64_t r = warp * rows_per_warp + g;
  const int gbase = g * a.lpr;
  const uint32_t gmask = (a.lpr == 32) ? 0xffffffffu : (((1u << a.lpr) - 1u) << gbase);
  if (r >= a.n_rows) return;
  const int64_t beg = a.indptr[r], end = a.indptr[r + 1];
  if (end - beg > LONG_ROW) return;   // chunked path
  float sum[MAX_T];
#pragma unroll
  for (int t = 0; t < MAX_T; ++t) sum[t] = 0.f;
  accumulate_range(a, beg, end, li, gmask, gbase, sum);
  row_epilogue(a, r, li, sum);
}

// ---- fast path: d % 4 == 0, 16-byte aligned rows: LPR lanes x float4 per row, 8 gathers in flight
template <int LPR>
__device__ __forceinline__ float4 accumulate_vec4(const Args& a, int64_t beg, int64_t end, int li,
                                                  uint32_t gmask, int gbase) {
  constexpr int U = 8;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const bool col_ok = li * 4 < a.d;
  for (int64_t j0 = beg; j0 < end; j0 += LPR) {
    const int64_t j = j0 + li;
    int32_t c = 0;
    float v = 0.f;
    if (j < end) { c = __ldg(a.col + j); v = __ldg(a.val + j); }
    const int cnt = (int)min((int64_t)LPR, end - j0);
    for (int q0 = 0; q0 < cnt; q0 += U) {
      int32_t cc[U];
      float vv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int src = gbase + min(q0 + u, cnt - 1);
        cc[u] = __shfl_sync(gmask, c, src);
        vv[u] = __shfl_sync(gmask, v, src);
        if (q0 + u >= cnt) vv[u] = 0.f;
      }
      float4 e[U];
#pragma unroll
      for (int u = 0; u < U; ++u)
        e[u] = col_ok ? __ldg(reinterpret_cast<const float4*>(a.E + (int64_t)cc[u] * a.ld_e) + li)
                      : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        acc.x = fmaf(vv[u], e[u].x, acc.x);
        acc.y = fmaf(vv[u], e[u].y, acc.y);
        acc.z = fmaf(vv[u], e[u].z, acc.z);
        acc.w = fmaf(vv[u], e[u].w, acc.w);
      }
    }
  }
  return acc;
}

__device__ __forceinline__ void row_epilogue_vec4(const Args& a, int64_t r, int li, float4 sum) {
  if (li * 4 >= a.d) return;
  if (a.out) *(reinterpret_cast<float4*>(a.out + r * a.ld_out) + li) = sum;
  if (a.acc) {
    float4 base = a.acc_init ? __ldg(reinterpret_cast<const float4*>(a.E + r * a.ld_e) + li)
                             : *(reinterpret_cast<const float4*>(a.acc + r * a.ld_acc) + li);
    float4 v = make_float4(base.x + sum.x, base.y + sum.y, base.z + sum.z, base.w + sum.w);
    if (a.final_div > 0.f) { v.x /= a.final_div; v.y /= a.final_div; v.z /= a.final_div; v.w /= a.final_div; }
    *(reinterpret_cast<float4*>(a.acc + r * a.ld_acc) + li) = v;
  }
}

template <int LPR>
__global__ void __launch_bounds__(256)
spmm_rows_vec4_kernel(const Args a) {
  const int lane = threadIdx.x & 31;
  constexpr int RPW = 32 / LPR;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int g = lane / LPR, li = lane % LPR;
  const int64_t r = warp * RPW + g;
  const int gbase = g * LPR;
  const uint32_t gmask = (LPR == 32) ? 0xffffffffu : (((1u << LPR) - 1u) << gbase);
  if (r >= a.n_rows) return;
  const int64_t beg = a.indptr[r], end = a.indptr[r + 1];
  if (end - beg > LONG_ROW) return;   // chunked path
  row_epilogue_vec4(a, r, li, accumulate_vec4<LPR>(a, beg, end, li, gmask, gbase));
}

template <int LPR>
__global__ void __launch_bounds__(256)
spmm_long_chunks_vec4_kernel(const Args a, const int32_t* __restrict__ chunk_row,
                             const int32_t* __restrict__ chunk_k, int64_t n_chunks,
                             float* __restrict__ partials) {
  const int lane = threadIdx.x & 31;
  constexpr int RPW = 32 / LPR;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int g = lane / LPR, li = lane % LPR;
  const int64_t chunk = warp * RPW + g;
  const int gbase = g * LPR;
  const uint32_t gmask = (LPR == 32) ? 0xffffffffu : (((1u << LPR) - 1u) << gbase);
  if (chunk >= n_chunks) return;
  const int64_t r = chunk_row[chunk];
  const int64_t beg = a.indptr[r] + (int64_t)chunk_k[chunk] * CHUNK;
  const int64_t end = min(beg + (int64_t)CHUNK, a.indptr[r + 1]);
  const float4 s = accumulate_vec4<LPR>(a, beg, end, li, gmask, gbase);
  if (li * 4 < a.d) *(reinterpret_cast<float4*>(partials + chunk * a.d) + li) = s;
}

template <int LPR>
static void launch_vec4(const Args& a, const int32_t* chunk_row, const int32_t* chunk_k, int64_t n_chunks,
                        float* partials, int64_t n_long, cudaStream_t stream) {
  constexpr int RPW = 32 / LPR;
  spmm_rows_vec4_kernel<LPR><<<(unsigned)ceil_div64(ceil_div64(a.n_rows, RPW) * 32, 256), 256, 0, stream>>>(a);
  if (n_long > 0)
    spmm_long_chunks_vec4_kernel<LPR><<<(unsigned)ceil_div64(ceil_div64(n_chunks, RPW) * 32, 256), 256, 0, stream>>>(
        a, chunk_row, chunk_k, n_chunks, partials);
}

// one warp (lpr forced to its row width) per CHUNK nnz of a long row -> partials[chunk, d]
__global__ void __launch_bounds__(256)
spmm_long_chunks_kernel(const Args a, const int32_t* __restrict__ chunk_row,
                        const int32_t* __restrict__ chunk_k, int64_t n_chunks,
                        float* __restrict__ partials) {
  const int lane = threadIdx.x & 31;
  const int64_t chunk = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (chunk >= n_chunks) return;
  // long rows always use the whole warp: lanes beyond lpr replicate work on other columns
  Args b = a;
  b.lpr = 32;
  b.T = (a.d + 31) / 32;
  const int64_t r = chunk_row[chunk];
  const int64_t beg = a.indptr[r] + (int64_t)chunk_k[chunk] * CHUNK;
  const int64_t end = min(beg + (int64_t)CHUNK, a.indptr[r + 1]);
  float sum[MAX_T];
#pragma unroll
  for (int t = 0; t < MAX_T; ++t) sum[t] = 0.f;
  accumulate_range(b, beg, end, lane, 0xffffffffu, 0, sum);
#pragma unroll
  for (int t = 0; t < MAX_T; ++t) {
    const int c = lane + t * 32;
    if (t < b.T && c < a.d) partials[chunk * a.d + c] = sum[t];
  }
}

// reduction of the partials of each long row in a FIXED order (deterministic): one CTA per row, warp w adds the
// chunks c = w, w + 8, ... into two alternating accumulators (two loads in flight per lane), warp 0 then adds the
// 8 x 2 partial sums in order.  (A single warp per row made the most popular item — ~2 M nnz = ~2000 chunks on a
// Zipf graph — a serial chain of ~2000 dependent loads that did not shrink when the rows were sharded: the
// sharded propagation's strong-scaling efficiency was 0.44 at 4 GPUs because of it.)
constexpr int RED_WARPS = 8;
__global__ void __launch_bounds__(RED_WARPS * 32)
spmm_long_reduce_kernel(const Args a, const int32_t* __restrict__ long_rows,
                        const int64_t* __restrict__ long_chunk_ptr, int64_t n_long,
                        const float* __restrict__ partials) {
  __shared__ float sh[RED_WARPS][2][32 * MAX_T];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t w = blockIdx.x;
  if (w >= n_long) return;
  Args b = a;
  b.lpr = 32;
  b.T = (a.d + 31) / 32;
  const int64_t r = long_rows[w];
  float s0[MAX_T], s1[MAX_T];
#pragma unroll
  for (int t = 0; t < MAX_T; ++t) { s0[t] = 0.f; s1[t] = 0.f; }
  const int64_t c_beg = long_chunk_ptr[w], c_end = long_chunk_ptr[w + 1];
  for (int64_t c = c_beg + wid; c < c_end; c += 2 * RED_WARPS) {
    const int64_t c2 = c + RED_WARPS;
#pragma unroll
    for (int t = 0; t < MAX_T; ++t) {
      const int cc = lane + t * 32;
      if (t < b.T && cc < a.d) {
        s0[t] += partials[c * a.d + cc];
        if (c2 < c_end) s1[t] += partials[c2 * a.d + cc];
      }
    }
  }
#pragma unroll
  for (int t = 0; t < MAX_T; ++t) {
    sh[wid][0][lane + t * 32] = s0[t];
    sh[wid][1][lane + t * 32] = s1[t];
  }
  __syncthreads();
  if (wid != 0) return;
  float sum[MAX_T];
#pragma unroll
  for (int t = 0; t < MAX_T; ++t) {
    float v = 0.f;
#pragma unroll
    for (int q = 0; q < RED_WARPS; ++q) { v += sh[q][0][lane + t * 32]; v += sh[q][1][lane + t * 32]; }
    sum[t] = v;
  }
  row_epilogue(b, r, lane, sum);
}

}  // namespace spmm
}  // namespace b200

using namespace b200;
using namespace b200::spmm;

extern "C" int b200_spmm_long_row_threshold(void) { return LONG_ROW; }
extern "C" int b200_spmm_chunk(void) { return CHUNK; }

extern "C" int b200_spmm_csr(const int64_t* indptr, const int32_t* col, const float* val,
                             int64_t n_rows, const float* E, int64_t ld_e, int32_t d, float* out,
                             int64_t ld_out, float* acc, int64_t ld_acc, int32_t acc_init,
                             float final_div, const int32_t* long_rows,
                             const int64_t* long_chunk_ptr, int64_t n_long,
                             const int32_t* chunk_row, const int32_t* chunk_k, int64_t n_chunks,
                             float* partials, void* stream_) {
  B200_REQUIRE(indptr && E && (out || acc), "b200_spmm_csr: null pointer");
  B200_REQUIRE(d >= 1 && d <= 32 * MAX_T, "b200_spmm_csr: embed width %d outside [1, %d]", d, 32 * MAX_T);
  B200_REQUIRE(n_long == 0 || (long_rows && long_chunk_ptr && chunk_row && chunk_k && partials),
               "b200_spmm_csr: long-row plan missing");
  if (n_rows == 0) return 0;
  cudaStream_t stream = (cudaStream_t)stream_;
  Args a;
  a.indptr = indptr; a.col = col; a.val = val; a.n_rows = n_rows; a.E = E; a.ld_e = ld_e;
  a.out = out; a.ld_out = ld_out; a.acc = acc; a.ld_acc = ld_acc; a.acc_init = acc_init;
  a.final_div = final_div; a.d = d;
  int lpr = 1;
  while (lpr < d && lpr < 32) lpr <<= 1;
  a.lpr = lpr;
  a.T = (d + lpr - 1) / lpr;
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  const bool vec4 = d % 4 == 0 && d <= 128 && ld_e % 4 == 0 && al16(E) &&
                    (!out || (ld_out % 4 == 0 && al16(out))) && (!acc || (ld_acc % 4 == 0 && al16(acc))) &&
                    (n_long == 0 || al16(partials));
  if (vec4) {
    const int q = d / 4;
    if (q <= 4) launch_vec4<4>(a, chunk_row, chunk_k, n_chunks, partials, n_long, stream);
    else if (q <= 8) launch_vec4<8>(a, chunk_row, chunk_k, n_chunks, partials, n_long, stream);
    else if (q <= 16) launch_vec4<16>(a, chunk_row, chunk_k, n_chunks, partials, n_long, stream);
    else launch_vec4<32>(a, chunk_row, chunk_k, n_chunks, partials, n_long, stream);
    count_launch(n_long > 0 ? 2 : 1);
  } else {
    const int rows_per_warp = 32 / lpr;
    const int64_t warps = ceil_div64(n_rows, rows_per_warp);
    spmm_rows_kernel<<<(unsigned)ceil_div64(warps * 32, 256), 256, 0, stream>>>(a);
    count_launch();
    if (n_long > 0) {
      spmm_long_chunks_kernel<<<(unsigned)ceil_div64(n_chunks * 32, 256), 256, 0, stream>>>(
          a, chunk_row, chunk_k, n_chunks, partials);
      count_launch();
    }
  }
  if (n_long > 0) {
    spmm_long_reduce_kernel<<<(unsigned)n_long, RED_WARPS * 32, 0, stream>>>(
        a, long_rows, long_chunk_ptr, n_long, partials);
    count_launch();
  }
  B200_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---- NGCF layer epilogue (reference libreco/algorithms/torch_modules/ngcf_module.py:104-121) ---------
// mul:      out[r, :] = a[r, :] * b[r, :]                      (side ⊙ E, the input of the pair GEMM)
// combine:  m = leaky_relu(a + b, 0.2); out[r, :] = m / max(||m||_2, 1e-12)   (F.normalize, eps 1e-12)
namespace b200 {
namespace ngcf {

__global__ void mul_rows_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n,
                                float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i] * b[i];
}

__global__ void combine_kernel(const float* __restrict__ a, int64_t lda, const float* __restrict__ b, int64_t ldb,
                               int64_t R, int d, float slope, float* __restrict__ out, int64_t ldo) {
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= R) return;
  float ss = 0.f;
  for (int k = lane; k < d; k += 32) {
    float m = a[r * lda + k] + b[r * ldb + k];
    m = m > 0.f ? m : slope * m;
    ss = fmaf(m, m, ss);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
  for (int k = lane; k < d; k += 32) {
    float m = a[r * lda + k] + b[r * ldb + k];
    m = m > 0.f ? m : slope * m;
    out[r * ldo + k] = m * inv;
  }
}

}  // namespace ngcf
}  // namespace b200

extern "C" int b200_mul_elementwise(const float* a, const float* b, int64_t n, float* out, void* stream) {
  B200_REQUIRE(a && b && out, "b200_mul_elementwise: null pointer");
  if (n == 0) return 0;
  b200::ngcf::mul_rows_kernel<<<(unsigned)b200::ceil_div64(n, 256), 256, 0, (cudaStream_t)stream>>>(a, b, n, out);
  B200_CUDA_OK(cudaGetLastError());
  b200::count_launch();
  return 0;
}

extern "C" int b200_ngcf_combine(const float* self_part, int64_t lda, const float* pair_part, int64_t ldb,
                                 int64_t R, int32_t d, float negative_slope, float* out, int64_t ld_out,
                                 void* stream) {
  B200_REQUIRE(self_part && pair_part && out && d > 0, "b200_ngcf_combine: bad arguments");
  if (R == 0) return 0;
  b200::ngcf::combine_kernel<<<(unsigned)b200::ceil_div64(R * 32, 256), 256, 0, (cudaStream_t)stream>>>(
      self_part, lda, pair_part, ldb, R, d, negative_slope, out, ld_out);
  B200_CUDA_OK(cudaGetLastError());
  b200::count_launch();
  return 0;
}
