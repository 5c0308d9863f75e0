// Hoisted all-items scoring of FM / DeepFM (a3 + a5 of §8).
//
// The reference evaluates the whole graph on B*N rows (libreco/recommendation/recommend.py:81-105).
// Everything that depends on one side only is computed ONCE per user / per item by
// b200_feat_forward (S = sum_f e_f, Q = sum_f e_f^2, the linear partial, and the first MLP layer's
// partial product of that side); here a (user, item) pair only pays for what truly couples the two
// sides: the square of the summed embedding and the small layers of the MLP.
//   sum_f over all fields = (sum over user-side fields) + (sum over item-side fields)   — exact
//   x W1 + b1 = x_u W1[user rows] + x_i W1[item rows] + b1                               — exact
// Re-association only changes fp32 rounding (tests: 1e-5 relative against the row-wise oracle).
#include "common.cuh"
#include "../../include/b200reco.h"

namespace b200 {
namespace pair {

constexpr int MAXK = 64;
constexpr int MAXH2 = 64;
constexpr int MAXH3 = 32;

__global__ void __launch_bounds__(256)
fm_pair_kernel(const float* __restrict__ Su, const float* __restrict__ Qu, const float* __restrict__ lu,
               const float* __restrict__ Si, const float* __restrict__ Qi, const float* __restrict__ li,
               int64_t N, int K, float lin_bias, const float* __restrict__ bn_scale,
               const float* __restrict__ bn_shift, const float* __restrict__ pw_kernel, float pw_bias,
               float* __restrict__ scores, int64_t lds) {
  __shared__ float su[MAXK], qu[MAXK], sc[MAXK], sh[MAXK], wk[MAXK];
  const int64_t b = blockIdx.y;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    su[k] = Su[b * K + k];
    qu[k] = Qu[b * K + k];
    sc[k] = bn_scale ? bn_scale[k] : 1.f;
    sh[k] = bn_shift ? bn_shift[k] : 0.f;
    wk[k] = pw_kernel[k];
  }
  __syncthreads();
  const float lub = lu[b] + lin_bias;
  for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += (int64_t)gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int k = 0; k < K; ++k) {
      const float s = su[k] + __ldg(Si + n * K + k);
      const float q = qu[k] + __ldg(Qi + n * K + k);
      const float pw = 0.5f * (s * s - q);
      acc = fmaf(fmaf(pw, sc[k], sh[k]), wk[k], acc);
    }
    acc += pw_bias;
    scores[b * lds + n] = lub + __ldg(li + n) + (acc > 0.f ? acc : expm1f(acc));
  }
}

struct DeepArgs {
  const float *Su, *Qu, *lu, *Pu, *Si, *Qi, *li, *Pi;
  int64_t N;
  int K, H1, H2, H3;
  float lin_bias;
  const float *W2, *b2, *W3, *b3, *w_out;
  float b_out;
  float* scores;
  int64_t lds;
};

// Two items per thread (the W2 / W3 rows read from shared memory are reused for both), the item-side
// first-layer rows Pi staged through shared memory in 32-column chunks with coalesced loads (a
// thread walking its own 512-byte row straight from global memory was the limiter of the first
// version: 0.8 G pairs/s).  fp32 fma chains in the same k order as before: results unchanged.
constexpr int PAIR_THREADS = 128;
constexpr int PAIR_ITEMS = 2 * PAIR_THREADS;   // items per block iteration
constexpr int KCH = 32;                        // Pi columns per staged chunk

__global__ void __launch_bounds__(PAIR_THREADS)
deepfm_pair_kernel(const DeepArgs a) {
  extern __shared__ float sm[];
  float* w2 = sm;                         // [H1][MAXH2]  (padded with zeros)
  float* w3 = w2 + (size_t)a.H1 * MAXH2;  // [MAXH2][MAXH3]
  float* pu = w3 + MAXH2 * MAXH3;         // [H1]
  float* su = pu + a.H1;                  // [K]
  float* qu = su + a.K;                   // [K]
  float* tile = qu + a.K;                 // 2 x [PAIR_ITEMS][KCH + 1]  (double buffer)
  const int64_t b = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  for (int i = tid; i < a.H1 * MAXH2; i += PAIR_THREADS) {
    const int k = i / MAXH2, j = i % MAXH2;
    w2[i] = j < a.H2 ? a.W2[(size_t)k * a.H2 + j] : 0.f;
  }
  for (int i = tid; i < MAXH2 * MAXH3; i += PAIR_THREADS) {
    const int k = i / MAXH3, j = i % MAXH3;
    w3[i] = (a.H3 > 0 && k < a.H2 && j < a.H3) ? a.W3[(size_t)k * a.H3 + j] : 0.f;
  }
  for (int i = tid; i < a.H1; i += PAIR_THREADS) pu[i] = a.Pu[b * a.H1 + i];
  for (int i = tid; i < a.K; i += PAIR_THREADS) { su[i] = a.Su[b * a.K + i]; qu[i] = a.Qu[b * a.K + i]; }
  __syncthreads();
  const float lub = a.lu[b] + a.lin_bias;
  const int n_deep = a.H3 > 0 ? a.H3 : a.H2;
  // The (item tile, k-chunk) sequence of this block is one stream of staged chunks: chunk c+1 is
  // fetched with cp.async into the other half of `tile` while chunk c is consumed.
  const int nch = (a.H1 + KCH - 1) / KCH;
  const int64_t n_iters = a.N > (int64_t)blockIdx.x * PAIR_ITEMS
                              ? (a.N - (int64_t)blockIdx.x * PAIR_ITEMS + (int64_t)gridDim.x * PAIR_ITEMS - 1) /
                                    ((int64_t)gridDim.x * PAIR_ITEMS)
                              : 0;
  const int64_t n_chunks = n_iters * nch;
  auto fetch = [&](int64_t c) {
    const int64_t base = ((int64_t)blockIdx.x + (c / nch) * gridDim.x) * PAIR_ITEMS;
    const int kc = (int)(c % nch) * KCH;
    const int kw = min(KCH, a.H1 - kc);
    float* dst = tile + (c & 1) * (PAIR_ITEMS * (KCH + 1));
    for (int r = wid; r < PAIR_ITEMS; r += PAIR_THREADS / 32) {      // one 128-byte row segment per warp step
      const int64_t n = base + r;
      float* d = dst + r * (KCH + 1) + lane;
      if (n < a.N && lane < kw) {
        const uint32_t sa = (uint32_t)__cvta_generic_to_shared(d);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(a.Pi + n * a.H1 + kc + lane) : "memory");
      } else {
        *d = 0.f;
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  if (n_chunks > 0) fetch(0);
  for (int64_t it = 0; it < n_iters; ++it) {
    const int64_t base = ((int64_t)blockIdx.x + it * gridDim.x) * PAIR_ITEMS;
    float h2[2][MAXH2];
#pragma unroll
    for (int j = 0; j < MAXH2; ++j) { h2[0][j] = 0.f; h2[1][j] = 0.f; }
    for (int ci = 0; ci < nch; ++ci) {
      const int64_t c = it * nch + ci;
      const int kc = ci * KCH;
      const int kw = min(KCH, a.H1 - kc);
      if (c + 1 < n_chunks) {
        fetch(c + 1);                                    // the other buffer: freed by the barrier below (previous round)
        asm volatile("cp.async.wait_group 1;" ::: "memory");
      } else {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
      }
      __syncthreads();                                   // chunk c landed for every thread
      const float* tb = tile + (c & 1) * (PAIR_ITEMS * (KCH + 1));
      const float* t0 = tb + tid * (KCH + 1);
      const float* t1 = tb + (tid + PAIR_THREADS) * (KCH + 1);
#pragma unroll 2      // keep the loop body inside the instruction cache
      for (int kk = 0; kk < kw; ++kk) {
        const float p = pu[kc + kk];
        const float ha = fmaxf(p + t0[kk], 0.f);
        const float hb = fmaxf(p + t1[kk], 0.f);
        const float4* wrow = reinterpret_cast<const float4*>(w2 + (size_t)(kc + kk) * MAXH2);
#pragma unroll
        for (int j4 = 0; j4 < MAXH2 / 4; ++j4) {
          const float4 w = wrow[j4];
          h2[0][4 * j4 + 0] = fmaf(ha, w.x, h2[0][4 * j4 + 0]);
          h2[0][4 * j4 + 1] = fmaf(ha, w.y, h2[0][4 * j4 + 1]);
          h2[0][4 * j4 + 2] = fmaf(ha, w.z, h2[0][4 * j4 + 2]);
          h2[0][4 * j4 + 3] = fmaf(ha, w.w, h2[0][4 * j4 + 3]);
          h2[1][4 * j4 + 0] = fmaf(hb, w.x, h2[1][4 * j4 + 0]);
          h2[1][4 * j4 + 1] = fmaf(hb, w.y, h2[1][4 * j4 + 1]);
          h2[1][4 * j4 + 2] = fmaf(hb, w.z, h2[1][4 * j4 + 2]);
          h2[1][4 * j4 + 3] = fmaf(hb, w.w, h2[1][4 * j4 + 3]);
        }
      }
      __syncthreads();                                   // chunk c consumed: its buffer may be refilled
    }
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int64_t n = base + tid + e * PAIR_THREADS;
      if (n >= a.N) continue;
      // output head starts with the linear and pairwise blocks of w_out
      float out = a.b_out + (lub + __ldg(a.li + n)) * __ldg(a.w_out);
      for (int k = 0; k < a.K; ++k) {
        const float s = su[k] + __ldg(a.Si + n * a.K + k);
        const float q = qu[k] + __ldg(a.Qi + n * a.K + k);
        out = fmaf(0.5f * (s * s - q), __ldg(a.w_out + 1 + k), out);
      }
      if (a.H3 > 0) {
        // third layer with the OUTPUT index as the (rolled) outer loop: 64 fma per iteration on
        // register-resident activations, a few hundred bytes of code instead of 2 x 2048 unrolled fma
        float v[MAXH2];
#pragma unroll
        for (int k = 0; k < MAXH2; ++k) v[k] = k < a.H2 ? fmaxf(h2[e][k] + __ldg(a.b2 + k), 0.f) : 0.f;
#pragma unroll 1
        for (int j = 0; j < a.H3; ++j) {
          float h3 = __ldg(a.b3 + j);
#pragma unroll
          for (int k = 0; k < MAXH2; ++k) h3 = fmaf(v[k], w3[k * MAXH3 + j], h3);
          out = fmaf(h3, __ldg(a.w_out + 1 + a.K + j), out);
        }
      } else {
#pragma unroll
        for (int j = 0; j < MAXH2; ++j)
          if (j < n_deep) out = fmaf(h2[e][j] + __ldg(a.b2 + j), __ldg(a.w_out + 1 + a.K + j), out);
      }
      a.scores[b * a.lds + n] = out;
    }
  }
}

}  // namespace pair
}  // namespace b200

using namespace b200;
using namespace b200::pair;

extern "C" int b200_fm_pair_scores(const float* Su, const float* Qu, const float* lu, int64_t B,
                                   const float* Si, const float* Qi, const float* li, int64_t N,
                                   int32_t K, float lin_bias, const float* bn_scale,
                                   const float* bn_shift, const float* pw_kernel, float pw_bias,
                                   float* scores, int64_t lds, void* stream) {
  B200_REQUIRE(Su && Qu && lu && Si && Qi && li && pw_kernel && scores, "b200_fm_pair_scores: null pointer");
  B200_REQUIRE(K >= 1 && K <= MAXK, "b200_fm_pair_scores: embed size %d outside [1, %d]", K, MAXK);
  B200_REQUIRE(B <= 65535, "b200_fm_pair_scores: at most 65535 users per call");
  if (B == 0 || N == 0) return 0;
  const unsigned gx = (unsigned)min((int64_t)1024, ceil_div64(N, 256));
  fm_pair_kernel<<<dim3(gx, (unsigned)B), 256, 0, (cudaStream_t)stream>>>(
      Su, Qu, lu, Si, Qi, li, N, K, lin_bias, bn_scale, bn_shift, pw_kernel, pw_bias, scores, lds);
  count_launch();
  B200_CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int b200_deepfm_pair_scores(const float* Su, const float* Qu, const float* lu, const float* Pu,
                                       int64_t B, const float* Si, const float* Qi, const float* li,
                                       const float* Pi, int64_t N, int32_t K, int32_t H1, int32_t H2,
                                       int32_t H3, float lin_bias, const float* W2, const float* b2,
                                       const float* W3, const float* b3, const float* w_out,
                                       float b_out, float* scores, int64_t lds, void* stream) {
  B200_REQUIRE(Su && Qu && lu && Pu && Si && Qi && li && Pi && W2 && b2 && w_out && scores,
               "b200_deepfm_pair_scores: null pointer");
  B200_REQUIRE(K >= 1 && K <= MAXK && H1 >= 1 && H1 <= 256 && H2 >= 1 && H2 <= MAXH2 && H3 >= 0 && H3 <= MAXH3,
               "b200_deepfm_pair_scores: unsupported layer sizes K=%d H=(%d,%d,%d)", K, H1, H2, H3);
  B200_REQUIRE(H3 == 0 || (W3 && b3), "b200_deepfm_pair_scores: third layer weights missing");
  B200_REQUIRE(B <= 65535, "b200_deepfm_pair_scores: at most 65535 users per call");
  if (B == 0 || N == 0) return 0;
  DeepArgs a;
  a.Su = Su; a.Qu = Qu; a.lu = lu; a.Pu = Pu; a.Si = Si; a.Qi = Qi; a.li = li; a.Pi = Pi; a.N = N;
  a.K = K; a.H1 = H1; a.H2 = H2; a.H3 = H3; a.lin_bias = lin_bias; a.W2 = W2; a.b2 = b2; a.W3 = W3;
  a.b3 = b3; a.w_out = w_out; a.b_out = b_out; a.scores = scores; a.lds = lds;
  const size_t smem = ((size_t)H1 * MAXH2 + MAXH2 * MAXH3 + H1 + 2 * K + 2 * PAIR_ITEMS * (KCH + 1)) * sizeof(float);
  static bool attr = false;
  if (!attr) {
    B200_CUDA_OK(cudaFuncSetAttribute(deepfm_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 140 * 1024));
    attr = true;
  }
  const unsigned gx = (unsigned)min((int64_t)512, ceil_div64(N, PAIR_ITEMS));
  deepfm_pair_kernel<<<dim3(gx, (unsigned)B), PAIR_THREADS, smem, (cudaStream_t)stream>>>(a);
  count_launch();
  B200_CUDA_OK(cudaGetLastError());
  return 0;
}
