// K7 — on-device negative sampler (the "fast mode" of SURVEY.md §7.2-1).
//
// Replaces, in distribution and in the rejection rules, the host samplers of
// libreco/sampling/negatives.py:17-82 used by BaseCollator.sample_neg_items
// (libreco/batch/collators.py:138-166).  The reference draws from numpy's PCG64 / Python's
// Mersenne Twister — sequential generators; bit-exact reproduction of those streams is the
// PARITY mode and stays on the host (librecommender_b200/sampling.py).  This kernel uses a
// counter-based generator (Philox4x32-10) so that every (sample, attempt) is independent and the
// result is a pure function of (seed, step, index): oracle/sampling.py restates it bit-exactly.
//
//   mode 0 "random"     (negatives.py:17-31): uniform over items; a draw equal to ITS OWN positive is
//                        re-drawn, at most `tolerance` times.
//   mode 1 "unconsumed" (negatives.py:55-82): additionally rejects items the user consumed (binary
//                        search in the per-user SORTED consumed list) and negatives already drawn for
//                        the same positive; after `tolerance` failures the consumed test is dropped
//                        for another `tolerance` tries, then the draw is accepted (same relaxation).
//   mode 2 "popular"    (negatives.py:34-43): inverse-CDF draw from p ~ freq^0.75 (cdf given),
//                        one re-draw when equal to the positive.
// Layout: negatives of positive j are out[j*num_neg : (j+1)*num_neg] (collators.py:231-232).
#include "common.cuh"
#include "../../include/b200reco.h"

namespace b200 {
namespace sampler {

struct U4 { uint32_t x, y, z, w; };

__host__ __device__ __forceinline__ U4 philox4x32_10(U4 ctr, uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)M0 * ctr.x;
    const uint64_t p1 = (uint64_t)M1 * ctr.z;
    U4 n;
    n.x = (uint32_t)(p1 >> 32) ^ ctr.y ^ k0;
    n.y = (uint32_t)p1;
    n.z = (uint32_t)(p0 >> 32) ^ ctr.w ^ k1;
    n.w = (uint32_t)p0;
    ctr = n;
    k0 += W0;
    k1 += W1;
  }
  return ctr;
}

// uniform integer in [0, n): high 64 bits of (64-bit random) * n
__device__ __forceinline__ int64_t bounded(uint32_t hi, uint32_t lo, int64_t n) {
  const uint64_t r = ((uint64_t)hi << 32) | lo;
  return (int64_t)__umul64hi(r, (uint64_t)n);
}

__device__ __forceinline__ bool in_sorted(const int32_t* __restrict__ a, int64_t beg, int64_t end, int32_t v) {
  while (beg < end) {
    const int64_t mid = (beg + end) >> 1;
    const int32_t x = __ldg(a + mid);
    if (x == v) return true;
    if (x < v) beg = mid + 1; else end = mid;
  }
  return false;
}

__device__ __forceinline__ int64_t draw(int mode, const float* __restrict__ cdf, int64_t n_items,
                                        uint64_t seed, uint64_t step, uint64_t index, uint32_t attempt) {
  U4 c;
  c.x = (uint32_t)index; c.y = (uint32_t)(index >> 32); c.z = attempt; c.w = (uint32_t)step;
  const U4 r = philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32) ^ (uint32_t)(step >> 32));
  if (mode != 2) return bounded(r.x, r.y, n_items);
  // inverse CDF: first index whose cdf >= u, u in [0,1) with 24 random bits
  const float u = (float)(r.x >> 8) * (1.0f / 16777216.0f);
  int64_t lo = 0, hi = n_items - 1;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (__ldg(cdf + mid) > u) hi = mid; else lo = mid + 1;
  }
  return lo;
}

__global__ void sample_negatives_kernel(const int64_t* __restrict__ users,
                                        const int64_t* __restrict__ items_pos, int64_t n_pos,
                                        int num_neg, int64_t n_items, int mode, int tolerance,
                                        uint64_t seed, uint64_t step,
                                        const int64_t* __restrict__ indptr,
                                        const int32_t* __restrict__ idx_sorted, int64_t n_users,
                                        const float* __restrict__ cdf, int64_t* __restrict__ out) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_pos) return;
  const int64_t pos = items_pos[j];
  int64_t beg = 0, end = 0;
  if (mode == 1) {
    const int64_t u = users[j];
    if (u >= 0 && u < n_users) { beg = indptr[u]; end = indptr[u + 1]; }
  }
  int64_t* o = out + j * num_neg;
  for (int t = 0; t < num_neg; ++t) {
    const uint64_t index = (uint64_t)(j * num_neg + t);
    uint32_t attempt = 0;
    int64_t n = draw(mode, cdf, n_items, seed, step, index, attempt++);
    if (mode == 0) {
      for (int a = 0; a < tolerance && n == pos; ++a) n = draw(mode, cdf, n_items, seed, step, index, attempt++);
    } else if (mode == 2) {
      if (n == pos) n = draw(mode, cdf, n_items, seed, step, index, attempt++);
    } else {
      bool ok = false;
      for (int a = 0; a < tolerance; ++a) {
        bool dup = false;
        for (int s = 0; s < t; ++s) dup |= (o[s] == n);
        if (n != pos && !dup && !in_sorted(idx_sorted, beg, end, (int32_t)n)) { ok = true; break; }
        n = draw(mode, cdf, n_items, seed, step, index, attempt++);
      }
      if (!ok) {
        for (int a = 0; a < tolerance; ++a) {
          bool dup = false;
          for (int s = 0; s < t; ++s) dup |= (o[s] == n);
          if (n != pos && !dup) break;
          n = draw(mode, cdf, n_items, seed, step, index, attempt++);
        }
      }
    }
    o[t] = n;
  }
}

// ---- per-sample behaviour sequences at collate time (libreco/batch/sequence.py:33-71, mode "recent")
// One warp per sample.  position = FIRST occurrence of the item in the user's time-ordered consumed
// list (list.index); an item the user never consumed (a sampled negative) takes a random position
// (random.randrange(0, len) in the reference: `rand_pos` carries that stream in parity mode, else
// Philox(seed, step, j)).  seq = consumed[max(0, position - L) : position], padded with pad_index;
// len = min(position, L), 1 when position == 0 (reference :56-58).
__global__ void interacted_seqs_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ idx,
                                       int64_t n_users, const int64_t* __restrict__ users,
                                       const int64_t* __restrict__ items, int64_t n, int L, int32_t pad_index,
                                       const int64_t* __restrict__ rand_pos, uint64_t seed, uint64_t step,
                                       int32_t* __restrict__ seqs, int32_t* __restrict__ lens) {
  const int64_t j = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (j >= n) return;
  const int64_t u = users[j];
  int64_t beg = 0, end = 0;
  if (u >= 0 && u < n_users) { beg = indptr[u]; end = indptr[u + 1]; }
  const int64_t clen = end - beg;
  const int64_t item = items[j];
  int64_t pos = -1;
  for (int64_t base = 0; base < clen && pos < 0; base += 32) {
    const bool in = base + lane < clen;
    const bool hit = in && (int64_t)__ldg(idx + beg + base + lane) == item;
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    if (m) pos = base + (__ffs(m) - 1);
  }
  if (pos < 0) {
    if (clen == 0) pos = 0;
    else if (rand_pos) pos = min(max(rand_pos[j], (int64_t)0), clen - 1);
    else {
      U4 c;
      c.x = (uint32_t)j; c.y = (uint32_t)((uint64_t)j >> 32); c.z = 0x5e9u; c.w = (uint32_t)step;
      const U4 r = philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32) ^ (uint32_t)(step >> 32));
      pos = bounded(r.x, r.y, clen);
    }
  }
  const int64_t count = pos < L ? pos : L;
  const int64_t start = pos - count;
  for (int t = lane; t < L; t += 32)
    seqs[j * L + t] = t < count ? __ldg(idx + beg + start + t) : pad_index;
  if (lane == 0) lens[j] = pos == 0 ? 1 : (int32_t)count;
}

}  // namespace sampler
}  // namespace b200

using namespace b200;

extern "C" int b200_interacted_seqs(const int64_t* indptr, const int32_t* idx, int64_t n_users,
                                    const int64_t* users, const int64_t* items, int64_t n,
                                    int32_t max_seq_len, int32_t pad_index, const int64_t* rand_pos,
                                    uint64_t seed, uint64_t step, int32_t* seqs, int32_t* lens, void* stream) {
  B200_REQUIRE(indptr && idx && users && items && seqs && lens, "b200_interacted_seqs: null pointer");
  B200_REQUIRE(max_seq_len >= 1, "b200_interacted_seqs: max_seq_len must be positive");
  if (n == 0) return 0;
  sampler::interacted_seqs_kernel<<<(unsigned)ceil_div64(n * 32, 256), 256, 0, (cudaStream_t)stream>>>(
      indptr, idx, n_users, users, items, n, max_seq_len, pad_index, rand_pos, seed, step, seqs, lens);
  count_launch();
  B200_CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int b200_sample_negatives(const int64_t* users, const int64_t* items_pos, int64_t n_pos,
                                     int32_t num_neg, int64_t n_items, int32_t mode,
                                     int32_t tolerance, uint64_t seed, uint64_t step,
                                     const int64_t* indptr, const int32_t* idx_sorted,
                                     int64_t n_users, const float* cdf, int64_t* out, void* stream) {
  B200_REQUIRE(items_pos && out, "b200_sample_negatives: null pointer");
  B200_REQUIRE(mode >= 0 && mode <= 2, "b200_sample_negatives: unknown mode %d", mode);
  B200_REQUIRE(mode != 1 || (users && indptr && idx_sorted), "unconsumed sampler needs users + sorted consumed CSR");
  B200_REQUIRE(mode != 2 || cdf, "popular sampler needs the cdf");
  B200_REQUIRE(num_neg >= 1 && n_items >= 2, "b200_sample_negatives: bad num_neg / n_items");
  if (n_pos == 0) return 0;
  sampler::sample_negatives_kernel<<<(unsigned)ceil_div64(n_pos, 128), 128, 0, (cudaStream_t)stream>>>(
      users, items_pos, n_pos, num_neg, n_items, mode, tolerance, seed, step, indptr, idx_sorted,
      n_users, cdf, out);
  count_launch();
  B200_CUDA_OK(cudaGetLastError());
  return 0;
}
