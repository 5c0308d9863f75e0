// Row-sharded embedding tables over NVLink peer memory (reference: one table,
// libreco/layers/embedding.py:16-23).
//
// Row r of a table lives on GPU r % G at slot r / G.  Every GPU's shard is allocated in symmetric
// (peer-mapped) memory, so a kernel on ANY GPU can address every shard: `shards[g]` is the device
// pointer of GPU g's shard as mapped into this process.  NVSwitch gives every peer the same
// bandwidth, so the lookup needs no request exchange, no bucketing by owner and no second
// all-to-all for the rows:
//
//   b200_peer_gather_rows      out[i, :] = shards[id % G][(id / G) * ld + :]   (P2P loads, the
//                              requester pulls; local rows take the same code path through HBM)
//   b200_peer_scatter_add_rows shards[id % G][(id / G) * ld + :] += rows[i, :] (P2P float atomics:
//                              the gradient of the lookup, pushed to the owners)
//
// One kernel per direction does the gather AND the exchange.  A sub-warp of LPR = d/4 lanes (16-byte
// accesses) serves one row; every sub-warp keeps UNROLL rows in flight so that the NVLink round trip
// (~2 us) is covered by independent requests.  With G == 1 the same kernels are the plain local
// gather / scatter-add.
#include "common.cuh"
#include "../../include/b200reco.h"

namespace b200 {

constexpr int kPeerMaxRanks = 16;
struct PeerShards {
  const float* p[kPeerMaxRanks];
};
struct PeerShardsMut {
  float* p[kPeerMaxRanks];
};

template <int LPR, int UNROLL>   // LPR lanes x float4 per row (d == 4 * LPR)
__global__ void __launch_bounds__(256)
peer_gather_vec4_kernel(PeerShards sh, int G, int64_t ld, const int64_t* __restrict__ ids, int64_t n,
                        float* __restrict__ out, int64_t ld_out) {
  const int sub = threadIdx.x / LPR, l = threadIdx.x % LPR;
  constexpr int SUBS = 256 / LPR;
  const int64_t base = ((int64_t)blockIdx.x * SUBS + sub) * UNROLL;
  float4 v[UNROLL];
  bool ok[UNROLL];
#pragma unroll
  for (int u = 0; u < UNROLL; ++u) {
    const int64_t i = base + u;
    ok[u] = i < n;
    v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ok[u]) {
      const int64_t id = __ldg(ids + i);
      const int g = (int)(id % G);
      const float* src = sh.p[g] + (id / G) * ld;
      v[u] = *reinterpret_cast<const float4*>(src + 4 * l);     // peer (or local) 16-byte load
    }
  }
#pragma unroll
  for (int u = 0; u < UNROLL; ++u)
    if (ok[u]) *reinterpret_cast<float4*>(out + (base + u) * ld_out + 4 * l) = v[u];
}

__global__ void peer_gather_generic_kernel(PeerShards sh, int G, int64_t ld, int d,
                                           const int64_t* __restrict__ ids, int64_t n,
                                           float* __restrict__ out, int64_t ld_out) {
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= n) return;
  const int64_t id = ids[r];
  const float* src = sh.p[(int)(id % G)] + (id / G) * ld;
  for (int k = lane; k < d; k += 32) out[r * ld_out + k] = src[k];
}

__global__ void peer_scatter_add_kernel(PeerShardsMut sh, int G, int64_t ld, int d,
                                        const int64_t* __restrict__ ids, int64_t n,
                                        const float* __restrict__ rows, int64_t ld_rows) {
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= n) return;
  const int64_t id = ids[r];
  float* dst = sh.p[(int)(id % G)] + (id / G) * ld;
  for (int k = lane; k < d; k += 32) atomicAdd(dst + k, rows[r * ld_rows + k]);   // RED over NVLink for peers
}

template <int LPR>
static void launch_gather_vec4(const PeerShards& sh, int G, int64_t ld, const int64_t* ids, int64_t n,
                               float* out, int64_t ld_out, cudaStream_t stream) {
  constexpr int UNROLL = 8;
  constexpr int SUBS = 256 / LPR;
  const int64_t rows_per_block = (int64_t)SUBS * UNROLL;
  peer_gather_vec4_kernel<LPR, UNROLL><<<(unsigned)ceil_div64(n, rows_per_block), 256, 0, stream>>>(
      sh, G, ld, ids, n, out, ld_out);
}

}  // namespace b200

using namespace b200;

extern "C" int b200_peer_gather_rows(const void* const* shards, int32_t n_ranks, int64_t ld, int32_t d,
                                     const int64_t* ids, int64_t n, float* out, int64_t ld_out,
                                     void* stream_) {
  if (n == 0) return 0;
  B200_REQUIRE(shards && ids && out && d > 0, "b200_peer_gather_rows: bad arguments");
  B200_REQUIRE(n_ranks >= 1 && n_ranks <= kPeerMaxRanks, "b200_peer_gather_rows: 1..%d ranks", kPeerMaxRanks);
  cudaStream_t stream = (cudaStream_t)stream_;
  PeerShards sh;
  bool aligned = (ld % 4 == 0) && (ld_out % 4 == 0) && (((uintptr_t)out & 15) == 0);
  for (int g = 0; g < n_ranks; ++g) {
    B200_REQUIRE(shards[g], "b200_peer_gather_rows: null shard pointer for rank %d", g);
    sh.p[g] = (const float*)shards[g];
    aligned = aligned && (((uintptr_t)shards[g] & 15) == 0);
  }
  if (aligned && d == 16) launch_gather_vec4<4>(sh, n_ranks, ld, ids, n, out, ld_out, stream);
  else if (aligned && d == 32) launch_gather_vec4<8>(sh, n_ranks, ld, ids, n, out, ld_out, stream);
  else if (aligned && d == 64) launch_gather_vec4<16>(sh, n_ranks, ld, ids, n, out, ld_out, stream);
  else if (aligned && d == 128) launch_gather_vec4<32>(sh, n_ranks, ld, ids, n, out, ld_out, stream);
  else
    peer_gather_generic_kernel<<<(unsigned)ceil_div64(n * 32, 256), 256, 0, stream>>>(sh, n_ranks, ld, d, ids, n,
                                                                                   out, ld_out);
  count_launch();
  B200_CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int b200_peer_scatter_add_rows(void* const* shards, int32_t n_ranks, int64_t ld, int32_t d,
                                          const int64_t* ids, int64_t n, const float* rows,
                                          int64_t ld_rows, void* stream_) {
  if (n == 0) return 0;
  B200_REQUIRE(shards && ids && rows && d > 0, "b200_peer_scatter_add_rows: bad arguments");
  B200_REQUIRE(n_ranks >= 1 && n_ranks <= kPeerMaxRanks, "b200_peer_scatter_add_rows: 1..%d ranks", kPeerMaxRanks);
  PeerShardsMut sh;
  for (int g = 0; g < n_ranks; ++g) {
    B200_REQUIRE(shards[g], "b200_peer_scatter_add_rows: null shard pointer for rank %d", g);
    sh.p[g] = (float*)shards[g];
  }
  peer_scatter_add_kernel<<<(unsigned)ceil_div64(n * 32, 256), 256, 0, (cudaStream_t)stream_>>>(
      sh, n_ranks, ld, d, ids, n, rows, ld_rows);
  count_launch();
  B200_CUDA_OK(cudaGetLastError());
  return 0;
}
