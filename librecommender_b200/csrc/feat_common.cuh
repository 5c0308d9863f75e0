// Helpers shared by the forward (feat.cu) and backward (train.cu) feature kernels.
#pragma once
#include "common.cuh"
#include "../../include/b200reco.h"

namespace b200 {
namespace feat {

__device__ __forceinline__ float subwarp_sum(float v, int lpr, uint32_t gmask) {
  for (int o = lpr >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(gmask, v, o);
  return v;
}

// field f of row r -> (table row index) for sparse fields, value for dense fields
__device__ __forceinline__ int32_t sparse_index(const b200_feat_layout& L, int64_t r, int64_t u, int64_t it, int f) {
  if (L.sparse_rows) return L.sparse_rows[r * L.ld_sparse_rows + f];
  return L.sparse_side[f] == 0 ? __ldg(L.user_sparse_unique + u * L.ld_us + L.sparse_col[f])
                               : __ldg(L.item_sparse_unique + it * L.ld_is + L.sparse_col[f]);
}
__device__ __forceinline__ float dense_value(const b200_feat_layout& L, int64_t r, int64_t u, int64_t it, int f) {
  if (L.dense_rows) return L.dense_rows[r * L.ld_dense_rows + f];
  return L.dense_side[f] == 0 ? __ldg(L.user_dense_unique + u * L.ld_ud + L.dense_col[f])
                              : __ldg(L.item_dense_unique + it * L.ld_id + L.dense_col[f]);
}

}  // namespace feat
}  // namespace b200
