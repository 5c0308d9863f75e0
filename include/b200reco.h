/*
 * b200reco.h — C-ABI of librecommender_b200 (sm_100a only).
 *
 * The reference (massquantity/LibRecommender @ 7463d9d) has no FFI on this path:
 * its seams are Python callables.  Each entry point below names
 * the reference function whose arithmetic it replaces; the Python shims in
 * librecommender_b200/ keep the reference signatures and call these through
 * ctypes.  Conventions:
 *   - every function returns 0 on success, <0 on error; b200_last_error() gives
 *     the message for the calling thread;
 *   - pointers are DEVICE pointers unless the name ends in _host / says host;
 *   - no ownership transfer: all buffers (incl. workspaces sized by the
 *     *_workspace_bytes queries) are allocated by the caller;
 *   - stream-ordered on `stream` (a cudaStream_t passed as void*), re-entrant,
 *     no global state except the launch counter.
 */
#ifndef B200RECO_H_
#define B200RECO_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200RECO_VERSION 100

int b200_version(void);
const char* b200_last_error(void);
/* number of kernels this library has launched in this process (bench.py: gpu_launches) */
unsigned long long b200_launch_count(void);

/* ---- a0: user_consumed -> CSR ---------------------------------------------------------
 * Replaces recfarm.build_consumed_unique (rust/src/utils.rs:8-35) as used by
 * libreco/data/consumed.py:7-19: interactions are grouped per user in arrival order and
 * CONSECUTIVE repeats are dropped.  HOST function (all pointers host).
 * indptr_host[n_users+1], idx_host capacity n; *nnz_host receives the kept count. */
int b200_build_consumed_csr_host(const int64_t* user_indices_host, const int64_t* item_indices_host,
                                 int64_t n, int64_t n_users, int64_t* indptr_host,
                                 int32_t* idx_host, int64_t* nnz_host);

/* ---- a2: rank_recommendations (libreco/recommendation/ranking.py:10-78) ----------------
 * b200_mask_consumed: filter_items (:59-61) under the rule of :38 — row r (user user_ids[r])
 * is masked iff c_u > 0 and K + c_u <= N, c_u = indptr[u+1]-indptr[u] (duplicates counted);
 * users >= n_users (OOV) are never masked.  Masked scores are overwritten with -inf. */
int b200_mask_consumed(float* scores, int64_t ld, const int64_t* user_ids, int64_t B, int64_t N,
                       int32_t K, const int64_t* indptr, const int32_t* idx, int64_t n_users,
                       void* stream);

/* b200_topk_rows: partition_select + argsort (:48-49,:76-78): per row the K largest scores,
 * sorted by (score desc, item id asc).  out_ids int64 [B,K]; out_scores float [B,K] or NULL.
 * K <= 4096 and K <= N. */
int b200_topk_rows_workspace_bytes(int64_t B, int64_t N, int32_t K, size_t* bytes);
int b200_topk_rows(const float* scores, int64_t ld, int64_t B, int64_t N, int32_t K,
                   int64_t* out_ids, float* out_scores, void* workspace, size_t workspace_bytes,
                   void* stream);

/* ---- a1: recommend_from_embedding (libreco/recommendation/recommend.py:57-78) ----------
 * scores[r, n] = sum_k U[user_ids[r], k] * I[n, k], fp32, one accumulator per output,
 * fused-multiply-add in increasing k (the library's exact-score definition). */
int b200_score_rows_f32(const float* U, int64_t ldu, const int64_t* user_ids, int64_t B,
                        const float* I, int64_t ldi, int64_t N, int32_t d, float* scores,
                        int64_t lds, void* stream);

/* ---- a1+a2 fused: recommend_from_embedding + rank_recommendations on the tensor cores -----
 * (recommend.py:57-78 + ranking.py:10-78, never materialising [B,N]).
 * catalog: device buffer prepared ONCE per item table (bf16 K-major copy + max row norm).
 * Result per row: the K best non-consumed items by EXACT fp32 score (same definition as
 * b200_score_rows_f32), sorted (score desc, id asc).  row_status[r] (device int32[B]) = 1 marks a
 * row the fused path could not bound (K + consumed > 288, or too many near-ties): its out_ids are
 * -1 and the caller re-runs it through b200_score_rows_f32 + b200_mask_consumed + b200_topk_rows.
 * Limits: d <= 256, K <= 288. */
int b200_embed_catalog_bytes(int64_t N, int32_t d, size_t* bytes);
int b200_embed_catalog_prepare(const float* I, int64_t ldi, int64_t N, int32_t d, void* catalog,
                               size_t bytes, void* stream);
int b200_recommend_embed_workspace_bytes(int64_t B, int64_t N, int32_t d, int32_t K, size_t* bytes);
int b200_recommend_embed(const float* U, int64_t ldu, const int64_t* user_ids, int64_t B,
                         const float* I, int64_t ldi, int64_t N, int32_t d, const void* catalog,
                         const int64_t* indptr, const int32_t* idx, int64_t n_users, int32_t filter,
                         int32_t K, int64_t* out_ids, float* out_scores, int32_t* row_status,
                         void* workspace, size_t workspace_bytes, void* stream,
                         void* ev_sweep_start /* cudaEvent_t or NULL: recorded on `stream` */,
                         void* ev_sweep_stop  /* just before / after the tcgen05 sweep kernel */);

/* ---- a10: LightGCN propagation (libreco/algorithms/torch_modules/lightgcn_module.py:66-88) -
 * out[r,:] = sum_j val[j] * E[col[j],:] over the CSR row r (fma in CSR order), optionally fused with
 * the layer-mean: acc = (acc_init ? E[r,:] : acc[r,:]) + out[r,:], then acc /= final_div if > 0.
 * Rows longer than b200_spmm_long_row_threshold() nnz must be listed in long_rows and split into
 * chunks of b200_spmm_chunk() nnz: chunk c covers nnz [indptr[row] + chunk_k[c]*chunk, ...) of
 * row chunk_row[c]; long_chunk_ptr[n_long+1] delimits each long row's chunks; partials is a
 * caller-provided float [n_chunks, d] scratch.  Deterministic (no float atomics).  d <= 256. */
int b200_spmm_long_row_threshold(void);
int b200_spmm_chunk(void);
int b200_spmm_csr(const int64_t* indptr, const int32_t* col, const float* val, int64_t n_rows,
                  const float* E, int64_t ld_e, int32_t d, float* out, int64_t ld_out, float* acc,
                  int64_t ld_acc, int32_t acc_init, float final_div, const int32_t* long_rows,
                  const int64_t* long_chunk_ptr, int64_t n_long, const int32_t* chunk_row,
                  const int32_t* chunk_k, int64_t n_chunks, float* partials, void* stream);

/* ---- a12: negative sampling (libreco/sampling/negatives.py:17-82; collators.py:138-166) ---
 * Counter-based (Philox4x32-10) device sampler; result = f(seed, step, index) only.
 * mode 0 random, 1 unconsumed (needs users + per-user SORTED consumed CSR), 2 popular (needs the
 * cdf of p ~ freq^0.75).  out[j*num_neg + t] = t-th negative of positive j.  The bit-exact
 * reproduction of the reference's numpy / Mersenne streams is the host parity mode
 * (librecommender_b200/sampling.py). */
int b200_sample_negatives(const int64_t* users, const int64_t* items_pos, int64_t n_pos,
                          int32_t num_neg, int64_t n_items, int32_t mode, int32_t tolerance,
                          uint64_t seed, uint64_t step, const int64_t* indptr,
                          const int32_t* idx_sorted, int64_t n_users, const float* cdf, int64_t* out,
                          void* stream);

/* ---- a14: predict_from_embedding (libreco/prediction/predict.py:36-40) -----------------
 * out[r] = sum_k U[users[r],k] * I[items[r],k]; mode 0: raw, 1: expit (ranking),
 * 2: clip to [lo, hi] (rating) — normalize_prediction (:18-23). */
int b200_gather_dot(const float* U, int64_t ldu, const int64_t* users, const float* I,
                    int64_t ldi, const int64_t* items, int64_t n, int32_t d, int32_t mode,
                    float lo, float hi, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200RECO_H_ */
