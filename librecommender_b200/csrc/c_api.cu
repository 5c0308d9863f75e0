// C-ABI plumbing: error strings, launch counter, host-side consumed -> CSR.
#include "common.cuh"
#include "../../include/b200reco.h"
#include <stdarg.h>
#include <vector>

namespace b200 {

static thread_local char g_err[512] = "";
unsigned long long g_launch_count = 0;

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return 0;
  set_last_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return -1;
}

}  // namespace b200

extern "C" int b200_version(void) { return B200RECO_VERSION; }
extern "C" const char* b200_last_error(void) { return b200::g_err; }
extern "C" unsigned long long b200_launch_count(void) { return b200::g_launch_count; }

// rust/src/utils.rs:8-35 (build_consumed_unique): group items per user in arrival order, then
// Vec::dedup() == drop CONSECUTIVE repeats only.  Two counting passes, no hashing.
extern "C" int b200_build_consumed_csr_host(const int64_t* user_indices, const int64_t* item_indices,
                                            int64_t n, int64_t n_users, int64_t* indptr,
                                            int32_t* idx, int64_t* nnz_out) {
  B200_REQUIRE(indptr && nnz_out && (n == 0 || (user_indices && item_indices && idx)),
               "b200_build_consumed_csr_host: null pointer");
  B200_REQUIRE(n_users >= 0 && n >= 0, "b200_build_consumed_csr_host: negative size");
  std::vector<int64_t> last((size_t)n_users, -1);  // last item appended per user (-1 = none)
  std::vector<int64_t> cnt((size_t)n_users + 1, 0);
  for (int64_t j = 0; j < n; ++j) {
    const int64_t u = user_indices[j];
    B200_REQUIRE(u >= 0 && u < n_users, "user index %lld out of range at %lld", (long long)u, (long long)j);
    B200_REQUIRE(item_indices[j] >= 0 && item_indices[j] < (1ll << 31), "item index out of range");
    if (last[u] != item_indices[j]) { cnt[u]++; last[u] = item_indices[j]; }
  }
  indptr[0] = 0;
  for (int64_t u = 0; u < n_users; ++u) indptr[u + 1] = indptr[u] + cnt[u];
  std::vector<int64_t> pos(indptr, indptr + n_users);
  std::fill(last.begin(), last.end(), -1);
  for (int64_t j = 0; j < n; ++j) {
    const int64_t u = user_indices[j];
    if (last[u] != item_indices[j]) {
      idx[pos[u]++] = (int32_t)item_indices[j];
      last[u] = item_indices[j];
    }
  }
  *nnz_out = indptr[n_users];
  return 0;
}
