// B1 — NN-descent kNN graph on device (placeholder until the tiled local-join kernel lands).
#include "internal.h"
namespace eps {
int nn_descent(Index* ix, int64_t n, int K, const eps_build_params& bp, unsigned long long* d_knn, eps_stats* st) {
  (void)ix; (void)n; (void)K; (void)bp; (void)d_knn; (void)st;
  return fail(EPS_ERR_UNSUPPORTED, "NN-descent build path not available; raise exact_knn_below");
}
}  // namespace eps
