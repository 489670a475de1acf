// scorer_mlp_tc.cuh -- NANN_SCORER_TENSOR: tcgen05 path of the mlp scorer (placeholder until the
// kernel lands; the EXACT path is the product path meanwhile).
#pragma once
namespace nann {
static nann_status mlp_tc_prepare(nann_scorer* s) {
  (void)s;
  return fail(NANN_UNIMPLEMENTED, "tensor-core scorer not built in this revision");
}
static nann_status mlp_tc_score(nann_scorer* s, const ScoreCall& c, cudaStream_t st) {
  (void)s; (void)c; (void)st;
  return fail(NANN_UNIMPLEMENTED, "tensor-core scorer not built in this revision");
}
static void mlp_tc_release(nann_scorer* s) { (void)s; }
}  // namespace nann
