/*
 * nann_oracle.h -- CPU restatement of alibaba/nann's model-scored HNSW retrieval hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (nann_b200/, include/nann_b200.h) never links, imports or calls anything in oracle/.
 *
 * Parity pinning (see DESIGN.md "Oracle"):
 *   - integer ops + TopKV2: pinned against the reference's own test vectors
 *     (tests/golden/kat_*.json, extracted from the reference test scripts cited there).
 *   - whole traversal (exec.pb dataflow): the reference holds no golden vector and cannot be
 *     built here (needs patched TF 1.15 + bazel) -> restated from build_opt_graph.py; it is
 *     pinned only through the op-level vectors.  "traversal: parity pinned at op level only".
 *   - scorer arithmetic: the reference's XLA/cuBLAS summation order is unspecified ->
 *     "scorer: parity unpinned"; this file FIXES one fp32 definition (sequential-k fmaf).
 *
 * Paths below are relative to /root/reference; UO = tensorflow/tensorflow/core/user_ops.
 */
#ifndef NANN_ORACLE_H_
#define NANN_ORACLE_H_

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Status codes: numeric values of tensorflow::error::Code
 * (tensorflow/tensorflow/core/lib/core/error_codes.proto). */
enum {
  ORC_OK = 0,
  ORC_INVALID_ARGUMENT = 3,
  ORC_NOT_FOUND = 5,
  ORC_UNIMPLEMENTED = 12,
  ORC_INTERNAL = 13
};

/* ---- ragged validation: UO/beam_search_op/GroupGather_kernel.cc:9-16 --------------------
 * returns 0 ok, 1 row_splits empty, 2 row_splits[0]!=0, 3 row_splits[-1]!=n_values */
int orc_validate_ragged(int64_t n_values, const int64_t* row_splits, int64_t n_row_splits);

/* ---- GroupGather: UO/beam_search_op/GroupGather_kernel.cc:55-170 -------------------------
 * Two-call protocol: call with ret_values==NULL to get *n_ret (and ret_row_splits filled),
 * then again with a buffer.  ret_row_splits must hold n_irs entries (or 1 for void inputs).
 * *n_ret_rs receives the number of row_splits written.  unique!=0 -> per-group first-occurrence
 * order (the reference's order is std::unordered_set iteration order, i.e. unspecified). */
int orc_group_gather_i32(const int32_t* pv, int64_t n_pv, const int64_t* prs, int64_t n_prs,
                         const int64_t* iv, int64_t n_iv, const int64_t* irs, int64_t n_irs,
                         int unique, int32_t* ret_values, int64_t* n_ret,
                         int64_t* ret_row_splits, int64_t* n_ret_rs, int* code);
int orc_group_gather_i64(const int64_t* pv, int64_t n_pv, const int64_t* prs, int64_t n_prs,
                         const int64_t* iv, int64_t n_iv, const int64_t* irs, int64_t n_irs,
                         int unique, int64_t* ret_values, int64_t* n_ret,
                         int64_t* ret_row_splits, int64_t* n_ret_rs, int* code);

/* ---- BitmapRefDifference: UO/bitmap_op/bitmap_ops.cc:175-257 ----------------------------
 * flags mutated in place.  c_values needs capacity n_v.  No range check on ids, as the
 * reference (bitmap_ops.cc:225-231); n_flags is only used by the optional debug bound check
 * (returns ORC_INVALID_ARGUMENT when check_bounds!=0 and an id is out of range). */
int orc_bitmap_ref_difference_i32(const int32_t* v, int64_t n_v, const int64_t* rs, int64_t n_rs,
                                  int32_t* flags, int64_t n_flags, int check_bounds,
                                  int32_t* c_values, int64_t* n_c,
                                  int64_t* c_row_splits, int64_t* n_c_rs, int* code);
int orc_bitmap_ref_difference_i64(const int64_t* v, int64_t n_v, const int64_t* rs, int64_t n_rs,
                                  int32_t* flags, int64_t n_flags, int check_bounds,
                                  int64_t* c_values, int64_t* n_c,
                                  int64_t* c_row_splits, int64_t* n_c_rs, int* code);

/* ---- TopKV2 (CPU functor): tensorflow/tensorflow/core/kernels/topk_op.cc:51-93,102-207 ---
 * input [rows, cols] row-major; outputs [rows, k]; sorted=1 as build_opt_graph.py uses it.
 * Order: value descending, ties -> smaller index first (topk_op.cc:142-150).
 * Errors: k<0 -> "Need k >= 0"; cols<k -> "input must have at least k columns" (:66-69). */
int orc_topk_v2_f32(const float* input, int64_t rows, int64_t cols, int k,
                    float* values, int32_t* indices);

/* ---- BatchTopKOnRT: UO/topk_op/BatchTopKOnRT_kernel.cc:62-156 ------------------------------
 * k: n_k==1 scalar or one per group; per group min(len,k) results, group-local indices (:146).
 * The reference's partial_sort_copy leaves tie order unspecified; this restatement breaks ties
 * by smaller position (a valid instance).  values_out/idx_out need capacity n_values. */
int orc_batch_topk_on_rt_f32(const float* values, int64_t n_values, const int64_t* rs, int64_t n_rs,
                             const int64_t* k, int64_t n_k, int ascending, float* values_out,
                             int64_t* idx_out, int64_t* rs_out, int64_t* n_out, int* code);

/* ---- row gather (stock GatherV2, build_opt_graph.py:92,144) ----------------------------- */
void orc_gather_rows_f32(const float* table, int64_t dim, const int32_t* ids, int64_t n, float* out);
void orc_gather_i64(const int64_t* table, const int32_t* ids, int64_t n, int64_t* out);

/* ---- synthetic scorer "mlp2x512" (BASELINE configs 2-5) ------------------
 * s(u,x) = w3 . relu(W2 . relu(W1 [u;x] + b1) + b2), u,x in R^d, hidden H.
 * fp32 DEFINITION fixed here: every dot product is a sequential chain of fmaf in increasing
 * k, starting from the bias (0 for the last layer); relu(a) = a>0?a:0.
 * W1 is [H][2d] row-major (cols 0..d-1 multiply u), W2 [H][H], w3 [H]. */
typedef struct orc_mlp orc_mlp_t;
orc_mlp_t* orc_mlp_create(int d, int H, const float* W1, const float* b1,
                          const float* W2, const float* b2, const float* w3);
void orc_mlp_destroy(orc_mlp_t* m);
/* scalar, un-hoisted, literal definition (slow; validates the blocked version) */
void orc_mlp_score_def(const orc_mlp_t* m, const float* u, const float* x, int64_t n, float* out);
/* blocked AVX2 version of the same definition: bit-identical results.  rows = table[ids[i]]
 * when ids!=NULL else table + i*d. */
void orc_mlp_score(const orc_mlp_t* m, const float* u, const float* table, const int32_t* ids,
                   int64_t n, float* out);

/* ---- reference scorer "nann_attention" (config 1): NANN_impls/nann/model/model.py:189-233,
 *      model_util.py:9-11,32-67,70-97; fp32, BN in inference form with eps=1e-3.
 * Weight blob layout documented in oracle/README.md; built by tests from a seeded generator. */
typedef struct orc_attn orc_attn_t;
orc_attn_t* orc_attn_create(const float* blob, int64_t n_floats);
void orc_attn_destroy(orc_attn_t* a);
int64_t orc_attn_blob_size(void);
/* user: [50][64]; rows: table[ids[i]] (64 floats each) */
void orc_attn_score(const orc_attn_t* a, const float* user, const float* table,
                    const int32_t* ids, int64_t n, float* out);

/* ---- index in Appendix-C layout (build_hnsw_index.py:33-67; widths as build_opt_graph.py:87) */
typedef struct {
  int64_t n_items;
  int dim;
  const float* emb;          /* [n_items][dim] */
  const int64_t* item_ids;   /* [n_items] */
  const int32_t* ep;         /* enter points, ascending node ids */
  int64_t n_ep;
  const int32_t* nbr_values[2];      /* level 0, level 1 */
  const int64_t* nbr_row_splits[2];  /* [n_items+1] */
} orc_index_t;

/* Scorer callback: fills out[0..n) for node ids[0..n).  round = 0..4 (level2, level1, L0x3). */
typedef void (*orc_score_fn)(void* ctx, int round, const int32_t* ids, int64_t n, float* out);

typedef struct {
  int64_t n_scored[5];   /* rows scored per round */
  int64_t n_expanded[5]; /* ids produced by GroupGather per round (before the bitmap) */
} orc_search_stats_t;

/* exec.pb dataflow, one query: NANN_impls/nann/delivery/build_opt_graph.py:109-149.
 * level_topn[6].  out_ids (item ids, int64) / out_scores / out_nodes need level_topn[5] slots.
 * trace (optional): per round r, trace_ids[r]/trace_scores[r] receive the scored node ids and
 * scores (caller provides capacity trace_cap each), trace_n[r] the counts.
 * Returns ORC_OK or ORC_INVALID_ARGUMENT (TopKV2 n<k, topk_op.cc:66-69; or the squeeze-to-
 * scalar rank error when exactly one candidate is scored, build_opt_graph.py:107). */
int orc_search(const orc_index_t* ix, orc_score_fn score, void* score_ctx,
               const int32_t* level_topn, int64_t* out_ids, float* out_scores, int32_t* out_nodes,
               orc_search_stats_t* stats,
               int32_t** trace_ids, float** trace_scores, int64_t* trace_n, int64_t trace_cap);

/* `main.py --job-type test` traversal, one query: NANN_impls/nann/model/model.py:299-362.
 * num_scoring_per_level[3] / top_k_per_level[3] are indexed by LEVEL (0..2) like the reference's flags
 * (config.py:52-55: defaults [3,1,1] / [400,200,100]); start level 2 scores every enter point once.
 * round numbers passed to `score` count scoring calls from 0.  out_* need topk_eval slots; *out_n = number
 * of valid results (clamped top-k, model.py:268).  ORC_INVALID_ARGUMENT when a round scores exactly one
 * candidate (tf.squeeze -> scalar) or num_scoring_per_level[2] != 1 (assert, :347). */
int orc_search_eval(const orc_index_t* ix, orc_score_fn score, void* score_ctx,
                    const int32_t* num_scoring_per_level, const int32_t* top_k_per_level, int topk_eval,
                    int64_t* out_ids, float* out_scores, int32_t* out_nodes, int32_t* out_n,
                    int64_t* n_scored_total);

/* convenience: mlp2x512 scorer, batch of queries, nthreads request-parallel workers
 * (one request per core, each request single-threaded, as blaze-benchmark's consumers:
 * blaze-benchmark/benchmark/core/benchmark.cc:126-132).  users [B][d].
 * status[B]; out_* [B][level_topn[5]].  Returns wall seconds spent. */
double orc_search_batch_mlp(const orc_index_t* ix, const orc_mlp_t* m, const float* users,
                            int64_t B, const int32_t* level_topn, int nthreads,
                            int64_t* out_ids, float* out_scores, int32_t* status,
                            int64_t* n_scored_total);

/* ---- HugeConst npy header check: UO/huge_const_op/huge_const_op.cc:85-147 ---------------
 * dtype codes: 0 f16, 1 f32, 2 f64, 3 i32, 4 i64.  Reads the payload into dst (capacity
 * dst_bytes) when dst!=NULL.  Returns ORC_NOT_FOUND / ORC_UNIMPLEMENTED (fortran) /
 * ORC_INTERNAL (shape or dtype mismatch) like the reference. */
int orc_huge_const_load(const char* path, int dtype, const int64_t* shape, int rank,
                        void* dst, int64_t dst_bytes);

/* ---- BloomFilterDifference: UO/bitmap_op/bitmap_ops.cc:264-432.  flags (int32[n_flags], n_flags >= bucket_size) is the
 * Ref input, mutated in place; c_values needs n_v slots, c_row_splits n_rs.  `code` = ValidateRaggedTensor result. */
uint64_t orc_fingerprint64(const char* s, int64_t len);       /* farmhash::Fingerprint64 for len <= 32 */
void orc_bloom_primes(int64_t bucket_size, int64_t primes[4]);
int orc_bloom_filter_difference_i32(const int32_t* v, int64_t n_v, const int64_t* rs, int64_t n_rs, int32_t* flags,
                                    int64_t n_flags, int64_t bucket, int64_t bucket_size, int32_t* c_values,
                                    int64_t* c_row_splits, int64_t* n_c, int* code);
int orc_bloom_filter_difference_i64(const int64_t* v, int64_t n_v, const int64_t* rs, int64_t n_rs, int32_t* flags,
                                    int64_t n_flags, int64_t bucket, int64_t bucket_size, int64_t* c_values,
                                    int64_t* c_row_splits, int64_t* n_c, int* code);

/* ---- index construction: CPU statement of the CUDA builder's batch construction (NOT faiss; see the .c file).
 * X [s][d] = the members of one level; out_links [s][cap] member-local ids closest-first, out_cnt [s]. */
int orc_build_level(const float* X, int64_t s, int d, int n_cand, int M, int cap, int nthreads,
                    int32_t* out_links, int32_t* out_cnt);

const char* orc_version(void);

#ifdef __cplusplus
}
#endif
#endif
