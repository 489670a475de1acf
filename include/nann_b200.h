/*
 * nann_b200.h -- C ABI of libnann_b200.so: a B200 (sm_100a) implementation of alibaba/nann's
 * model-scored HNSW retrieval hot path, shaped so the reference's TensorFlow custom-op
 * boundary can bind it (see INTEGRATION.md for the shim).
 *
 * Reference interfaces replaced (paths relative to the reference checkout,
 * UO = tensorflow/tensorflow/core/user_ops):
 *   nann_huge_const_*             HugeConstantOp          UO/huge_const_op/huge_const_op.cc:58-252
 *   nann_group_gather_*           GroupGather<T>::Compute UO/beam_search_op/GroupGather_kernel.cc:18-182
 *   nann_bitmap_ref_difference_*  BitmapRefDifference<T>  UO/bitmap_op/bitmap_ops.cc:150-257
 *   nann_topk_v2_f32              TopK<CPU,float> (TopKV2) tensorflow/tensorflow/core/kernels/topk_op.cc:40-230
 *   nann_gather_rows              GatherV2 as used at NANN_impls/nann/delivery/build_opt_graph.py:92,144
 *   nann_scorer_* / nann_blaze_xla_run   BlazeXlaOp + BlazeXlaPredictor::Compute
 *                                 UO/blaze_op/blaze_xla_kernel.cc:24-33,194-258, blaze_xla_predictor.cc:360-459
 *   nann_index_*, nann_search_*   the whole exec.pb dataflow, build_opt_graph.py:69-160, for a
 *                                 BATCH of queries in one call (the reference runs batch=1)
 *   nann_merge_topk               new: per-shard top-k merge after an all-gather (torch.distributed transport)
 *   nann_shard_group_*, nann_search_sharded   new: sharded HNSW, exchange + merge inside the library over NVLink peer windows
 *   nann_dist_group_*, nann_search_distributed, nann_index_create_sharded
 *                                 new: one graph, embedding table row-sharded, distributed scoring (bit-identical results)
 *   nann_bloom_filter_difference_*  BloomFilterDifference<T>  UO/bitmap_op/bitmap_ops.cc:264-432
 *   nann_scorer_set_admission     BlazeXlaOp::Schedule    UO/blaze_op/blaze_xla_kernel.cc:87-101,221-258
 *   nann_hnsw_build               faiss IndexHNSWFlat + CSR dump  NANN_impls/nann/delivery/build_hnsw_index.py:33-67
 *   nann_executor_*               blaze-benchmark's session pool + consumers
 *                                 blaze-benchmark/benchmark/core/model.cc:128-237, predict_request_consumer.cc:17-54
 *
 * Conventions
 *   - Status: every call returns a tensorflow::error::Code value (0 = OK).  The message a TF
 *     kernel would put in the Status is available from nann_last_error() (thread local).
 *   - Memory space: any data pointer may be HOST or DEVICE memory; the library inspects it
 *     (cudaPointerGetAttributes) and stages host buffers over the device itself.  There is NO CPU
 *     implementation behind this ABI: without a CUDA device every compute entry point fails with
 *     NANN_FAILED_PRECONDITION.
 *   - Streams: `stream` is a cudaStream_t passed as void* (NULL = default stream).  Calls that
 *     hand results back to host memory synchronise that stream before returning; nann_search_batch and
 *     nann_search_sharded with all-device outputs (and no host-side stats) only enqueue work.
 *     A small-batch nann_search_batch call with stream == NULL and only HOST pointers runs on a stream of the searcher's
 *     own (so that its launch sequence can replay as a CUDA graph; the NULL stream cannot be captured); it synchronises
 *     before returning, like every call with host outputs.
 *   - Data-dependent output sizes (GroupGather, BitmapRefDifference) use an allocator callback,
 *     the C equivalent of OpKernelContext::allocate_output: the library calls
 *     alloc(ctx, output_index, n_elems) once per output and writes n_elems elements there.
 *   - Ownership: the library never frees caller memory; handles are opaque and freed by their
 *     *_destroy.  All entry points are re-entrant; a handle may be used from several threads as
 *     long as each call uses its own stream (nann_search_* additionally needs its own searcher).
 */
#ifndef NANN_B200_H_
#define NANN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NANN_B200_ABI_VERSION 2

typedef int nann_status;
enum {
  NANN_OK = 0,
  NANN_INVALID_ARGUMENT = 3,
  NANN_DEADLINE_EXCEEDED = 4,
  NANN_NOT_FOUND = 5,
  NANN_RESOURCE_EXHAUSTED = 8,
  NANN_FAILED_PRECONDITION = 9,
  NANN_UNIMPLEMENTED = 12,
  NANN_INTERNAL = 13
};

/* dtype codes (the five HugeConst is registered for, huge_const_op.cc:230-252) */
enum { NANN_F16 = 0, NANN_F32 = 1, NANN_F64 = 2, NANN_I32 = 3, NANN_I64 = 4 };

typedef void* (*nann_alloc_fn)(void* ctx, int output_index, int64_t n_elems);

int nann_abi_version(void);
const char* nann_last_error(void);
/* number of CUDA kernels this library has launched since load (process wide) */
uint64_t nann_kernel_launch_count(void);
/* device_count, SM count and HBM bytes of `device`; NANN_FAILED_PRECONDITION without a GPU */
nann_status nann_device_info(int device, int* device_count, int* sm_count, int64_t* hbm_bytes,
                             int* cc_major, int* cc_minor);

/* ------------------------------------------------------------------------------------------
 * HugeConst: npy file -> tensor (host copy + one cached device copy)
 * checks and error codes as huge_const_op.cc:85-147: NotFound (open), Unimplemented (fortran
 * order / dtype), Internal (shape or dtype mismatch; only the header's dims are compared).
 * device < 0: host only.
 * ---------------------------------------------------------------------------------------- */
typedef struct nann_huge_const nann_huge_const_t;
nann_status nann_huge_const_create(const char* path, int dtype, const int64_t* shape, int rank,
                                   int device, nann_huge_const_t** out);
const void* nann_huge_const_host(const nann_huge_const_t* h);
const void* nann_huge_const_device(const nann_huge_const_t* h);
int64_t nann_huge_const_bytes(const nann_huge_const_t* h);
void nann_huge_const_destroy(nann_huge_const_t* h);
/* header only: dtype code, rank (<=8) and dims of an npy file */
nann_status nann_npy_peek(const char* path, int* dtype, int* rank, int64_t* shape8);

/* ------------------------------------------------------------------------------------------
 * GroupGather (GroupGather_kernel.cc:18-42): outputs 0 = ret_values (T), 1 = ret_row_splits (i64)
 * InvalidArgument "Invalid RaggedTensor ... code: 1|2|3" as :62-67; void inputs -> values=[],
 * row_splits=[0] (:69-77).  unique!=0 emits each group's distinct values in first-occurrence
 * order (the reference's order is unordered_set iteration order, i.e. unspecified).
 * ---------------------------------------------------------------------------------------- */
nann_status nann_group_gather_i32(const int32_t* params_values, int64_t n_params_values,
                                  const int64_t* params_row_splits, int64_t n_params_row_splits,
                                  const int64_t* indices_values, int64_t n_indices_values,
                                  const int64_t* indices_row_splits, int64_t n_indices_row_splits,
                                  int unique, nann_alloc_fn alloc, void* alloc_ctx, void* stream);
nann_status nann_group_gather_i64(const int64_t* params_values, int64_t n_params_values,
                                  const int64_t* params_row_splits, int64_t n_params_row_splits,
                                  const int64_t* indices_values, int64_t n_indices_values,
                                  const int64_t* indices_row_splits, int64_t n_indices_row_splits,
                                  int unique, nann_alloc_fn alloc, void* alloc_ctx, void* stream);

/* ------------------------------------------------------------------------------------------
 * BitmapRefDifference (bitmap_ops.cc:150-167): outputs 0 = c_values (T), 1 = c_row_splits (i64);
 * idx_flag (int32[n_flags]) is the Ref input: mutated in place and "forwarded" (:179,:238).
 * Order-preserving test-and-set over ALL groups in order, one shared bitmap (:221-234).
 * Ids >= 32*n_flags are InvalidArgument here (the reference writes out of bounds, :225-231).
 * ---------------------------------------------------------------------------------------- */
nann_status nann_bitmap_ref_difference_i32(const int32_t* idx_next_values, int64_t n_values,
                                           const int64_t* idx_next_row_splits, int64_t n_row_splits,
                                           int32_t* idx_flag, int64_t n_flags,
                                           nann_alloc_fn alloc, void* alloc_ctx, void* stream);
nann_status nann_bitmap_ref_difference_i64(const int64_t* idx_next_values, int64_t n_values,
                                           const int64_t* idx_next_row_splits, int64_t n_row_splits,
                                           int32_t* idx_flag, int64_t n_flags,
                                           nann_alloc_fn alloc, void* alloc_ctx, void* stream);

/* ------------------------------------------------------------------------------------------
 * TopKV2 (topk_op.cc:51-93): input [rows, cols] -> values/indices [rows, k]; value descending,
 * ties -> smaller index first (:142-150).  InvalidArgument for k<0 (:60-61) and cols<k (:66-69).
 * sorted==0 is accepted and returns the same (sorted) order, a valid "unsorted" result.
 * ---------------------------------------------------------------------------------------- */
nann_status nann_topk_v2_f32(const float* input, int64_t rows, int64_t cols, int32_t k, int sorted,
                             float* values, int32_t* indices, void* stream);

/* ------------------------------------------------------------------------------------------
 * BatchTopKOnRT (UO/topk_op/BatchTopKOnRT_kernel.cc:25-156): ragged per-group top-k.
 * k: n_k == 1 -> scalar k for every group, else n_k must equal the number of groups (:98-107);
 * each group yields min(len, k) entries (:117-121).  outputs 0 = values_out (f32),
 * 1 = idx_out (i64, GROUP-LOCAL positions, :146), 2 = row_splits_out (i64).
 * ascending != 0 returns the smallest first.  Ties (unspecified by the reference's
 * std::partial_sort_copy) resolve to the smaller position first, TopKV2's rule.
 * Void input (row_splits == [0]) -> [], [], [0] (:88-96).
 * ---------------------------------------------------------------------------------------- */
nann_status nann_batch_topk_on_rt_f32(const float* values_in, int64_t n_values, const int64_t* row_splits_in,
                                      int64_t n_row_splits, const int64_t* k, int64_t n_k, int ascending,
                                      nann_alloc_fn alloc, void* alloc_ctx, void* stream);

/* ------------------------------------------------------------------------------------------
 * Ragged-batch helpers (not in exec.pb; the reference's batch>1 plumbing).  T in {i32, i64}; outputs
 * through the allocator callback: 0 = values (T), 1 = row_splits (i64).
 *   BatchGatherOnRT  UO/beam_search_op/BatchGatherOnRT_kernel.cc:17-97: ret[j] = params[params_rs[g] + idx[j]]
 *   BatchConcatOnRT  UO/beam_search_op/BatchConcatOnRT_kernel.cc:18-115: group g = left[g] ++ right[g]
 *   SplitsGather     UO/beam_search_op/SplitsGather_kernel.cc:20-105: ranges [splits[i], splits[i+1]) expanded
 *   BitmapInit       UO/bitmap_op/bitmap_ops.cc:28-75: bitmap[length] with the bits of idx set
 *   BitmapDifference UO/bitmap_op/bitmap_ops.cc:83-143: like BitmapRefDifference on ONE list, but the flags
 *                    are copied (idx_flag -> idx_flag_new) instead of mutated; output 0 = idx_next_new
 * Error behaviour as the reference's (InvalidArgument on ragged validation / mismatching row_splits;
 * void inputs -> [], [0]).
 * ---------------------------------------------------------------------------------------- */
#define NANN_RAGGED_DECL(T, SFX)                                                                                       \
  nann_status nann_batch_gather_on_rt_##SFX(const T* params_values, int64_t n_pv, const int64_t* params_row_splits,      \
                                            int64_t n_prs, const int64_t* indices_values, int64_t n_iv,                  \
                                            const int64_t* indices_row_splits, int64_t n_irs, nann_alloc_fn alloc,       \
                                            void* alloc_ctx, void* stream);                                              \
  nann_status nann_batch_concat_on_rt_##SFX(const T* left_values, int64_t n_lv, const int64_t* left_row_splits,          \
                                            int64_t n_lrs, const T* right_values, int64_t n_rv,                          \
                                            const int64_t* right_row_splits, int64_t n_rrs, nann_alloc_fn alloc,         \
                                            void* alloc_ctx, void* stream);                                              \
  nann_status nann_splits_gather_##SFX(const T* splits, int64_t n_splits, const int64_t* indices_values, int64_t n_iv,   \
                                       const int64_t* indices_row_splits, int64_t n_irs, nann_alloc_fn alloc,            \
                                       void* alloc_ctx, void* stream);                                                   \
  nann_status nann_bitmap_init_##SFX(const T* idx, int64_t n, int32_t length, int32_t* bitmap, void* stream);            \
  nann_status nann_bitmap_difference_##SFX(const T* idx_next, int64_t n, const int32_t* idx_flag, int64_t n_flags,       \
                                           int32_t* idx_flag_new, nann_alloc_fn alloc, void* alloc_ctx, void* stream);
NANN_RAGGED_DECL(int32_t, i32)
NANN_RAGGED_DECL(int64_t, i64)
#undef NANN_RAGGED_DECL

/* BloomFilterDifference (UO/bitmap_op/bitmap_ops.cc:264-432): BitmapRefDifference with a 4-hash Bloom filter over
 * idx_flag (int32[n_flags], n_flags >= bucket_size; Ref input, mutated in place) instead of an exact bitmap: a value is
 * emitted when at least one of its four bits (Fingerprint64 of the decimal string, optional `% bucket`, the four
 * prime-modulus hashes of :340-351) was still clear; all four are set afterwards.  Values are processed in order over
 * all groups, like the reference's loop.  outputs 0 = c_values (T), 1 = c_row_splits (i64).  InvalidArgument for an
 * invalid ragged input (code 1|2|3) and for n_flags < bucket_size (the reference writes out of bounds). */
nann_status nann_bloom_filter_difference_i32(const int32_t* idx_next_values, int64_t n_values,
                                             const int64_t* idx_next_row_splits, int64_t n_row_splits, int32_t* idx_flag,
                                             int64_t n_flags, int64_t bucket, int64_t bucket_size, nann_alloc_fn alloc,
                                             void* alloc_ctx, void* stream);
nann_status nann_bloom_filter_difference_i64(const int64_t* idx_next_values, int64_t n_values,
                                             const int64_t* idx_next_row_splits, int64_t n_row_splits, int32_t* idx_flag,
                                             int64_t n_flags, int64_t bucket, int64_t bucket_size, nann_alloc_fn alloc,
                                             void* alloc_ctx, void* stream);

/* GatherV2 on axis 0: out[i] = table[ids[i]], row_bytes per row (build_opt_graph.py:92,144).
 * InvalidArgument when an id is outside [0, n_rows). */
nann_status nann_gather_rows(const void* table, int64_t n_rows, int64_t row_bytes,
                             const int32_t* ids, int64_t n, void* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Scorer = what BlazeXlaOp runs (a nested session over frozen_graph.pb in the reference).
 *   mlp:        s(u,x) = w3 . relu(W2 . relu(W1 [u;x] + b1) + b2)   (BASELINE configs 2-5)
 *               W1 [H][2d] row-major (cols 0..d-1 act on u), W2 [H][H], w3 [H]; d=128, H=512.
 *   attention:  Model.forward, NANN_impls/nann/model/model.py:189-233 (config 1); weights as one
 *               fp32 blob in the order documented in nann_b200/scorer_weights.py (BN folded).
 * precision: NANN_SCORER_EXACT  = fp32 FFMA, every dot a sequential chain in k (bit-identical
 *                                 to the oracle's definition);
 *            NANN_SCORER_TENSOR = tcgen05 tensor cores, fp16 hi/lo split operands, fp32
 *                                 accumulate (|score - exact| <= 1e-5); mlp only.
 * ---------------------------------------------------------------------------------------- */
typedef struct nann_scorer nann_scorer_t;
enum { NANN_SCORER_EXACT = 0, NANN_SCORER_TENSOR = 1 };
nann_status nann_scorer_create_mlp(int d, int H, const float* W1, const float* b1, const float* W2,
                                   const float* b2, const float* w3, int device, nann_scorer_t** out);
nann_status nann_scorer_create_attention(const float* blob, int64_t n_floats, int device,
                                         nann_scorer_t** out);
int64_t nann_scorer_attention_blob_size(void);
nann_status nann_scorer_set_precision(nann_scorer_t* s, int precision);
int nann_scorer_user_floats(const nann_scorer_t* s);  /* 128 (mlp) or 3200 (attention) */
int nann_scorer_item_dim(const nann_scorer_t* s);     /* 128 or 64 */
void nann_scorer_destroy(nann_scorer_t* s);
/* BlazeXlaOp::Compute with inputs [user, item_emb[n, d]] -> logits[n] (fp32 in, fp32 out) */
nann_status nann_blaze_xla_run(nann_scorer_t* s, const float* user, const float* item_emb, int64_t n,
                               float* logits, void* stream);
/* Admission control of the two run calls, as BlazeXlaOp::Schedule (blaze_xla_kernel.cc:221-258): at most running_max
 * runs at a time (default: env BLAZE_THREADS_NUM or 2, :87-93); a call that finds the scorer busy blocks -- with
 * wait_ms > 0 (BlazeKernelOptions.wait_ms, config.proto:840) for at most that long, then fails with
 * NANN_INTERNAL "blaze wait too long" (still queued) or NANN_DEADLINE_EXCEEDED "blaze wait too long" (admitted too
 * late); with wait_ms == 0 it fails at once with NANN_INTERNAL "waiting pool is full" when max_waiting calls (default:
 * env DENSE_MAX_WAITING_COUNT or 10, :95-101) are queued already.  A negative argument leaves that setting alone. */
nann_status nann_scorer_set_admission(nann_scorer_t* s, int running_max, int max_waiting, int wait_ms);
nann_status nann_scorer_admission_state(nann_scorer_t* s, int* running, int* waiting, int* running_max,
                                        int* max_waiting, int* wait_ms);
/* fused GatherV2 + BlazeXlaOp: logits[i] = score(user, table[ids[i]]); table f32 [n_rows][d] */
nann_status nann_scorer_run_ids(nann_scorer_t* s, const float* user, const float* table,
                                int64_t n_rows, const int32_t* ids, int64_t n, float* logits,
                                void* stream);

/* ------------------------------------------------------------------------------------------
 * Index: the Appendix-C files (build_hnsw_index.py:33-67) resident in HBM.
 *   emb f32[n_items][dim] (f16 accepted via emb_dtype, widened once), item_ids i64[n_items],
 *   enter_points (i32 or i64, ascending, unique), per level l in {0,1}: CSR values (i32 or i64,
 *   narrowed once like build_opt_graph.py:87) + row_splits i64[n_items+1].
 * Arrays may be host or device pointers; the index keeps its own device copy.
 * ---------------------------------------------------------------------------------------- */
typedef struct nann_index nann_index_t;
nann_status nann_index_create(int64_t n_items, int dim, const void* emb, int emb_dtype,
                              const int64_t* item_ids, const void* enter_points, int ep_dtype,
                              int64_t n_enter_points, const void* const nbr_values[2],
                              const int64_t n_nbr_values[2], int nbr_dtype,
                              const int64_t* const nbr_row_splits[2], int device, nann_index_t** out);
/* loads item_embs.npy, item_ids.npy from embs_dir and enter_points.npy,
 * neighbors_level_{0,1}_{values,row_splits}.npy from index_dir through nann_huge_const_create */
nann_status nann_index_load(const char* embs_dir, const char* index_dir, int device,
                            nann_index_t** out);
int64_t nann_index_n_items(const nann_index_t* ix);
int nann_index_dim(const nann_index_t* ix);
int64_t nann_index_n_enter_points(const nann_index_t* ix);
const float* nann_index_emb_device(const nann_index_t* ix);
void nann_index_destroy(nann_index_t* ix);

/* ------------------------------------------------------------------------------------------
 * Search: exec.pb's dataflow (build_opt_graph.py:109-149) for B queries per call, all on device.
 * level_topn[6] is a per-call input like the reference's placeholder (:75), shared by the batch.
 * Per query q: out_item_ids[q][0..k) (k = level_topn[5], i64, the op's 'top_k' output :149),
 * out_scores[q][0..k) (additional), out_status[q] = NANN_OK or NANN_INVALID_ARGUMENT where the
 * reference's session.run would fail (TopKV2 n<k, topk_op.cc:66-69; exactly one candidate to
 * score, the squeeze-to-scalar case of build_opt_graph.py:107).
 * users: [B][nann_scorer_user_floats].  Host or device pointers throughout.
 * ---------------------------------------------------------------------------------------- */
typedef struct nann_searcher nann_searcher_t;
typedef struct {
  int64_t n_scored[5];    /* rows scored per round, summed over the batch (main.py:179-185) */
  int64_t n_expanded[5];  /* ids GroupGather produced per round (level 2: enter points) */
  int64_t n_failed;       /* queries with status != OK */
} nann_search_stats_t;
nann_status nann_searcher_create(const nann_index_t* ix, nann_scorer_t* scorer, int max_batch,
                                 const int32_t max_level_topn[6], nann_searcher_t** out);
void nann_searcher_destroy(nann_searcher_t* s);
/* keep per-round (node ids, scores) of every query for parity checks (costly; off by default) */
nann_status nann_searcher_set_trace(nann_searcher_t* s, int enable);
nann_status nann_search_batch(nann_searcher_t* s, const float* users, int B,
                              const int32_t level_topn[6], int64_t* out_item_ids,
                              float* out_scores, int32_t* out_status, nann_search_stats_t* stats,
                              void* stream);
/* after a traced call: copies round r (0..4) of query q to host: ids/scores capacity cap */
nann_status nann_searcher_get_trace(nann_searcher_t* s, int q, int round, int32_t* ids,
                                    float* scores, int64_t cap, int64_t* n);
/* Stage timing with CUDA events on the launching stream, accumulated over calls while enabled
 * (enable resets).  stage: 0 = scorer (row gather + model), 1 = expand+filter, 2 = top-k,
 * 3 = bitmap reset+mark.  rows_scored = rows the scorer processed in those calls. */
nann_status nann_searcher_set_profile(nann_searcher_t* s, int enable);
nann_status nann_searcher_get_profile(nann_searcher_t* s, double stage_ms[4], int64_t stage_launches[4],
                                      int64_t* rows_scored, int64_t* calls);

/* ---- the `main.py --job-type test` traversal ----------------------------------
 * Replaces Model.retrieval / Model.search_level (NANN_impls/nann/model/model.py:299-362), the graph
 * `main.py test` runs once per user (main.py:150-190), for a batch of users: start level = 2, every
 * enter point scored once, then per level the UNIQUE unvisited neighbours of the frontier in ascending
 * id order (tf.unique + tf.sets.set_difference, :319-322), merged into the running result with a
 * clamped k (:268, :329-331), next frontier = new nodes whose score reaches the worst kept score
 * (:333-334).  num_scoring_per_level / top_k_per_level are indexed by LEVEL 0..2 like the reference's
 * flags (nann/config.py:52-55, defaults [3,1,1] / [400,200,100]); topk_eval = --topk-eval.
 * out_item_ids/out_scores/out_nodes [B][topk_eval] (-1 / 0 padded), out_n[B] = valid results per
 * query, out_status[B]: InvalidArgument where the reference's graph fails (a round with exactly one
 * new candidate: tf.squeeze yields a scalar that cannot be concatenated), ResourceExhausted when
 * a round produced more candidates than the workspace holds.  Host or device pointers. */
typedef struct nann_eval_searcher nann_eval_searcher_t;
nann_status nann_eval_searcher_create(const nann_index_t* ix, nann_scorer_t* scorer, int max_batch,
                                      const int32_t max_top_k_per_level[3], int max_topk_eval,
                                      nann_eval_searcher_t** out);
void nann_eval_searcher_destroy(nann_eval_searcher_t* s);
nann_status nann_search_eval_batch(nann_eval_searcher_t* s, const float* users, int B,
                                   const int32_t num_scoring_per_level[3], const int32_t top_k_per_level[3],
                                   int topk_eval, int64_t* out_item_ids, float* out_scores, int32_t* out_nodes,
                                   int32_t* out_n, int32_t* out_status, int64_t* n_scored_total, void* stream);
/* node ids (rows of the table) of the last call's final top-k, [B][k], before the item_ids gather */
nann_status nann_searcher_get_nodes(nann_searcher_t* s, int32_t* out_nodes, int64_t cap);

/* debug: device buffer of 64*48 int64 that CTA 0 of the tensor-core scorer fills with clock64()
 * stamps per (tile, pipeline event); NULL turns it off (scripts/tc_timeline.py) */
nann_status nann_debug_tc_trace(long long* device_buffer_64x48);

/* ------------------------------------------------------------------------------------------
 * Shard merge: G per-shard results [G][B][k_in] (score f32, id i64), as laid out by
 * an allgather, -> global top k_out per query.  Order: score descending, ties -> lower shard,
 * then lower per-shard rank.  InvalidArgument if G*k_in < k_out.
 * ---------------------------------------------------------------------------------------- */
nann_status nann_merge_topk(const float* scores, const int64_t* ids, int G, int B, int k_in,
                            int k_out, float* out_scores, int64_t* out_ids, void* stream);

/* ------------------------------------------------------------------------------------------
 * Sharded search: the corpus is row-sharded over `world` ranks (one per GPU), each with its own
 * index + searcher; every query visits every shard and the per-shard top-k are exchanged and merged.  The
 * exchange is part of the library: every rank owns a receive window in its HBM, maps the peers' windows
 * (CUDA IPC across processes -- export / connect -- or nann_shard_group_connect_local for several members in
 * one process) and the final top-k kernel of nann_search_sharded stores its (score, item id) records straight
 * into every rank's window over NVLink; a merge kernel on the group's own stream then produces the global
 * top k_out (order as nann_merge_topk: score descending, ties -> lower shard, then lower per-shard rank).
 * Every rank must make the same sequence of nann_search_sharded calls (same B, level_topn_shard, k_out): it is
 * a collective.  A query that failed on any shard fails as a whole (status of the first failing shard; ids -1).
 * Outputs [B][k_out] / [B]:
 *   - host pointers: the call returns when the merged results are in them;
 *   - device pointers: the call only enqueues work (searches on `stream`, exchange + merge on the group's
 *     stream, so that the next call's search overlaps them).  The outputs are complete once work ordered by
 *     nann_shard_group_wait has run; give consecutive calls different output buffers.
 * nann_shard_group_wait: host_block != 0 blocks the host until every merge issued so far has finished
 * (DeadlineExceeded if a peer never delivered); otherwise makes `stream` wait for them.
 * ---------------------------------------------------------------------------------------- */
typedef struct nann_shard_group nann_shard_group_t;
nann_status nann_shard_group_create(int device, int rank, int world, int max_batch, int max_k_shard,
                                    nann_shard_group_t** out);
int nann_shard_group_handle_bytes(void);  /* 64: size of the opaque window handle (cudaIpcMemHandle_t) */
nann_status nann_shard_group_export(nann_shard_group_t* g, void* handle);
/* handles: [world][nann_shard_group_handle_bytes()] in rank order (exchanged out of band, e.g. an
 * all-gather over torch.distributed / MPI); the own entry is ignored */
nann_status nann_shard_group_connect(nann_shard_group_t* g, const void* handles);
nann_status nann_shard_group_connect_local(nann_shard_group_t* const* members, int world);
void nann_shard_group_destroy(nann_shard_group_t* g);
nann_status nann_search_sharded(nann_searcher_t* s, nann_shard_group_t* g, const float* users, int B,
                                const int32_t level_topn_shard[6], int k_out, int64_t* out_item_ids,
                                float* out_scores, int32_t* out_status, void* stream);
/* The two phases of nann_search_sharded as separate calls, for ONE host thread that drives several members
 * (shards on one GPU, or one thread per box driving all GPUs): push every member first, then merge every member.
 * (A member's merge waits on the device for the other members' pushes; enqueueing it before those pushes exist
 * can stall the GPU's work queues behind the waiting kernel.)  At most one push may be pending per group. */
nann_status nann_search_sharded_push(nann_searcher_t* s, nann_shard_group_t* g, const float* users, int B,
                                     const int32_t level_topn_shard[6], void* stream);
nann_status nann_search_sharded_merge(nann_shard_group_t* g, int k_out, int64_t* out_item_ids, float* out_scores,
                                      int32_t* out_status);
nann_status nann_shard_group_wait(nann_shard_group_t* g, void* stream, int host_block);

/* ------------------------------------------------------------------------------------------
 * Distributed scoring: the second multi-GPU form (csrc/lib_dist.inl).  ONE graph (replicated, like the enter points
 * and item ids), the embedding table row-sharded -- rank r of `world` holds rows [r*per, (r+1)*per), per =
 * ceil(n_items / world), nann_index_create_sharded -- and the global batch partitioned: every rank runs the
 * traversal of ITS B queries; each scoring round the candidates are sent to the ranks that own their rows, scored there
 * with the ordinary scorer kernel, and the scores come back (stores into IPC-mapped peer windows over NVLink + flags,
 * no host synchronisation).  The result is BIT-IDENTICAL to nann_search_batch on the unsharded index; per-GPU scoring
 * work is 1/world of it.  Every rank must call nann_search_distributed with the same B and level_topn (collective);
 * out_* are this rank's queries only.  Both scorers (the attention scorer's raw user sequences travel with the key
 * projections, once per call).  Device outputs: the call only enqueues on `stream`
 * (nann_dist_group_check reports a timed-out exchange); host outputs: it returns when they are filled -- every rank
 * then needs its own thread or process.
 * ---------------------------------------------------------------------------------------- */
nann_status nann_index_create_sharded(int64_t n_items, int dim, const void* emb_rows, int emb_dtype, int64_t row_lo,
                                      int64_t n_rows_local, const int64_t* item_ids, const void* enter_points,
                                      int ep_dtype, int64_t n_enter_points, const void* const nbr_values[2],
                                      const int64_t n_nbr_values[2], int nbr_dtype,
                                      const int64_t* const nbr_row_splits[2], int device, nann_index_t** out);
typedef struct nann_dist_group nann_dist_group_t;
nann_status nann_dist_group_create(const nann_searcher_t* s, int rank, int world, nann_dist_group_t** out);
nann_status nann_dist_group_export(nann_dist_group_t* g, void* handle);   /* nann_shard_group_handle_bytes() bytes */
nann_status nann_dist_group_connect(nann_dist_group_t* g, const void* handles);
nann_status nann_dist_group_connect_local(nann_dist_group_t* const* members, int world);
nann_status nann_dist_group_check(nann_dist_group_t* g);
void nann_dist_group_destroy(nann_dist_group_t* g);
nann_status nann_search_distributed(nann_searcher_t* s, nann_dist_group_t* g, const float* users, int B,
                                    const int32_t level_topn[6], int64_t* out_item_ids, float* out_scores,
                                    int32_t* out_status, nann_search_stats_t* stats, void* stream);

/* ------------------------------------------------------------------------------------------
 * Index construction.  The reference builds its graph offline with faiss
 * IndexHNSWFlat(d, M) and dumps per-level CSR files (NANN_impls/nann/delivery/build_hnsw_index.py:33-67);
 * this builds the same FILES on the GPU: per level l < n_levels the links of the nodes whose level reaches l
 * (levels[i] = highest level of node i, 0-based, drawn by the caller with faiss's distribution), at most 2M
 * links at level 0 and M above, rows closest-first.  Batch construction instead of faiss's sequential
 * insertion: exact k-NN candidates (tensor-core brute force + fp32 refinement), HNSW's diversity heuristic for
 * the forward links, reverse links, truncation -- so the graph is a valid HNSW in the reference's layout, not
 * faiss's graph (which no reference test pins).  dim must be 128, 2 <= M <= 32.
 * Outputs through the allocator callback (host or device memory): 2*l = values of level l (int32 node ids),
 * 2*l+1 = row_splits of level l (int64 [n+1]).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  double seconds_knn;        /* candidate search + refinement + heuristic, all levels */
  double seconds_links;      /* reverse links, de-duplication, truncation */
  int64_t n_forward_links;
  int64_t n_overflow;        /* candidate pairs dropped because a row's append buffer was full (0 on shuffled data) */
} nann_hnsw_build_stats_t;
nann_status nann_hnsw_build(const float* emb, int64_t n, int dim, const int32_t* levels, int M, int n_levels,
                            int device, nann_alloc_fn alloc, void* alloc_ctx, nann_hnsw_build_stats_t* stats);

/* ------------------------------------------------------------------------------------------
 * Executor: blaze-benchmark's load generator + session pool around the search call
 * (blaze-benchmark/benchmark/proto/bench_conf.proto:5-42; core/benchmark.cc:101-146;
 * core/model.cc:19-53,192-234; core/predict_request_consumer.cc:17-54; core/metrics.cc:5-94).
 * predictor_num searchers, each with its own CUDA stream and consumer thread, replace the
 * reference's sessions on virtual GPUs / CUDA contexts; a consumer coalesces up to max_batch_size
 * queued requests into one nann_search_batch call (1 = the reference: one run per request).
 * Histograms are {count,min,max,mean,stddev,median,p75,p95,p98,p99,p99.9} like cppmetrics'
 * ConsoleReporter; latency_us = duration of the run a request was part of (what the reference
 * records), e2e_latency_us = enqueue -> completion.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  int predictor_num;       /* bench_conf.proto:21-22 */
  int bench_thread_count;  /* :36-37 producer threads */
  double duration_s;       /* :38-39 */
  int qps;                 /* :23-24, <= 0 = maximum (closed loop) */
  int max_queue_size;      /* :40-41, <= 0 never drop */
  int max_batch_size;      /* new: dynamic batching width */
  int batch_timeout_us;    /* new: how long a consumer waits for a fuller batch (0 = take what is queued) */
  int report_interval_s;   /* 3 s in the reference (metrics.cc:15) */
} nann_bench_conf_t;
typedef struct {
  double seconds;
  int64_t throughput_count;        /* "<model>_throughput" meter */
  double mean_rate;                /* events/second */
  int64_t failures;                /* "<model>_failures" */
  int64_t get_predictor_failures;  /* "<model>_get_predictor_failures" (dropped: queue too long) */
  double latency_us[11];
  double e2e_latency_us[11];
  double batchsize[11];
} nann_bench_report_t;
/* queries: HOST memory [n_queries][user_floats], replayed round-robin like mock.runmeta */
nann_status nann_executor_run(const nann_index_t* ix, nann_scorer_t* scorer, const nann_bench_conf_t* conf,
                              const int32_t level_topn[6], const float* queries, int64_t n_queries,
                              int print_reports, nann_bench_report_t* report);

#ifdef __cplusplus
}
#endif
#endif /* NANN_B200_H_ */
