/*
 * nann_oracle.c -- CPU restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY
 * (see nann_oracle.h).  Plain C11 (+ AVX2/FMA intrinsics for the blocked scorer, OpenMP for
 * the request-parallel baseline).  Every function cites the reference lines it follows;
 * paths are relative to /root/reference, UO = tensorflow/tensorflow/core/user_ops.
 */
#define _GNU_SOURCE
#include "nann_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <pthread.h>
#include <stdatomic.h>
#include <unistd.h>
#if defined(__AVX2__) && defined(__FMA__)
#include <immintrin.h>
#define ORC_AVX2 1
#else
#define ORC_AVX2 0
#endif

const char* orc_version(void) { return "nann-oracle 1 (avx2="
#if ORC_AVX2
  "1"
#else
  "0"
#endif
  ")"; }

/* ------------------------------------------------------------------------------------------
 * ragged validation -- GroupGather_kernel.cc:9-16 (same helper in bitmap_ops.cc)
 * ---------------------------------------------------------------------------------------- */
int orc_validate_ragged(int64_t n_values, const int64_t* row_splits, int64_t n_row_splits) {
  if (n_row_splits == 0) return 1;
  if (row_splits[0] != 0) return 2;
  if (row_splits[n_row_splits - 1] != n_values) return 3;
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * GroupGather -- GroupGather_kernel.cc:55-170
 *   :62-67  validate both ragged inputs (InvalidArgument, code 1/2/3)
 *   :69-77  void inputs -> values=[], row_splits=[0]
 *   :136-149 pass 1: count; :151-168 pass 2: fill in index order
 *   :91-131 unique: per-group set; order unspecified in the reference (unordered_set) ->
 *           we emit first-occurrence order.
 * ---------------------------------------------------------------------------------------- */
#define ORC_GROUP_GATHER(NAME, T)                                                              \
  int NAME(const T* pv, int64_t n_pv, const int64_t* prs, int64_t n_prs, const int64_t* iv,     \
           int64_t n_iv, const int64_t* irs, int64_t n_irs, int unique, T* ret_values,          \
           int64_t* n_ret, int64_t* ret_row_splits, int64_t* n_ret_rs, int* code) {             \
    int valid = orc_validate_ragged(n_pv, prs, n_prs);                                          \
    if (code) *code = valid;                                                                    \
    if (valid != 0) return ORC_INVALID_ARGUMENT;                                                \
    valid = orc_validate_ragged(n_iv, irs, n_irs);                                              \
    if (code) *code = valid;                                                                    \
    if (valid != 0) return ORC_INVALID_ARGUMENT;                                                \
    if (n_prs == 1 || n_irs == 1) {                                                             \
      *n_ret = 0;                                                                               \
      ret_row_splits[0] = 0;                                                                    \
      *n_ret_rs = 1;                                                                            \
      return ORC_OK;                                                                            \
    }                                                                                           \
    int64_t num_groups = n_irs - 1;                                                             \
    *n_ret_rs = n_irs;                                                                          \
    ret_row_splits[0] = 0;                                                                      \
    if (!unique) {                                                                              \
      int64_t sum = 0;                                                                          \
      for (int64_t i = 0; i < num_groups; ++i) {                                                \
        for (int64_t j = irs[i]; j < irs[i + 1]; ++j) {                                         \
          int64_t idx = iv[j];                                                                  \
          sum += prs[idx + 1] - prs[idx];                                                       \
        }                                                                                       \
        ret_row_splits[i + 1] = sum;                                                            \
      }                                                                                         \
      *n_ret = sum;                                                                             \
      if (!ret_values) return ORC_OK;                                                           \
      for (int64_t i = 0; i < num_groups; ++i) {                                                \
        int64_t o = ret_row_splits[i];                                                          \
        for (int64_t j = irs[i]; j < irs[i + 1]; ++j) {                                         \
          int64_t g = iv[j];                                                                    \
          for (int64_t k = prs[g]; k < prs[g + 1]; ++k) ret_values[o++] = pv[k];                \
        }                                                                                       \
      }                                                                                         \
      return ORC_OK;                                                                            \
    }                                                                                           \
    /* unique: O(n^2) per group is fine for an oracle */                                        \
    int64_t cap = 0;                                                                            \
    for (int64_t j = 0; j < n_iv; ++j) cap += prs[iv[j] + 1] - prs[iv[j]];                       \
    T* tmp = (T*)malloc((size_t)(cap > 0 ? cap : 1) * sizeof(T));                               \
    int64_t o = 0;                                                                              \
    for (int64_t i = 0; i < num_groups; ++i) {                                                  \
      int64_t start = o;                                                                        \
      for (int64_t j = irs[i]; j < irs[i + 1]; ++j) {                                           \
        int64_t g = iv[j];                                                                      \
        for (int64_t k = prs[g]; k < prs[g + 1]; ++k) {                                         \
          T v = pv[k];                                                                          \
          int seen = 0;                                                                         \
          for (int64_t q = start; q < o; ++q)                                                   \
            if (tmp[q] == v) { seen = 1; break; }                                               \
          if (!seen) tmp[o++] = v;                                                              \
        }                                                                                       \
      }                                                                                         \
      ret_row_splits[i + 1] = o;                                                                \
    }                                                                                           \
    *n_ret = o;                                                                                 \
    if (ret_values) memcpy(ret_values, tmp, (size_t)o * sizeof(T));                             \
    free(tmp);                                                                                  \
    return ORC_OK;                                                                              \
  }
ORC_GROUP_GATHER(orc_group_gather_i32, int32_t)
ORC_GROUP_GATHER(orc_group_gather_i64, int64_t)

/* ------------------------------------------------------------------------------------------
 * BitmapRefDifference -- bitmap_ops.cc:175-257
 *   :182-184 validate; :187-196 void input; :221-234 Differ: for every value in group order,
 *   flag_index = node>>5, bit = node&31; keep iff bit unset, then set it.  All groups share the
 *   one bitmap (groups are processed in order, :236 Differ(0,num_groups)).
 * ---------------------------------------------------------------------------------------- */
#define ORC_BITMAP_DIFF(NAME, T)                                                               \
  int NAME(const T* v, int64_t n_v, const int64_t* rs, int64_t n_rs, int32_t* flags,            \
           int64_t n_flags, int check_bounds, T* c_values, int64_t* n_c, int64_t* c_row_splits, \
           int64_t* n_c_rs, int* code) {                                                        \
    int valid = orc_validate_ragged(n_v, rs, n_rs);                                             \
    if (code) *code = valid;                                                                    \
    if (valid != 0) return ORC_INVALID_ARGUMENT;                                                \
    if (n_rs == 1) {                                                                            \
      *n_c = 0;                                                                                 \
      c_row_splits[0] = 0;                                                                      \
      *n_c_rs = 1;                                                                              \
      return ORC_OK;                                                                            \
    }                                                                                           \
    int64_t num_groups = n_rs - 1;                                                              \
    uint32_t* f = (uint32_t*)flags;                                                             \
    int64_t o = 0;                                                                              \
    c_row_splits[0] = 0;                                                                        \
    for (int64_t i = 0; i < num_groups; ++i) {                                                  \
      for (int64_t j = rs[i]; j < rs[i + 1]; ++j) {                                             \
        T node = v[j];                                                                          \
        int64_t w = (int64_t)(node >> 5);                                                       \
        uint32_t bit = 1u << (uint32_t)(node & 31);                                             \
        if (check_bounds && (w < 0 || w >= n_flags)) return ORC_INVALID_ARGUMENT;               \
        if (!(f[w] & bit)) {                                                                    \
          c_values[o++] = node;                                                                 \
          f[w] |= bit;                                                                          \
        }                                                                                       \
      }                                                                                         \
      c_row_splits[i + 1] = o;                                                                  \
    }                                                                                           \
    *n_c = o;                                                                                   \
    *n_c_rs = n_rs;                                                                             \
    return ORC_OK;                                                                              \
  }
ORC_BITMAP_DIFF(orc_bitmap_ref_difference_i32, int32_t)
ORC_BITMAP_DIFF(orc_bitmap_ref_difference_i64, int64_t)

/* ------------------------------------------------------------------------------------------
 * TopKV2 -- topk_op.cc:51-93 (checks) and :139-207 (CPU functor).
 * stable_comp(a,b) (:142-150): a precedes b iff in[a] > in[b], or equal and a < b.  TopN keeps
 * the k best under it and Extract() returns best-first (lib/gtl/top_n.h:128-131,279-288); the
 * k==cols branch (:159-180) is std::sort + tie-run index sort = the same total order.  Since the
 * comparator is a strict total order on indices (NaN aside) the result is algorithm-independent;
 * we use a bounded binary heap + heap sort.
 * ---------------------------------------------------------------------------------------- */
static inline int topk_before(const float* in, int32_t a, int32_t b) {
  if (in[b] < in[a]) return 1;
  if (in[b] > in[a]) return 0;
  return a < b;
}

/* heap with the WORST kept element at the root (so root is evicted first) */
static void topk_sift_down(const float* in, int32_t* h, int64_t n, int64_t i) {
  for (;;) {
    int64_t l = 2 * i + 1, r = l + 1, w = i;
    if (l < n && topk_before(in, h[w], h[l])) w = l; /* h[l] is worse than h[w] */
    if (r < n && topk_before(in, h[w], h[r])) w = r;
    if (w == i) return;
    int32_t t = h[i]; h[i] = h[w]; h[w] = t;
    i = w;
  }
}

static void topk_row(const float* in, int64_t cols, int k, float* values, int32_t* indices) {
  int32_t* h = indices; /* build in place */
  for (int i = 0; i < k; ++i) h[i] = i;
  for (int64_t i = k / 2 - 1; i >= 0; --i) topk_sift_down(in, h, k, i);
  for (int64_t c = k; c < cols; ++c) {
    if (topk_before(in, (int32_t)c, h[0])) { /* c beats current worst */
      h[0] = (int32_t)c;
      topk_sift_down(in, h, k, 0);
    }
  }
  /* heap sort: repeatedly move the worst to the end -> best-first order */
  for (int64_t n = k; n > 1; --n) {
    int32_t t = h[0]; h[0] = h[n - 1]; h[n - 1] = t;
    topk_sift_down(in, h, n - 1, 0);
  }
  for (int i = 0; i < k; ++i) values[i] = in[indices[i]];
}

int orc_topk_v2_f32(const float* input, int64_t rows, int64_t cols, int k, float* values,
                    int32_t* indices) {
  if (k < 0) return ORC_INVALID_ARGUMENT;    /* topk_op.cc:60-61 "Need k >= 0" */
  if (cols < k) return ORC_INVALID_ARGUMENT; /* topk_op.cc:66-69 */
  if (k == 0 || rows == 0) return ORC_OK;    /* topk_op.cc:84-85 */
  for (int64_t r = 0; r < rows; ++r)
    topk_row(input + r * cols, cols, k, values + r * (int64_t)k, indices + r * (int64_t)k);
  return ORC_OK;
}

/* ------------------------------------------------------------------------------------------
 * BatchTopKOnRT -- BatchTopKOnRT_kernel.cc:62-156
 * ---------------------------------------------------------------------------------------- */
int orc_batch_topk_on_rt_f32(const float* values, int64_t n_values, const int64_t* rs, int64_t n_rs,
                             const int64_t* k, int64_t n_k, int ascending, float* values_out,
                             int64_t* idx_out, int64_t* rs_out, int64_t* n_out, int* code) {
  int valid = orc_validate_ragged(n_values, rs, n_rs);              /* :75-77 */
  if (code) *code = valid;
  if (valid != 0) return ORC_INVALID_ARGUMENT;
  int64_t G = n_rs - 1;
  rs_out[0] = 0;
  *n_out = 0;
  if (G == 0) return ORC_OK;                                        /* :88-96 */
  if (n_k != 1 && n_k != G) return ORC_INVALID_ARGUMENT;            /* :103-105 */
  int64_t o = 0;
  for (int64_t g = 0; g < G; ++g) {
    int64_t b = rs[g], len = rs[g + 1] - rs[g];
    int64_t kk = n_k == 1 ? k[0] : k[g];
    if (kk > len) kk = len;                                         /* :119 */
    if (kk < 0) kk = 0;
    /* selection sort of the kk best under (value, position) -- small oracle, clarity over speed */
    char* used = (char*)calloc((size_t)(len > 0 ? len : 1), 1);
    for (int64_t r = 0; r < kk; ++r) {
      int64_t best = -1;
      for (int64_t i = 0; i < len; ++i) {
        if (used[i]) continue;
        if (best < 0) { best = i; continue; }
        float vi = values[b + i], vb = values[b + best];
        if (ascending ? (vi < vb) : (vi > vb)) best = i;           /* strict: earlier position wins ties */
      }
      used[best] = 1;
      values_out[o] = values[b + best];
      idx_out[o] = best;                                            /* group-local (:146) */
      ++o;
    }
    free(used);
    rs_out[g + 1] = o;
  }
  *n_out = o;
  return ORC_OK;
}

/* ------------------------------------------------------------------------------------------
 * GatherV2 (stock) as used at build_opt_graph.py:92 (rows) and :144 (item ids)
 * ---------------------------------------------------------------------------------------- */
void orc_gather_rows_f32(const float* table, int64_t dim, const int32_t* ids, int64_t n, float* out) {
  for (int64_t i = 0; i < n; ++i) memcpy(out + i * dim, table + (int64_t)ids[i] * dim, (size_t)dim * sizeof(float));
}
void orc_gather_i64(const int64_t* table, const int32_t* ids, int64_t n, int64_t* out) {
  for (int64_t i = 0; i < n; ++i) out[i] = table[ids[i]];
}

/* ------------------------------------------------------------------------------------------
 * mlp2x512 scorer.  Definition in nann_oracle.h.  The blocked kernel keeps, per output
 * element, the same sequential fmaf chain (k ascending), so it is bit-identical to
 * orc_mlp_score_def; only independent outputs are vectorised.
 * ---------------------------------------------------------------------------------------- */
struct orc_mlp {
  int d, H;
  float *W1, *b1, *W2, *b2, *w3; /* as given (row-major [out][in]) */
  float *W1uT;                   /* [d][H]  : W1uT[k][j] = W1[j][k]       */
  float *W1xT;                   /* [d][H]  : W1xT[k][j] = W1[j][d+k]     */
  float *W2T;                    /* [H][H]  : W2T[k][j]  = W2[j][k]       */
};

static float* dup_f32(const float* p, size_t n) {
  float* q = (float*)aligned_alloc(64, ((n * sizeof(float) + 63) / 64) * 64);
  memcpy(q, p, n * sizeof(float));
  return q;
}

orc_mlp_t* orc_mlp_create(int d, int H, const float* W1, const float* b1, const float* W2,
                          const float* b2, const float* w3) {
  if (H % 16 != 0) return NULL;
  orc_mlp_t* m = (orc_mlp_t*)calloc(1, sizeof(*m));
  m->d = d; m->H = H;
  m->W1 = dup_f32(W1, (size_t)H * 2 * d); m->b1 = dup_f32(b1, H);
  m->W2 = dup_f32(W2, (size_t)H * H);     m->b2 = dup_f32(b2, H);
  m->w3 = dup_f32(w3, H);
  m->W1uT = (float*)aligned_alloc(64, (size_t)d * H * sizeof(float));
  m->W1xT = (float*)aligned_alloc(64, (size_t)d * H * sizeof(float));
  m->W2T = (float*)aligned_alloc(64, (size_t)H * H * sizeof(float));
  for (int j = 0; j < H; ++j) {
    for (int k = 0; k < d; ++k) {
      m->W1uT[(size_t)k * H + j] = W1[(size_t)j * 2 * d + k];
      m->W1xT[(size_t)k * H + j] = W1[(size_t)j * 2 * d + d + k];
    }
    for (int k = 0; k < H; ++k) m->W2T[(size_t)k * H + j] = W2[(size_t)j * H + k];
  }
  return m;
}

void orc_mlp_destroy(orc_mlp_t* m) {
  if (!m) return;
  free(m->W1); free(m->b1); free(m->W2); free(m->b2); free(m->w3);
  free(m->W1uT); free(m->W1xT); free(m->W2T);
  free(m);
}

static inline float relu_f(float a) { return a > 0.0f ? a : 0.0f; }

void orc_mlp_score_def(const orc_mlp_t* m, const float* u, const float* x, int64_t n, float* out) {
  const int d = m->d, H = m->H;
  float* h1 = (float*)malloc(sizeof(float) * H);
  float* h2 = (float*)malloc(sizeof(float) * H);
  for (int64_t i = 0; i < n; ++i) {
    const float* xi = x + i * d;
    for (int j = 0; j < H; ++j) {
      float a = m->b1[j];
      const float* w = m->W1 + (size_t)j * 2 * d;
      for (int k = 0; k < d; ++k) a = fmaf(w[k], u[k], a);
      for (int k = 0; k < d; ++k) a = fmaf(w[d + k], xi[k], a);
      h1[j] = relu_f(a);
    }
    for (int j = 0; j < H; ++j) {
      float a = m->b2[j];
      const float* w = m->W2 + (size_t)j * H;
      for (int k = 0; k < H; ++k) a = fmaf(w[k], h1[k], a);
      h2[j] = relu_f(a);
    }
    float s = 0.0f;
    for (int k = 0; k < H; ++k) s = fmaf(m->w3[k], h2[k], s);
    out[i] = s;
  }
  free(h1); free(h2);
}

/* hu[j] = b1[j] (+) sum_k W1[j][k] u[k]  -- the per-query prefix of every layer-1 chain */
static void mlp_hoist(const orc_mlp_t* m, const float* u, float* hu) {
  const int d = m->d, H = m->H;
  for (int j = 0; j < H; ++j) hu[j] = m->b1[j];
  for (int k = 0; k < d; ++k) {
    const float* w = m->W1uT + (size_t)k * H;
    const float uk = u[k];
    for (int j = 0; j < H; ++j) hu[j] = fmaf(w[j], uk, hu[j]);
  }
}

#define MLP_MB 6 /* candidates per register block */

/* C[mb][H] = relu(init[j] (+) sum_k A[c][k] * BT[k][j]);  A rows given by pointers */
static void mlp_layer_block(const float* const* arows, int mb, int K, const float* BT, int H,
                            const float* init, float* C /* [MLP_MB][H] */) {
#if ORC_AVX2
  for (int j0 = 0; j0 < H; j0 += 16) {
    __m256 acc[MLP_MB][2];
    const __m256 i0 = _mm256_loadu_ps(init + j0), i1 = _mm256_loadu_ps(init + j0 + 8);
    for (int c = 0; c < MLP_MB; ++c) { acc[c][0] = i0; acc[c][1] = i1; }
    for (int k = 0; k < K; ++k) {
      const __m256 b0 = _mm256_loadu_ps(BT + (size_t)k * H + j0);
      const __m256 b1 = _mm256_loadu_ps(BT + (size_t)k * H + j0 + 8);
      for (int c = 0; c < MLP_MB; ++c) {
        const __m256 a = _mm256_broadcast_ss(arows[c] + k);
        acc[c][0] = _mm256_fmadd_ps(a, b0, acc[c][0]);
        acc[c][1] = _mm256_fmadd_ps(a, b1, acc[c][1]);
      }
    }
    const __m256 z = _mm256_setzero_ps();
    for (int c = 0; c < mb; ++c) {
      /* relu: a>0?a:0  (max(a,0) differs only for NaN; and_ps with the compare mask is exact) */
      __m256 m0 = _mm256_cmp_ps(acc[c][0], z, _CMP_GT_OQ), m1 = _mm256_cmp_ps(acc[c][1], z, _CMP_GT_OQ);
      _mm256_storeu_ps(C + (size_t)c * H + j0, _mm256_and_ps(acc[c][0], m0));
      _mm256_storeu_ps(C + (size_t)c * H + j0 + 8, _mm256_and_ps(acc[c][1], m1));
    }
  }
#else
  for (int c = 0; c < mb; ++c)
    for (int j = 0; j < H; ++j) {
      float a = init[j];
      for (int k = 0; k < K; ++k) a = fmaf(arows[c][k], BT[(size_t)k * H + j], a);
      C[(size_t)c * H + j] = relu_f(a);
    }
#endif
}

static void mlp_score_hoisted(const orc_mlp_t* m, const float* hu, const float* table,
                              const int32_t* ids, int64_t n, float* out, float* scratch) {
  const int d = m->d, H = m->H;
  float* h1 = scratch;                       /* [MLP_MB][H] */
  float* h2 = scratch + (size_t)MLP_MB * H;  /* [MLP_MB][H] */
  const float* arows[MLP_MB];
  for (int64_t i0 = 0; i0 < n; i0 += MLP_MB) {
    int mb = (int)((n - i0) < MLP_MB ? (n - i0) : MLP_MB);
    for (int c = 0; c < MLP_MB; ++c) {
      int64_t i = i0 + (c < mb ? c : 0);
      arows[c] = ids ? table + (int64_t)ids[i] * d : table + i * d;
    }
    mlp_layer_block(arows, mb, d, m->W1xT, H, hu, h1);
    const float* hrows[MLP_MB];
    for (int c = 0; c < MLP_MB; ++c) hrows[c] = h1 + (size_t)(c < mb ? c : 0) * H;
    mlp_layer_block(hrows, mb, H, m->W2T, H, m->b2, h2);
    for (int c = 0; c < mb; ++c) {
      float s = 0.0f;
      const float* h = h2 + (size_t)c * H;
      for (int k = 0; k < H; ++k) s = fmaf(m->w3[k], h[k], s);
      out[i0 + c] = s;
    }
  }
}

void orc_mlp_score(const orc_mlp_t* m, const float* u, const float* table, const int32_t* ids,
                   int64_t n, float* out) {
  float* hu = (float*)aligned_alloc(64, sizeof(float) * m->H);
  float* scratch = (float*)aligned_alloc(64, sizeof(float) * 2 * MLP_MB * m->H);
  mlp_hoist(m, u, hu);
  mlp_score_hoisted(m, hu, table, ids, n, out, scratch);
  free(hu); free(scratch);
}

/* ------------------------------------------------------------------------------------------
 * nann_attention scorer (config 1).  model.py:189-233; model_util.py:70-97 (attention),
 * :32-67 (DNN), :9-11 (prelu).  fp32; dense = chain of fmaf starting from the bias, k ascending.
 * BN (tf.layers.batch_normalization, inference) arrives folded as y = x*scale + shift.
 * Blob order (floats): see oracle/README.md and nann_b200/scorer_weights.py.
 * ---------------------------------------------------------------------------------------- */
#define AT_L 50
#define AT_E 64
struct orc_attn {
  float* blob;
  const float *Wq1, *bq1, *aq, *Wq2, *bq2, *Wk1, *bk1, *ak, *Wk2, *bk2;
  const float *W1, *b1, *s1, *t1, *a1, *W2, *b2, *s2, *t2, *a2, *W3, *b3, *s3, *t3, *a3, *W4;
};

int64_t orc_attn_blob_size(void) {
  return 64 * 128 + 128 + 128 + 128 * 256 + 256 + 64 * 128 + 128 + 128 + 128 * 256 + 256 +
         128 * 128 + 4 * 128 + 128 * 64 + 4 * 64 + 64 * 32 + 4 * 32 + 32;
}

orc_attn_t* orc_attn_create(const float* blob, int64_t n_floats) {
  if (n_floats != orc_attn_blob_size()) return NULL;
  orc_attn_t* a = (orc_attn_t*)calloc(1, sizeof(*a));
  a->blob = dup_f32(blob, (size_t)n_floats);
  const float* p = a->blob;
#define TAKE(field, n) a->field = p; p += (n)
  TAKE(Wq1, 64 * 128); TAKE(bq1, 128); TAKE(aq, 128); TAKE(Wq2, 128 * 256); TAKE(bq2, 256);
  TAKE(Wk1, 64 * 128); TAKE(bk1, 128); TAKE(ak, 128); TAKE(Wk2, 128 * 256); TAKE(bk2, 256);
  TAKE(W1, 128 * 128); TAKE(b1, 128); TAKE(s1, 128); TAKE(t1, 128); TAKE(a1, 128);
  TAKE(W2, 128 * 64); TAKE(b2, 64); TAKE(s2, 64); TAKE(t2, 64); TAKE(a2, 64);
  TAKE(W3, 64 * 32); TAKE(b3, 32); TAKE(s3, 32); TAKE(t3, 32); TAKE(a3, 32);
  TAKE(W4, 32);
#undef TAKE
  return a;
}
void orc_attn_destroy(orc_attn_t* a) { if (a) { free(a->blob); free(a); } }

/* y[j] = b[j] (+) sum_k x[k] W[k][j]   (W in TF kernel layout [in][out]) */
static void dense_f(const float* x, int in, const float* W, const float* b, int out, float* y) {
  for (int j = 0; j < out; ++j) y[j] = b ? b[j] : 0.0f;
  for (int k = 0; k < in; ++k) {
    const float xk = x[k];
    const float* w = W + (size_t)k * out;
    for (int j = 0; j < out; ++j) y[j] = fmaf(xk, w[j], y[j]);
  }
}
static inline float prelu_f(float x, float alpha) { /* model_util.py:9-11 */
  float pos = x > 0.0f ? x : 0.0f, neg = x < 0.0f ? x : 0.0f;
  return fmaf(alpha, neg, pos);
}

void orc_attn_score(const orc_attn_t* a, const float* user, const float* table,
                    const int32_t* ids, int64_t n, float* out) {
  /* key side once per call: k' = dense_3(prelu_k(dense_2(u_l)))  model_util.py:84-85 */
  float kp[AT_L][256];
  for (int l = 0; l < AT_L; ++l) {
    float t[128];
    dense_f(user + l * AT_E, 64, a->Wk1, a->bk1, 128, t);
    for (int j = 0; j < 128; ++j) t[j] = prelu_f(t[j], a->ak[j]);
    dense_f(t, 128, a->Wk2, a->bk2, 256, kp[l]);
  }
  for (int64_t i = 0; i < n; ++i) {
    const float* x = ids ? table + (int64_t)ids[i] * AT_E : table + i * AT_E;
    float q[128], qp[256], lg[AT_L], h[128], y1[128], y2[64], y3[32];
    dense_f(x, 64, a->Wq1, a->bq1, 128, q);                       /* :81 */
    for (int j = 0; j < 128; ++j) q[j] = prelu_f(q[j], a->aq[j]);
    dense_f(q, 128, a->Wq2, a->bq2, 256, qp);                     /* :82 */
    float mx = -INFINITY;
    for (int l = 0; l < AT_L; ++l) {                              /* :90-91 einsum / sqrt(256) */
      float s = 0.0f;
      for (int dd = 0; dd < 256; ++dd) s = fmaf(qp[dd], kp[l][dd], s);
      lg[l] = s * 0.0625f;
      if (lg[l] > mx) mx = lg[l];
    }
    float Z = 0.0f;
    for (int l = 0; l < AT_L; ++l) { lg[l] = expf(lg[l] - mx); Z += lg[l]; } /* :93 softmax */
    for (int dd = 0; dd < AT_E; ++dd) h[dd] = 0.0f;
    for (int l = 0; l < AT_L; ++l) {                              /* :95 + model.py:208 reduce_sum */
      const float p = lg[l] / Z;
      for (int dd = 0; dd < AT_E; ++dd) h[dd] = fmaf(p, user[l * AT_E + dd], h[dd]);
    }
    for (int dd = 0; dd < AT_E; ++dd) h[AT_E + dd] = x[dd];       /* model.py:214 concat */
    dense_f(h, 128, a->W1, a->b1, 128, y1);
    for (int j = 0; j < 128; ++j) y1[j] = prelu_f(fmaf(y1[j], a->s1[j], a->t1[j]), a->a1[j]);
    dense_f(y1, 128, a->W2, a->b2, 64, y2);
    for (int j = 0; j < 64; ++j) y2[j] = prelu_f(fmaf(y2[j], a->s2[j], a->t2[j]), a->a2[j]);
    dense_f(y2, 64, a->W3, a->b3, 32, y3);
    for (int j = 0; j < 32; ++j) y3[j] = prelu_f(fmaf(y3[j], a->s3[j], a->t3[j]), a->a3[j]);
    float s = 0.0f;
    for (int k = 0; k < 32; ++k) s = fmaf(y3[k], a->W4[k], s);    /* 4_dnn, no bias :220 */
    out[i] = s;
  }
}

/* ------------------------------------------------------------------------------------------
 * exec.pb traversal -- build_opt_graph.py:109-149
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  int32_t* ids; float* sc; int64_t n, cap;
} vec_t;
static void vec_reserve(vec_t* v, int64_t cap) {
  if (cap <= v->cap) return;
  v->ids = (int32_t*)realloc(v->ids, (size_t)cap * sizeof(int32_t));
  v->sc = (float*)realloc(v->sc, (size_t)cap * sizeof(float));
  v->cap = cap;
}
static void vec_free(vec_t* v) { free(v->ids); free(v->sc); memset(v, 0, sizeof(*v)); }

/* py top_k (build_opt_graph.py:52-66): TopKV2 then gather ids by the returned indices */
static int topk_ids(const int32_t* ids, const float* sc, int64_t n, int k, int32_t* oid, float* osc,
                    int32_t* idx_tmp) {
  int st = orc_topk_v2_f32(sc, 1, n, k, osc, idx_tmp);
  if (st != ORC_OK) return st;
  for (int i = 0; i < k; ++i) oid[i] = ids[idx_tmp[i]];
  return ORC_OK;
}

/* ragged_gather (:39-49) with one group: fake_row_splits = [0, len] (:29-30) */
static int expand(const orc_index_t* ix, int level, const int32_t* ids, int64_t t, vec_t* out) {
  int64_t* iv = (int64_t*)malloc((size_t)(t > 0 ? t : 1) * sizeof(int64_t));
  for (int64_t i = 0; i < t; ++i) iv[i] = ids[i]; /* tf.cast(idx, int64) :46 */
  int64_t irs[2] = {0, t}, rrs[2], n_rrs = 0, n_ret = 0;
  int64_t n_pv = ix->nbr_row_splits[level][ix->n_items];
  int st = orc_group_gather_i32(ix->nbr_values[level], n_pv, ix->nbr_row_splits[level],
                                ix->n_items + 1, iv, t, irs, 2, 0, NULL, &n_ret, rrs, &n_rrs, NULL);
  if (st == ORC_OK) {
    vec_reserve(out, n_ret + 1);
    st = orc_group_gather_i32(ix->nbr_values[level], n_pv, ix->nbr_row_splits[level],
                              ix->n_items + 1, iv, t, irs, 2, 0, out->ids, &n_ret, rrs, &n_rrs, NULL);
    out->n = n_ret;
  }
  free(iv);
  return st;
}

/* set_difference (:33-36) in place on v */
static int diff_inplace(int32_t* ids, int64_t* n, int32_t* flags, int64_t n_flags) {
  int64_t rs[2] = {0, *n}, crs[2], n_crs = 0, n_c = 0;
  int st = orc_bitmap_ref_difference_i32(ids, *n, rs, 2, flags, n_flags, 1, ids, &n_c, crs, &n_crs, NULL);
  if (st == ORC_OK) *n = n_c;
  return st;
}

int orc_search(const orc_index_t* ix, orc_score_fn score, void* score_ctx,
               const int32_t* T, int64_t* out_ids, float* out_scores, int32_t* out_nodes,
               orc_search_stats_t* stats, int32_t** trace_ids, float** trace_scores,
               int64_t* trace_n, int64_t trace_cap) {
  int st = ORC_OK;
  const int64_t W = (ix->n_items + 31) / 32; /* bucket_size :115 */
  int32_t* flags = (int32_t*)malloc((size_t)(W > 0 ? W : 1) * sizeof(int32_t));
  vec_t R = {0}, Nx = {0}, cat = {0};
  int64_t maxk = 1;
  for (int i = 0; i < 6; ++i) if (T[i] > maxk) maxk = T[i];
  int32_t* idx_tmp = (int32_t*)malloc((size_t)maxk * sizeof(int32_t));
  if (stats) memset(stats, 0, sizeof(*stats));
  if (trace_n) for (int r = 0; r < 5; ++r) trace_n[r] = 0;

#define SCORE(round, v)                                                                    \
  do {                                                                                     \
    vec_reserve(&(v), (v).n + 1);                                                          \
    if ((v).n == 1) { st = ORC_INVALID_ARGUMENT; goto done; } /* tf.squeeze -> scalar :107 */ \
    score(score_ctx, (round), (v).ids, (v).n, (v).sc);                                     \
    if (stats) stats->n_scored[round] = (v).n;                                             \
    if (trace_n) {                                                                         \
      int64_t m = (v).n < trace_cap ? (v).n : trace_cap;                                   \
      trace_n[round] = (v).n;                                                              \
      if (trace_ids && trace_ids[round]) memcpy(trace_ids[round], (v).ids, (size_t)m * 4); \
      if (trace_scores && trace_scores[round]) memcpy(trace_scores[round], (v).sc, (size_t)m * 4); \
    }                                                                                      \
  } while (0)

  /* level 2 (:109-112): score every enter point, keep top T[0] */
  vec_reserve(&Nx, ix->n_ep + 1);
  memcpy(Nx.ids, ix->ep, (size_t)ix->n_ep * sizeof(int32_t));
  Nx.n = ix->n_ep;
  if (stats) stats->n_expanded[0] = Nx.n;
  SCORE(0, Nx);
  vec_reserve(&R, maxk * 8 + 8);
  if ((st = topk_ids(Nx.ids, Nx.sc, Nx.n, T[0], R.ids, R.sc, idx_tmp)) != ORC_OK) goto done;
  R.n = T[0];

  /* level 1 (:114-127) */
  if ((st = expand(ix, 1, R.ids, R.n, &Nx)) != ORC_OK) goto done;
  if (stats) stats->n_expanded[1] = Nx.n;
  memset(flags, 0, (size_t)W * sizeof(int32_t));                      /* Assign zeros :118 */
  if ((st = diff_inplace(R.ids, &R.n, flags, W)) != ORC_OK) goto done; /* :119-120 */
  if ((st = diff_inplace(Nx.ids, &Nx.n, flags, W)) != ORC_OK) goto done; /* :121-122 */
  SCORE(1, Nx);
  vec_reserve(&cat, R.n + Nx.n + 1);                                  /* concat :125-126 */
  memcpy(cat.ids, R.ids, (size_t)R.n * 4); memcpy(cat.ids + R.n, Nx.ids, (size_t)Nx.n * 4);
  memcpy(cat.sc, R.sc, (size_t)R.n * 4);   memcpy(cat.sc + R.n, Nx.sc, (size_t)Nx.n * 4);
  cat.n = R.n + Nx.n;
  vec_reserve(&R, (int64_t)T[1] + T[2] + T[3] + T[4] + 8);
  if ((st = topk_ids(cat.ids, cat.sc, cat.n, T[1], R.ids, R.sc, idx_tmp)) != ORC_OK) goto done;
  R.n = T[1];

  /* level 0 (:128-141) */
  {
    vec_t C = {0};
    vec_reserve(&C, maxk + 8);
    memcpy(C.ids, R.ids, (size_t)R.n * 4);
    C.n = R.n;
    memset(flags, 0, (size_t)W * sizeof(int32_t));                    /* Assign zeros :131 */
    st = diff_inplace(C.ids, &C.n, flags, W);                         /* :132-133 */
    for (int i = 0; i < 3 && st == ORC_OK; ++i) {
      if ((st = expand(ix, 0, C.ids, C.n, &Nx)) != ORC_OK) break;     /* :136 */
      if (stats) stats->n_expanded[2 + i] = Nx.n;
      if ((st = diff_inplace(Nx.ids, &Nx.n, flags, W)) != ORC_OK) break; /* :137 */
      vec_reserve(&Nx, Nx.n + 1);
      if (Nx.n == 1) { st = ORC_INVALID_ARGUMENT; break; }
      score(score_ctx, 2 + i, Nx.ids, Nx.n, Nx.sc);                   /* :138 */
      if (stats) stats->n_scored[2 + i] = Nx.n;
      if (trace_n) {
        int64_t m = Nx.n < trace_cap ? Nx.n : trace_cap;
        trace_n[2 + i] = Nx.n;
        if (trace_ids && trace_ids[2 + i]) memcpy(trace_ids[2 + i], Nx.ids, (size_t)m * 4);
        if (trace_scores && trace_scores[2 + i]) memcpy(trace_scores[2 + i], Nx.sc, (size_t)m * 4);
      }
      vec_reserve(&C, (int64_t)T[i + 2] + 8);
      if ((st = topk_ids(Nx.ids, Nx.sc, Nx.n, T[i + 2], C.ids, C.sc, idx_tmp)) != ORC_OK) break; /* :139 */
      C.n = T[i + 2];
      memcpy(R.ids + R.n, C.ids, (size_t)C.n * 4);                    /* concat :140-141 */
      memcpy(R.sc + R.n, C.sc, (size_t)C.n * 4);
      R.n += C.n;
    }
    vec_free(&C);
    if (st != ORC_OK) goto done;
  }

  /* final (:143-144) */
  vec_reserve(&cat, (int64_t)T[5] + 8);
  if ((st = topk_ids(R.ids, R.sc, R.n, T[5], cat.ids, cat.sc, idx_tmp)) != ORC_OK) goto done;
  for (int i = 0; i < T[5]; ++i) {
    if (out_nodes) out_nodes[i] = cat.ids[i];
    if (out_scores) out_scores[i] = cat.sc[i];
    if (out_ids) out_ids[i] = ix->item_ids[cat.ids[i]];
  }
done:
#undef SCORE
  vec_free(&R); vec_free(&Nx); vec_free(&cat);
  free(idx_tmp); free(flags);
  return st;
}

/* ------------------------------------------------------------------------------------------
 * `main.py --job-type test` traversal -- NANN_impls/nann/model/model.py:299-362.
 * Differs from exec.pb: candidates of a round are the UNIQUE unvisited neighbours in ascending id
 * order (tf.unique + tf.sets.set_difference, :319-322), the running result is merged with clamped k
 * (:268,:329-331), the next frontier is the new nodes that made the cut (:333-334), and the visited
 * set of a level starts as the level's entry results (:312).
 * ---------------------------------------------------------------------------------------- */
int orc_search_eval(const orc_index_t* ix, orc_score_fn score, void* score_ctx,
                    const int32_t* num_scoring_per_level, const int32_t* top_k_per_level, int topk_eval,
                    int64_t* out_ids, float* out_scores, int32_t* out_nodes, int32_t* out_n,
                    int64_t* n_scored_total) {
  int st = ORC_OK;
  const int start_level = 2;                               /* levels 1 and 0 carry neighbour lists */
  const int64_t W = (ix->n_items + 31) / 32;
  uint32_t* visited = (uint32_t*)malloc((size_t)(W > 0 ? W : 1) * sizeof(uint32_t));
  uint32_t* cand = (uint32_t*)calloc((size_t)(W > 0 ? W : 1), sizeof(uint32_t));
  vec_t R = {0}, Nx = {0}, cat = {0}, C = {0};
  int64_t scored = 0;
  int round = 0;
  if (out_n) *out_n = 0;
  if (num_scoring_per_level[start_level] != 1) { st = ORC_INVALID_ARGUMENT; goto done; }   /* assert :347 */

#define TOPK_CLAMPED(src, kk, dst)                                                              \
  do {                                                                                          \
    const int kc = (int)((src).n < (int64_t)(kk) ? (src).n : (int64_t)(kk));  /* reduce_min :268 */ \
    vec_reserve(&(dst), kc + 1);                                                                \
    int32_t* idx_tmp_ = (int32_t*)malloc((size_t)(kc > 0 ? kc : 1) * sizeof(int32_t));          \
    st = kc > 0 ? topk_ids((src).ids, (src).sc, (src).n, kc, (dst).ids, (dst).sc, idx_tmp_) : ORC_OK; \
    free(idx_tmp_);                                                                             \
    (dst).n = kc;                                                                               \
    if (st != ORC_OK) goto done;                                                                \
  } while (0)

  /* start level (:350-354): every enter point is scored */
  vec_reserve(&Nx, ix->n_ep + 1);
  memcpy(Nx.ids, ix->ep, (size_t)ix->n_ep * sizeof(int32_t));
  Nx.n = ix->n_ep;
  if (Nx.n == 1) { st = ORC_INVALID_ARGUMENT; goto done; }   /* tf.squeeze -> scalar, top_k needs rank >= 1 */
  if (Nx.n > 0) score(score_ctx, round, Nx.ids, Nx.n, Nx.sc);
  scored += Nx.n; ++round;
  TOPK_CLAMPED(Nx, top_k_per_level[start_level], R);

  for (int level = start_level - 1; level >= 0; --level) {   /* :356-357 */
    memset(visited, 0, (size_t)W * sizeof(uint32_t));         /* visited_idx = idx_ep :312 */
    for (int64_t i = 0; i < R.n; ++i) visited[R.ids[i] >> 5] |= 1u << (R.ids[i] & 31);
    vec_reserve(&C, R.n + 1);
    memcpy(C.ids, R.ids, (size_t)R.n * 4);
    C.n = R.n;
    for (int it = 0; it < num_scoring_per_level[level]; ++it) {   /* :317 */
      /* neighbours of the frontier, unique, minus visited, ascending (:319-322) */
      int64_t lo_w = W, hi_w = -1;
      for (int64_t i = 0; i < C.n; ++i) {
        const int64_t b = ix->nbr_row_splits[level][C.ids[i]], e = ix->nbr_row_splits[level][C.ids[i] + 1];
        for (int64_t j = b; j < e; ++j) {
          const int32_t v = ix->nbr_values[level][j];
          const int64_t w = v >> 5;
          if (!((visited[w] >> (v & 31)) & 1u)) {
            cand[w] |= 1u << (v & 31);
            if (w < lo_w) lo_w = w;
            if (w > hi_w) hi_w = w;
          }
        }
      }
      Nx.n = 0;
      for (int64_t w = lo_w; w <= hi_w; ++w) {
        uint32_t bits = cand[w];
        if (!bits) continue;
        visited[w] |= bits;                                    /* set_union :324 */
        cand[w] = 0;
        while (bits) {
          const int b = __builtin_ctz(bits);
          bits &= bits - 1;
          vec_reserve(&Nx, Nx.n + 64);
          Nx.ids[Nx.n++] = (int32_t)(w * 32 + b);
        }
      }
      if (Nx.n == 1) { st = ORC_INVALID_ARGUMENT; goto done; } /* scalar scores cannot be concatenated :329 */
      vec_reserve(&Nx, Nx.n + 1);
      if (Nx.n > 0) score(score_ctx, round, Nx.ids, Nx.n, Nx.sc);   /* :326 */
      scored += Nx.n; ++round;
      /* merged running result, clamped k (:329-331) */
      vec_reserve(&cat, R.n + Nx.n + 1);
      memcpy(cat.ids, R.ids, (size_t)R.n * 4); memcpy(cat.ids + R.n, Nx.ids, (size_t)Nx.n * 4);
      memcpy(cat.sc, R.sc, (size_t)R.n * 4);   memcpy(cat.sc + R.n, Nx.sc, (size_t)Nx.n * 4);
      cat.n = R.n + Nx.n;
      TOPK_CLAMPED(cat, top_k_per_level[level], R);
      /* frontier = new nodes whose score reaches the worst kept score (:333-334) */
      C.n = 0;
      vec_reserve(&C, Nx.n + 1);
      if (R.n > 0) {
        const float cut = R.sc[R.n - 1];
        for (int64_t i = 0; i < Nx.n; ++i)
          if (Nx.sc[i] >= cut) C.ids[C.n++] = Nx.ids[i];
      }
    }
  }
  {
    const int64_t n_out = R.n < topk_eval ? R.n : topk_eval;   /* results[:topk_eval] :359 */
    for (int64_t i = 0; i < n_out; ++i) {
      if (out_nodes) out_nodes[i] = R.ids[i];
      if (out_scores) out_scores[i] = R.sc[i];
      if (out_ids) out_ids[i] = ix->item_ids[R.ids[i]];
    }
    if (out_n) *out_n = (int32_t)n_out;
  }
done:
#undef TOPK_CLAMPED
  if (n_scored_total) *n_scored_total = scored;
  vec_free(&R); vec_free(&Nx); vec_free(&cat); vec_free(&C);
  free(visited); free(cand);
  return st;
}

/* ---- batch of queries with the mlp scorer, request-parallel ------------------------------ */
typedef struct {
  const orc_index_t* ix; const orc_mlp_t* m; float* hu; float* scratch;
} mlp_ctx_t;
static void mlp_score_cb(void* ctx, int round, const int32_t* ids, int64_t n, float* out) {
  (void)round;
  mlp_ctx_t* c = (mlp_ctx_t*)ctx;
  mlp_score_hoisted(c->m, c->hu, c->ix->emb, ids, n, out, c->scratch);
}

static double now_s(void) {
  struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

typedef struct {
  const orc_index_t* ix; const orc_mlp_t* m; const float* users; int64_t B; const int32_t* T;
  int64_t* out_ids; float* out_scores; int32_t* status;
  atomic_llong next; atomic_llong total;
} batch_job_t;

static void* batch_worker(void* arg) {
  batch_job_t* j = (batch_job_t*)arg;
  const int k = j->T[5];
  mlp_ctx_t c;
  c.ix = j->ix; c.m = j->m;
  c.hu = (float*)aligned_alloc(64, sizeof(float) * j->m->H);
  c.scratch = (float*)aligned_alloc(64, sizeof(float) * 2 * MLP_MB * j->m->H);
  long long local = 0;
  for (;;) {
    int64_t q = atomic_fetch_add(&j->next, 1); /* one request at a time per worker */
    if (q >= j->B) break;
    orc_search_stats_t s;
    mlp_hoist(j->m, j->users + q * j->ix->dim, c.hu);
    int st = orc_search(j->ix, mlp_score_cb, &c, j->T, j->out_ids ? j->out_ids + q * k : NULL,
                        j->out_scores ? j->out_scores + q * k : NULL, NULL, &s, NULL, NULL, NULL, 0);
    if (j->status) j->status[q] = st;
    for (int r = 0; r < 5; ++r) local += s.n_scored[r];
  }
  atomic_fetch_add(&j->total, local);
  free(c.hu); free(c.scratch);
  return NULL;
}

double orc_search_batch_mlp(const orc_index_t* ix, const orc_mlp_t* m, const float* users,
                            int64_t B, const int32_t* T, int nthreads, int64_t* out_ids,
                            float* out_scores, int32_t* status, int64_t* n_scored_total) {
  if (nthreads <= 0) nthreads = (int)sysconf(_SC_NPROCESSORS_ONLN);
  if (nthreads > B) nthreads = (int)(B > 0 ? B : 1);
  batch_job_t job = {ix, m, users, B, T, out_ids, out_scores, status, 0, 0};
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)nthreads);
  double t0 = now_s();
  for (int i = 1; i < nthreads; ++i) pthread_create(&th[i], NULL, batch_worker, &job);
  batch_worker(&job);
  for (int i = 1; i < nthreads; ++i) pthread_join(th[i], NULL);
  double dt = now_s() - t0;
  free(th);
  if (n_scored_total) *n_scored_total = (int64_t)atomic_load(&job.total);
  return dt;
}

/* ------------------------------------------------------------------------------------------
 * HugeConst loader checks -- huge_const_op.cc:85-182 (npy.h read_header/parse_header)
 * ---------------------------------------------------------------------------------------- */
static const char* npy_descr(int dtype) {
  switch (dtype) {
    case 0: return "<f2"; case 1: return "<f4"; case 2: return "<f8";
    case 3: return "<i4"; case 4: return "<i8"; default: return NULL;
  }
}
static int dtype_size(int dtype) { static const int s[5] = {2, 4, 8, 4, 8}; return s[dtype]; }

int orc_huge_const_load(const char* path, int dtype, const int64_t* shape, int rank, void* dst,
                        int64_t dst_bytes) {
  if (dtype < 0 || dtype > 4) return ORC_UNIMPLEMENTED; /* :143-146 */
  FILE* f = fopen(path, "rb");
  if (!f) return ORC_NOT_FOUND;                          /* :94-96 */
  unsigned char magic[10];
  int st = ORC_INTERNAL;
  char* hdr = NULL;
  if (fread(magic, 1, 8, f) != 8 || memcmp(magic, "\x93NUMPY", 6) != 0) goto out;
  size_t hlen;
  if (magic[6] == 1) {
    if (fread(magic + 8, 1, 2, f) != 2) goto out;
    hlen = magic[8] | ((size_t)magic[9] << 8);
  } else {
    unsigned char l4[4];
    if (fread(l4, 1, 4, f) != 4) goto out;
    hlen = l4[0] | ((size_t)l4[1] << 8) | ((size_t)l4[2] << 16) | ((size_t)l4[3] << 24);
  }
  hdr = (char*)malloc(hlen + 1);
  if (fread(hdr, 1, hlen, f) != hlen) goto out;
  hdr[hlen] = 0;
  {
    const char* p = strstr(hdr, "'fortran_order'");
    if (!p) goto out;
    p = strchr(p, ':');
    while (*++p == ' ') {}
    if (strncmp(p, "True", 4) == 0) { st = ORC_UNIMPLEMENTED; goto out; } /* :105-107 */
    /* shape: only the header's dims are compared (:110-115) */
    p = strstr(hdr, "'shape'");
    if (!p) goto out;
    p = strchr(p, '(');
    int i = 0;
    int64_t count = 1;
    ++p;
    while (*p && *p != ')') {
      while (*p == ' ' || *p == ',') ++p;
      if (*p == ')') break;
      char* e;
      long long v = strtoll(p, &e, 10);
      if (e == p) goto out;
      if (i >= rank || v != shape[i]) { st = ORC_INTERNAL; goto out; }
      count *= v; ++i; p = e;
    }
    p = strstr(hdr, "'descr'");
    if (!p) goto out;
    p = strchr(p, ':');
    p = strchr(p, '\'');
    const char* want = npy_descr(dtype);
    if (strncmp(p + 1, want, 3) != 0) { st = ORC_INTERNAL; goto out; }    /* :118-147 */
    if (dst) {
      int64_t bytes = count * dtype_size(dtype);
      if (bytes > dst_bytes) { st = ORC_INTERNAL; goto out; }
      if ((int64_t)fread(dst, 1, (size_t)bytes, f) != bytes) { st = ORC_INTERNAL; goto out; }
    }
    st = ORC_OK;
  }
out:
  free(hdr);
  fclose(f);
  return st;
}


/* ============================================================================================
 * Index construction.  NOT a restatement of faiss: the reference builds its graph with
 * faiss IndexHNSWFlat(d, 32) (NANN_impls/nann/delivery/build_hnsw_index.py:33-36), a third-party
 * library that is absent here and whose sequential-insertion graph no reference test pins.  This is
 * the CPU statement of the batch construction the CUDA builder implements (builder_kernels.cuh), so
 * that the GPU files can be compared bit for bit:
 *   d2(i,j)   = (sq_i + sq_j) - 2 * dot(i,j), dot and sq sequential fmaf chains over k = 0..d-1
 *   candidates of i = the n_cand members closest to i by (d2, id), i excluded
 *   forward links   = HNSW's diversity heuristic over the candidates in that order: keep j unless an
 *                     already kept m has d2(j,m) < d2(i,j); stop at M links
 *   links of i      = forward(i) U {j : i in forward(j)}, one entry per neighbour, the `cap` closest
 *                     by (d2, id), closest first.
 * parity unpinned by the reference (see header); pinned against the torch builder nann_b200/index.py by edge overlap.
 * ============================================================================================ */
static float dot_seq(const float* a, const float* b, int d) {
  float acc = 0.f;
  for (int k = 0; k < d; ++k) acc = fmaf(a[k], b[k], acc);
  return acc;
}
static uint32_t okey(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u << 1) == 0) u = 0;
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
static int cmp_u64(const void* a, const void* b) {
  uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b;
  return x < y ? -1 : (x > y ? 1 : 0);
}
typedef struct { const float* X; const float* sq; int64_t s; int d, n_cand, M; int32_t* fwd; float* fwd_d; int32_t* fwd_cnt; int64_t next; pthread_mutex_t mu; } build_job_t;

static void build_forward_row(build_job_t* J, int64_t i, uint64_t* keys, float* dist) {
  const int d = J->d;
  const float* xi = J->X + i * d;
  int64_t n = 0;
  for (int64_t j = 0; j < J->s; ++j) {
    if (j == i) continue;
    const float d2 = (J->sq[i] + J->sq[j]) - 2.0f * dot_seq(xi, J->X + j * d, d);
    keys[n++] = ((uint64_t)okey(d2) << 32) | (uint32_t)j;
    (void)dist;
  }
  qsort(keys, (size_t)n, sizeof(uint64_t), cmp_u64);
  const int c = (int)(n < J->n_cand ? n : J->n_cand);
  int nk = 0;
  int32_t kept[64];
  for (int a = 0; a < c && nk < J->M; ++a) {
    const int32_t j = (int32_t)(uint32_t)keys[a];
    const uint32_t ok = (uint32_t)(keys[a] >> 32);
    const uint32_t fb = (ok & 0x80000000u) ? (ok & 0x7fffffffu) : ~ok;
    float dj;
    memcpy(&dj, &fb, 4);
    int blocked = 0;
    for (int t = 0; t < nk && !blocked; ++t) {
      const int32_t m = kept[t];
      const float pd = (J->sq[j] + J->sq[m]) - 2.0f * dot_seq(J->X + (int64_t)j * d, J->X + (int64_t)m * d, d);
      if (pd < dj) blocked = 1;
    }
    if (!blocked) {
      kept[nk] = j;
      J->fwd[i * J->M + nk] = j;
      J->fwd_d[i * J->M + nk] = dj;
      ++nk;
    }
  }
  J->fwd_cnt[i] = nk;
}
static void* build_worker(void* arg) {
  build_job_t* J = (build_job_t*)arg;
  uint64_t* keys = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)(J->s > 0 ? J->s : 1));
  for (;;) {
    pthread_mutex_lock(&J->mu);
    const int64_t i = J->next++;
    pthread_mutex_unlock(&J->mu);
    if (i >= J->s) break;
    build_forward_row(J, i, keys, NULL);
  }
  free(keys);
  return NULL;
}

/* X [s][d] members of one level; out_links [s][cap] (member-local ids), out_cnt [s] */
int orc_build_level(const float* X, int64_t s, int d, int n_cand, int M, int cap, int nthreads,
                    int32_t* out_links, int32_t* out_cnt) {
  if (s < 0 || d <= 0 || M <= 0 || M > 64 || cap <= 0 || n_cand < 0) return ORC_INVALID_ARGUMENT;
  if (s == 0) return ORC_OK;
  float* sq = (float*)malloc(sizeof(float) * (size_t)s);
  for (int64_t i = 0; i < s; ++i) sq[i] = dot_seq(X + i * d, X + i * d, d);
  build_job_t J;
  J.X = X; J.sq = sq; J.s = s; J.d = d; J.n_cand = n_cand; J.M = M; J.next = 0;
  J.fwd = (int32_t*)malloc(sizeof(int32_t) * (size_t)(s * M));
  J.fwd_d = (float*)malloc(sizeof(float) * (size_t)(s * M));
  J.fwd_cnt = (int32_t*)calloc((size_t)s, sizeof(int32_t));
  pthread_mutex_init(&J.mu, NULL);
  if (nthreads < 1) nthreads = 1;
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)nthreads);
  for (int t = 1; t < nthreads; ++t) pthread_create(&th[t], NULL, build_worker, &J);
  build_worker(&J);
  for (int t = 1; t < nthreads; ++t) pthread_join(th[t], NULL);
  free(th);
  pthread_mutex_destroy(&J.mu);
  /* reverse links */
  int64_t* rev_off = (int64_t*)calloc((size_t)s + 1, sizeof(int64_t));
  for (int64_t i = 0; i < s; ++i)
    for (int e = 0; e < J.fwd_cnt[i]; ++e) rev_off[J.fwd[i * M + e] + 1]++;
  for (int64_t i = 0; i < s; ++i) rev_off[i + 1] += rev_off[i];
  uint64_t* rev = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)(rev_off[s] > 0 ? rev_off[s] : 1));
  int64_t* fill = (int64_t*)calloc((size_t)s, sizeof(int64_t));
  for (int64_t i = 0; i < s; ++i)
    for (int e = 0; e < J.fwd_cnt[i]; ++e) {
      const int32_t j = J.fwd[i * M + e];
      rev[rev_off[j] + fill[j]++] = ((uint64_t)okey(J.fwd_d[i * M + e]) << 32) | (uint32_t)i;
    }
  int64_t max_list = M;
  for (int64_t i = 0; i < s; ++i) if (M + rev_off[i + 1] - rev_off[i] > max_list) max_list = M + rev_off[i + 1] - rev_off[i];
  uint64_t* list = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)max_list);
  for (int64_t i = 0; i < s; ++i) {
    int64_t n = 0;
    for (int e = 0; e < J.fwd_cnt[i]; ++e) list[n++] = ((uint64_t)okey(J.fwd_d[i * M + e]) << 32) | (uint32_t)J.fwd[i * M + e];
    for (int64_t e = rev_off[i]; e < rev_off[i + 1]; ++e) list[n++] = rev[e];
    qsort(list, (size_t)n, sizeof(uint64_t), cmp_u64);
    int out = 0;
    for (int64_t e = 0; e < n && out < cap; ++e) {
      if (e > 0 && list[e] == list[e - 1]) continue;      /* forward and reverse entry of the same neighbour */
      out_links[i * cap + out++] = (int32_t)(uint32_t)list[e];
    }
    out_cnt[i] = out;
  }
  free(list); free(fill); free(rev); free(rev_off);
  free(J.fwd); free(J.fwd_d); free(J.fwd_cnt); free(sq);
  return ORC_OK;
}


/* ============================================================================================
 * BloomFilterDifference  (UO/bitmap_op/bitmap_ops.cc:264-432)
 * Fingerprint64 = farmhash::Fingerprint64 (tensorflow/core/platform/fingerprint.h:80-90), a third-party
 * dependency that is not vendored in the reference tree (tensorflow/workspace.bzl:250-257 pins
 * google/farmhash @ 816a4ae622e964763ca0862d9dbd19324a1eaf45).  Restated below from the published
 * algorithm (farmhashna::Hash64, inputs up to 32 bytes -- decimal strings of int64 have at most 20) and
 * pinned against the vectors the reference's own tests hold:
 *   tensorflow/python/kernel_tests/string_to_hash_bucket_op_test.py:46-49  ('a','b','c','d')
 *   tensorflow/core/platform/fingerprint_test.cc:27-28                     ("Hello","World")
 * (both exercise the 1..3 and 4..7 byte branches; the 8..16 and 17..32 byte branches follow the published
 * source with no reference-held vector).
 * ============================================================================================ */
#define FH_K0 0xc3a5c85c97cb3127ULL
#define FH_K1 0xb492b66fbe98f273ULL
#define FH_K2 0x9ae16a3b2f90404fULL
static uint64_t fh_fetch64(const char* p) { uint64_t v; memcpy(&v, p, 8); return v; }
static uint32_t fh_fetch32(const char* p) { uint32_t v; memcpy(&v, p, 4); return v; }
static uint64_t fh_rot(uint64_t v, int s) { return s == 0 ? v : ((v >> s) | (v << (64 - s))); }
static uint64_t fh_shift_mix(uint64_t v) { return v ^ (v >> 47); }
static uint64_t fh_len16(uint64_t u, uint64_t v, uint64_t mul) {
  uint64_t a = (u ^ v) * mul;
  a ^= (a >> 47);
  uint64_t b = (v ^ a) * mul;
  b ^= (b >> 47);
  return b * mul;
}
uint64_t orc_fingerprint64(const char* s, int64_t len) {
  if (len <= 16) {
    if (len >= 8) {
      const uint64_t mul = FH_K2 + (uint64_t)len * 2, a = fh_fetch64(s) + FH_K2, b = fh_fetch64(s + len - 8);
      const uint64_t c = fh_rot(b, 37) * mul + a, d = (fh_rot(a, 25) + b) * mul;
      return fh_len16(c, d, mul);
    }
    if (len >= 4) {
      const uint64_t mul = FH_K2 + (uint64_t)len * 2, a = fh_fetch32(s);
      return fh_len16((uint64_t)len + (a << 3), fh_fetch32(s + len - 4), mul);
    }
    if (len > 0) {
      const uint8_t a = (uint8_t)s[0], b = (uint8_t)s[len >> 1], c = (uint8_t)s[len - 1];
      const uint32_t y = (uint32_t)a + ((uint32_t)b << 8), z = (uint32_t)len + ((uint32_t)c << 2);
      return fh_shift_mix(y * FH_K2 ^ z * FH_K0) * FH_K2;
    }
    return FH_K2;
  }
  if (len <= 32) {
    const uint64_t mul = FH_K2 + (uint64_t)len * 2, a = fh_fetch64(s) * FH_K1, b = fh_fetch64(s + 8);
    const uint64_t c = fh_fetch64(s + len - 8) * mul, d = fh_fetch64(s + len - 16) * FH_K2;
    return fh_len16(fh_rot(a + b, 43) + fh_rot(c, 30) + d, a + fh_rot(b + FH_K2, 18) + c, mul);
  }
  return 0;   /* longer inputs never occur for decimal integers */
}

static int bloom_is_prime(int64_t x) {                       /* bitmap_ops.cc:395-402 ("not fit when x equal 1") */
  for (int64_t i = (int64_t)(sqrt((double)x) + 1e-6); i > 1; i--)
    if ((x % i) == 0) return 0;
  return 1;
}
void orc_bloom_primes(int64_t bucket_size, int64_t primes[4]) {   /* :404-421 */
  static const int mod_param[4] = {29, 47, 67, 83};
  for (int i = 0; i < 4; i++) {
    int64_t target = (int64_t)mod_param[i] * bucket_size * 32, p = 0;
    for (int64_t n = target; n > 0; n--) if (bloom_is_prime(n)) { p = n; break; }
    primes[i] = p;
  }
}
/* the Differ loop (:334-359) for one node; returns `miss` */
static int bloom_touch(int64_t node, int64_t bucket, int64_t bucket_size, const int64_t primes[4], int32_t* flags) {
  static const int mult[4] = {1, 3, 5, 7};
  char buf[32];
  const int len = snprintf(buf, sizeof(buf), "%lld", (long long)node);      /* std::to_string(node) */
  uint64_t raw = orc_fingerprint64(buf, len);
  if (bucket > 0) raw = raw % (uint64_t)bucket;
  int miss = 0;
  for (int l = 0; l < 4; l++) {
    const uint64_t lp = (uint64_t)primes[l];
    const uint64_t tmp = ((raw * (uint64_t)mult[l]) % lp + lp) % lp;
    const int64_t bucket_id = (int64_t)(tmp % (uint64_t)(bucket_size * 32));
    const int fi = (int)(bucket_id >> 5), bi = (int)(bucket_id & 31);
    if (!((uint32_t)flags[fi] & (1u << bi))) { miss++; flags[fi] = (int32_t)((uint32_t)flags[fi] | (1u << bi)); }
  }
  return miss;
}
#define ORC_BLOOM(T, SFX)                                                                                          \
  int orc_bloom_filter_difference_##SFX(const T* v, int64_t n_v, const int64_t* rs, int64_t n_rs, int32_t* flags,  \
                                        int64_t n_flags, int64_t bucket, int64_t bucket_size, T* c_values,          \
                                        int64_t* c_row_splits, int64_t* n_c, int* code) {                           \
    *code = orc_validate_ragged(n_v, rs, n_rs);                                                                     \
    if (*code) return ORC_INVALID_ARGUMENT;                                                                         \
    if (bucket_size < 1 || bucket < 0 || n_flags < bucket_size) return ORC_INVALID_ARGUMENT;                        \
    *n_c = 0;                                                                                                       \
    if (n_rs == 1) { c_row_splits[0] = 0; return ORC_OK; }                       /* void input, :312-322 */         \
    int64_t primes[4];                                                                                              \
    orc_bloom_primes(bucket_size, primes);                                                                          \
    c_row_splits[0] = 0;                                                                                            \
    for (int64_t g = 0; g + 1 < n_rs; ++g) {                                                                        \
      for (int64_t j = rs[g]; j < rs[g + 1]; ++j)                                                                   \
        if (bloom_touch((int64_t)v[j], bucket, bucket_size, primes, flags) > 0) c_values[(*n_c)++] = v[j];           \
      c_row_splits[g + 1] = *n_c;                                                                                   \
    }                                                                                                               \
    return ORC_OK;                                                                                                  \
  }
ORC_BLOOM(int32_t, i32)
ORC_BLOOM(int64_t, i64)
