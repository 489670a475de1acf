// scorer_attn.cuh -- the reference's own scorer (config 1) fused with the row gather:
// Model.forward, NANN_impls/nann/model/model.py:189-233; nonlinear_attention
// model_util.py:70-97; DNN :32-67; prelu :9-11.  fp32 FFMA; every dense output is a sequential
// fmaf chain starting from the bias (as the oracle defines it); BN arrives folded (scale, shift).
// One CTA = 32 candidates of one query; all intermediates stay in shared memory.
#pragma once

namespace nann {

constexpr int ATT_L = 50;    // user sequence length  (build_opt_graph.py:26)
constexpr int ATT_E = 64;    // item / sequence embedding dim (:25)
constexpr int ATT_QK = 256;  // attention projection dim (4*emb_dim, model_util.py:82)
constexpr int ATT_NC = 32;   // candidates per CTA
constexpr int ATT_THREADS = 256;

struct AttnW {  // device pointers into the blob, TF kernel layout [in][out]
  const float *Wq1, *bq1, *aq, *Wq2, *bq2, *Wk1, *bk1, *ak, *Wk2, *bk2;
  const float *W1, *b1, *s1, *t1, *a1, *W2, *b2, *s2, *t2, *a2, *W3, *b3, *s3, *t3, *a3, *W4;
};
constexpr int64_t ATT_BLOB = 64 * 128 + 128 + 128 + 128 * 256 + 256 + 64 * 128 + 128 + 128 + 128 * 256 + 256 +
                             128 * 128 + 4 * 128 + 128 * 64 + 4 * 64 + 64 * 32 + 4 * 32 + 32;

static AttnW attn_weights(const float* p) {
  AttnW w;
  auto take = [&](int64_t n) { const float* r = p; p += n; return r; };
  w.Wq1 = take(64 * 128); w.bq1 = take(128); w.aq = take(128); w.Wq2 = take(128 * 256); w.bq2 = take(256);
  w.Wk1 = take(64 * 128); w.bk1 = take(128); w.ak = take(128); w.Wk2 = take(128 * 256); w.bk2 = take(256);
  w.W1 = take(128 * 128); w.b1 = take(128); w.s1 = take(128); w.t1 = take(128); w.a1 = take(128);
  w.W2 = take(128 * 64); w.b2 = take(64); w.s2 = take(64); w.t2 = take(64); w.a2 = take(64);
  w.W3 = take(64 * 32); w.b3 = take(32); w.s3 = take(32); w.t3 = take(32); w.a3 = take(32);
  w.W4 = take(32);
  return w;
}

__device__ __forceinline__ float prelu_dev(float x, float alpha) {
  const float pos = x > 0.f ? x : 0.f, neg = x < 0.f ? x : 0.f;
  return fmaf(alpha, neg, pos);
}

// key side, once per query: kp[q][l][:] = dense_3(prelu_k(dense_2(u_l)))   (model_util.py:84-85)
__global__ void __launch_bounds__(ATT_QK)
attn_keys_kernel(const float* __restrict__ users, AttnW w, float* __restrict__ kp) {
  __shared__ float su[ATT_E], st[128];
  const int l = blockIdx.x, q = blockIdx.y, j = threadIdx.x;
  if (j < ATT_E) su[j] = users[((int64_t)q * ATT_L + l) * ATT_E + j];
  __syncthreads();
  if (j < 128) {
    float a = w.bk1[j];
    for (int k = 0; k < ATT_E; ++k) a = fmaf(su[k], w.Wk1[k * 128 + j], a);
    st[j] = prelu_dev(a, w.ak[j]);
  }
  __syncthreads();
  float a = w.bk2[j];
  for (int k = 0; k < 128; ++k) a = fmaf(st[k], w.Wk2[k * ATT_QK + j], a);
  kp[((int64_t)q * ATT_L + l) * ATT_QK + j] = a;
}

// out[c][j] = post(b[j] (+) sum_k in[c][k] * W[k][j]) for the CTA's ATT_NC candidates.
// Thread owns column j = tid % OUT and CPT = NC*OUT/256 candidates.
template <int IN, int OUT, int IN_PITCH, int OUT_PITCH, int MODE /*0 none, 1 prelu, 2 bn+prelu*/>
__device__ __forceinline__ void dense_tile(const float* __restrict__ in, const float* __restrict__ W,
                                           const float* __restrict__ b, const float* __restrict__ sc,
                                           const float* __restrict__ sh, const float* __restrict__ al,
                                           float* __restrict__ out, int tid) {
  constexpr int COLS = OUT < ATT_THREADS ? OUT : ATT_THREADS;
  constexpr int GROUPS = ATT_THREADS / COLS;
  constexpr int CPT = ATT_NC / GROUPS;
  static_assert(OUT <= ATT_THREADS && ATT_NC % GROUPS == 0, "tile shape");
  const int j = tid % COLS, c0 = (tid / COLS) * CPT;
  float acc[CPT];
  const float bj = b ? b[j] : 0.f;
#pragma unroll
  for (int c = 0; c < CPT; ++c) acc[c] = bj;
#pragma unroll 4
  for (int k = 0; k < IN; ++k) {
    const float wv = W[k * OUT + j];
#pragma unroll
    for (int c = 0; c < CPT; ++c) acc[c] = fmaf(in[(c0 + c) * IN_PITCH + k], wv, acc[c]);
  }
#pragma unroll
  for (int c = 0; c < CPT; ++c) {
    float v = acc[c];
    if (MODE == 1) v = prelu_dev(v, al[j]);
    if (MODE == 2) v = prelu_dev(fmaf(v, sc[j], sh[j]), al[j]);
    out[(c0 + c) * OUT_PITCH + j] = v;
  }
}

struct AttnArgs {
  const float* table; const int32_t* ids; int64_t ids_stride; int64_t rows_stride;
  const int32_t* n_ptr; int n_fixed;
  const float* users;   // [B][50][64]
  const float* kp;      // [B][50][256]
  AttnW w;
  float* out; int64_t out_stride; const int32_t* status;
};

constexpr int ATT_KP_PITCH = ATT_QK + 1;
// smem floats: x 32*64, q 32*128 (reused: y1), qp 32*256 (reused: y2,y3), kp 50*257, u 50*64, p 32*52, h 32*128
constexpr int ATT_SMEM_FLOATS = ATT_NC * 64 + ATT_NC * 128 + ATT_NC * 256 + ATT_L * ATT_KP_PITCH + ATT_L * ATT_E +
                                ATT_NC * 52 + ATT_NC * 128;
constexpr int ATT_SMEM_BYTES = ATT_SMEM_FLOATS * 4;

__global__ void __launch_bounds__(ATT_THREADS)
attn_score_kernel(AttnArgs p) {
  extern __shared__ __align__(16) float sm[];
  float* xs = sm;                          // [32][64]
  float* qs = xs + ATT_NC * 64;            // [32][128]
  float* qp = qs + ATT_NC * 128;           // [32][256]
  float* kps = qp + ATT_NC * 256;          // [50][257]
  float* us = kps + ATT_L * ATT_KP_PITCH;  // [50][64]
  float* ps = us + ATT_L * ATT_E;          // [32][52]
  float* hs = ps + ATT_NC * 52;            // [32][128]
  const int q = blockIdx.y, tid = threadIdx.x;
  if (p.status && p.status[q] != 0) return;
  const int n = p.n_ptr ? p.n_ptr[q] : p.n_fixed;
  const int t0 = blockIdx.x * ATT_NC;
  if (t0 >= n) return;
  const int nt = min(ATT_NC, n - t0);
  // stage: candidate rows (16 lanes x float4 = one 256-B row), user sequence, key projections
  for (int v = tid; v < ATT_NC * 16; v += ATT_THREADS) {
    const int c = v >> 4, part = v & 15, cc = c < nt ? c : 0;
    const float* src = p.ids ? p.table + (int64_t)p.ids[(int64_t)q * p.ids_stride + t0 + cc] * ATT_E
                             : p.table + ((int64_t)q * p.rows_stride + t0 + cc) * ATT_E;
    *reinterpret_cast<float4*>(xs + c * 64 + part * 4) = *reinterpret_cast<const float4*>(src + part * 4);
  }
  for (int v = tid; v < ATT_L * ATT_E; v += ATT_THREADS) us[v] = p.users[(int64_t)q * ATT_L * ATT_E + v];
  for (int v = tid; v < ATT_L * ATT_QK; v += ATT_THREADS)
    kps[(v / ATT_QK) * ATT_KP_PITCH + (v % ATT_QK)] = p.kp[(int64_t)q * ATT_L * ATT_QK + v];
  __syncthreads();
  // q = prelu_q(dense(x)); q' = dense_1(q)                                    model_util.py:81-82
  dense_tile<64, 128, 64, 128, 1>(xs, p.w.Wq1, p.w.bq1, nullptr, nullptr, p.w.aq, qs, tid);
  __syncthreads();
  dense_tile<128, 256, 128, 256, 0>(qs, p.w.Wq2, p.w.bq2, nullptr, nullptr, nullptr, qp, tid);
  __syncthreads();
  // att logits = q'.k'_l / sqrt(256)                                           :90-91
  for (int v = tid; v < ATT_NC * ATT_L; v += ATT_THREADS) {
    const int c = v / ATT_L, l = v % ATT_L;
    float s = 0.f;
    const float* a = qp + c * 256;
    const float* b = kps + l * ATT_KP_PITCH;
#pragma unroll 8
    for (int d = 0; d < ATT_QK; ++d) s = fmaf(a[d], b[d], s);
    ps[c * 52 + l] = s * 0.0625f;
  }
  __syncthreads();
  if (tid < ATT_NC) {  // softmax over the 50 positions (no mask, as the reference)   :93
    float* row = ps + tid * 52;
    float mx = -INFINITY;
    for (int l = 0; l < ATT_L; ++l) mx = row[l] > mx ? row[l] : mx;
    float Z = 0.f;
    for (int l = 0; l < ATT_L; ++l) { row[l] = expf(row[l] - mx); Z += row[l]; }
    for (int l = 0; l < ATT_L; ++l) row[l] = row[l] / Z;
  }
  __syncthreads();
  // h = [sum_l p_l u_l ; x]                                                    :95, model.py:208,214
  for (int v = tid; v < ATT_NC * ATT_E; v += ATT_THREADS) {
    const int c = v >> 6, d = v & 63;
    float a = 0.f;
    for (int l = 0; l < ATT_L; ++l) a = fmaf(ps[c * 52 + l], us[l * ATT_E + d], a);
    hs[c * 128 + d] = a;
    hs[c * 128 + 64 + d] = xs[c * 64 + d];
  }
  __syncthreads();
  float* y1 = qs;        // [32][128]
  float* y2 = qp;        // [32][64]
  float* y3 = qp + ATT_NC * 64;  // [32][32]
  dense_tile<128, 128, 128, 128, 2>(hs, p.w.W1, p.w.b1, p.w.s1, p.w.t1, p.w.a1, y1, tid);
  __syncthreads();
  dense_tile<128, 64, 128, 64, 2>(y1, p.w.W2, p.w.b2, p.w.s2, p.w.t2, p.w.a2, y2, tid);
  __syncthreads();
  dense_tile<64, 32, 64, 32, 2>(y2, p.w.W3, p.w.b3, p.w.s3, p.w.t3, p.w.a3, y3, tid);
  __syncthreads();
  if (tid < nt) {  // 4_dnn: Dense(1), no bias                                    model.py:220
    float s = 0.f;
    for (int k = 0; k < 32; ++k) s = fmaf(y3[tid * 32 + k], p.w.W4[k], s);
    p.out[(int64_t)q * p.out_stride + t0 + tid] = s;
  }
}

static nann_status attn_prepare_users(nann_scorer* s, const float* users_dev, int B, float* kp, cudaStream_t st) {
  dim3 grid(ATT_L, (unsigned)B);
  NANN_LAUNCH(attn_keys_kernel, grid, ATT_QK, 0, st, users_dev, attn_weights(s->blob), kp);
  return NANN_OK;
}

static nann_status attn_score(nann_scorer* s, const ScoreCall& c, cudaStream_t st) {
  AttnArgs a{};
  a.table = c.table; a.ids = c.ids; a.ids_stride = c.ids_stride; a.rows_stride = c.rows_stride;
  a.n_ptr = c.n_ptr; a.n_fixed = c.n_fixed; a.users = c.users; a.kp = c.hu; a.w = attn_weights(s->blob);
  a.out = c.out; a.out_stride = c.out_stride; a.status = c.status;
  static bool attr_set[64] = {false};
  if (!attr_set[s->device & 63]) {
    NANN_CUDA(cudaFuncSetAttribute(attn_score_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM_BYTES));
    attr_set[s->device & 63] = true;
  }
  dim3 grid((unsigned)ceil_div(c.max_n, ATT_NC), (unsigned)c.B);
  NANN_LAUNCH(attn_score_kernel, grid, ATT_THREADS, ATT_SMEM_BYTES, st, a);
  return NANN_OK;
}

}  // namespace nann

extern "C" {
int64_t nann_scorer_attention_blob_size(void) { return nann::ATT_BLOB; }

nann_status nann_scorer_create_attention(const float* blob, int64_t n_floats, int device, nann_scorer_t** out) {
  if (!out) return nann::fail(NANN_INVALID_ARGUMENT, "null out");
  *out = nullptr;
  NANN_TRY(nann::require_device());
  if (!blob || n_floats != nann::ATT_BLOB)
    return nann::fail(NANN_INVALID_ARGUMENT, "attention blob must hold %lld floats (got %lld)", (long long)nann::ATT_BLOB,
                      (long long)n_floats);
  NANN_CUDA(cudaSetDevice(device));
  auto* s = new nann_scorer();
  s->kind = 1; s->device = device; s->d = nann::ATT_E; s->H = 0;
  nann_status rc = nann::to_device_copy(blob, n_floats, &s->blob, (cudaStream_t)0);
  if (rc == NANN_OK && cudaStreamSynchronize(0) != cudaSuccess) rc = nann::fail(NANN_INTERNAL, "blob upload failed");
  if (rc != NANN_OK) { nann_scorer_destroy(s); return rc; }
  *out = s;
  return NANN_OK;
}
}
