// lib_search.inl -- scorer handles, the HBM-resident index, and the batched exec.pb dataflow
// (NANN_impls/nann/delivery/build_opt_graph.py:109-149) as a fixed sequence of kernel launches.

// BlazeXlaOp's admission control (UO/blaze_op/blaze_xla_kernel.cc:87-101,221-258): at most `running_max` runs at a
// time (BLAZE_THREADS_NUM, default 2); a request that finds the op busy waits -- up to `wait_ms` when that is set in
// the blaze options (then "blaze wait too long": Internal while still queued, DeadlineExceeded when it would start
// late), otherwise as long as fewer than `max_waiting` requests are queued (DENSE_MAX_WAITING_COUNT, default 10; else
// Internal "waiting pool is full").  The reference re-posts a waiting request to its thread pool in a loop; here the
// caller's thread waits on a condition variable -- same admissions, same errors, no spinning.
struct BlazeAdmission {
  int running_max = 2, max_waiting = 10;
  long long wait_ns = 0;
  int running = 0, waiting = 0;
  std::mutex mu;
  std::condition_variable cv;
  static long long env_or(const char* name, long long dflt) {
    const char* e = std::getenv(name);
    return (e && *e) ? atoll(e) : dflt;
  }
  BlazeAdmission() {
    running_max = (int)env_or("BLAZE_THREADS_NUM", 2);           // BlazeThreadsCount(), :87-93
    max_waiting = (int)env_or("DENSE_MAX_WAITING_COUNT", 10);    // BlazeWatingCount(), :95-101
  }
  static long long now_ns() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (long long)ts.tv_sec * 1000000000ll + ts.tv_nsec;
  }
  nann_status enter() {                                           // Schedule(), :221-258
    const long long begin = now_ns();
    std::unique_lock<std::mutex> lk(mu);
    bool first = true;
    for (;;) {
      const long long waited = now_ns() - begin;
      if (running >= running_max) {
        if (wait_ns > 0) {
          if (waited > wait_ns) {
            if (!first) --waiting;
            return nann::fail(NANN_INTERNAL, "blaze wait too long %lld", waited);
          }
        } else if (waiting >= max_waiting && first) {
          return nann::fail(NANN_INTERNAL, "waiting pool is full %d", waiting);
        }
        if (first) { ++waiting; first = false; }
        if (wait_ns > 0) cv.wait_for(lk, std::chrono::nanoseconds(std::max<long long>(wait_ns - waited, 1000)));
        else cv.wait(lk);
        continue;
      }
      if (!first) --waiting;
      if (wait_ns > 0 && waited > wait_ns) {                      // would start too late (:243-248)
        cv.notify_one();
        return nann::fail(NANN_DEADLINE_EXCEEDED, "blaze wait too long %lld", waited);
      }
      ++running;
      return NANN_OK;
    }
  }
  void leave() {
    { std::lock_guard<std::mutex> lk(mu); --running; }
    cv.notify_one();
  }
};

struct nann_scorer {
  BlazeAdmission gate;
  int kind = 0;  // 0 = mlp, 1 = attention
  int device = 0;
  int precision = NANN_SCORER_EXACT;
  int d = 0, H = 0;
  // mlp (device): transposed so that a k-row of weights is contiguous over neurons
  float *W1uT = nullptr, *W1xT = nullptr, *W2T = nullptr, *b1 = nullptr, *b2 = nullptr, *w3 = nullptr;
  // attention (device blob + sub-pointers)
  float* blob = nullptr;
  // tensor-core operand planes (filled lazily by the TENSOR path)
  void* tc = nullptr;
};

struct nann_index {
  int device = 0;
  int64_t n_items = 0;
  int dim = 0;
  float* emb = nullptr;        // [n_local][dim] f32: rows row_lo .. row_lo + n_local of the table (all of it unless sharded)
  int64_t row_lo = 0, n_local = 0;
  int64_t* item_ids = nullptr; // [n_items]
  int32_t* ep = nullptr;       // [n_ep]
  int64_t n_ep = 0;
  int32_t* nbr_values[2] = {nullptr, nullptr};
  int64_t* nbr_rs[2] = {nullptr, nullptr};
  int64_t n_nbr[2] = {0, 0};
  int max_deg[2] = {0, 0};
};

namespace nann {

// ---- small conversion / validation kernels ------------------------------------------------
__global__ void f16_to_f32_kernel(const __half* __restrict__ in, float* __restrict__ out, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = __half2float(in[i]);
}
__global__ void i64_to_i32_kernel(const int64_t* __restrict__ in, int32_t* __restrict__ out, int64_t n,
                                  int64_t limit, int* __restrict__ bad) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t v = in[i];
    if (v < 0 || v >= limit) *bad = 1;
    out[i] = (int32_t)v;
  }
}
__global__ void check_i32_range_kernel(const int32_t* __restrict__ in, int64_t n, int64_t limit,
                                       int* __restrict__ bad) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    if (in[i] < 0 || in[i] >= limit) *bad = 1;
}
// CSR sanity + max degree; strictly ascending check for enter points
__global__ void csr_check_kernel(const int64_t* __restrict__ rs, int64_t n_rows, int64_t n_vals,
                                 int* __restrict__ max_deg, int* __restrict__ bad) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_rows; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t d = rs[i + 1] - rs[i];
    if (d < 0 || d > (1 << 20)) *bad = 1;
    else atomicMax(max_deg, (int)d);
    if (i == 0 && rs[0] != 0) *bad = 2;
    if (i == n_rows - 1 && rs[n_rows] != n_vals) *bad = 3;
  }
}
__global__ void ascending_check_kernel(const int32_t* __restrict__ v, int64_t n, int* __restrict__ bad) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i + 1 < n; i += (int64_t)gridDim.x * blockDim.x)
    if (v[i] >= v[i + 1]) *bad = 1;
}
__global__ void transpose_kernel(const float* __restrict__ in, int rows, int cols, int ld_in, int col0,
                                 float* __restrict__ out /* [cols][rows] */) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= (int64_t)rows * cols) return;
  const int r = (int)(t / cols), c = (int)(t % cols);
  out[(int64_t)c * rows + r] = in[(int64_t)r * ld_in + col0 + c];
}

template <typename T>
static nann_status to_device_copy(const T* src, int64_t n, T** out, cudaStream_t st) {
  *out = nullptr;
  NANN_CUDA(cudaMalloc(out, (size_t)(n > 0 ? n : 1) * sizeof(T)));
  if (n > 0)
    NANN_CUDA(cudaMemcpyAsync(*out, src, (size_t)n * sizeof(T),
                              is_device_ptr(src) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
  return NANN_OK;
}

// ids (i32 or i64, host or device) -> owned device i32, range-checked against [0, limit)
static nann_status ids_to_device_i32(const void* src, int dtype, int64_t n, int64_t limit, int32_t** out,
                                     int* d_bad, cudaStream_t st) {
  *out = nullptr;
  NANN_CUDA(cudaMalloc(out, (size_t)(n > 0 ? n : 1) * sizeof(int32_t)));
  if (n == 0) return NANN_OK;
  const int blocks = (int)std::min<int64_t>(ceil_div(n, 256), 148 * 8);
  if (dtype == NANN_I32) {
    NANN_CUDA(cudaMemcpyAsync(*out, src, (size_t)n * 4,
                              is_device_ptr(src) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
    NANN_LAUNCH(check_i32_range_kernel, blocks, 256, 0, st, *out, n, limit, d_bad);
  } else if (dtype == NANN_I64) {
    DevIn<int64_t> tmp;
    NANN_TRY(tmp.init((const int64_t*)src, n, st));
    NANN_LAUNCH(i64_to_i32_kernel, blocks, 256, 0, st, tmp.d, *out, n, limit, d_bad);
    NANN_CUDA(cudaStreamSynchronize(st));
  } else {
    return fail(NANN_INVALID_ARGUMENT, "id arrays must be int32 or int64");
  }
  return NANN_OK;
}

// ---- scorer dispatch ------------------------------------------------------------------------
struct TcWorkspace;
struct ScoreCall {
  TcWorkspace* ws = nullptr;   // tensor-core path scratch, owned by the caller (one per stream)
  const float* table; const int32_t* ids; int64_t ids_stride; int64_t rows_stride;
  const int32_t* n_ptr; int n_fixed; int max_n; int B;
  const float* hu;      // mlp: hoisted layer-1 prefix [B][H];  attention: key side [B][50][256]
  const float* users;   // [B][user_floats]
  float* out; int64_t out_stride; const int32_t* status;
};

static nann_status scorer_prepare_users(nann_scorer* s, const float* users_dev, int B, float* hu,
                                        cudaStream_t st);
static nann_status scorer_score(nann_scorer* s, const ScoreCall& c, cudaStream_t st);

}  // namespace nann

#include "scorer_attn.cuh"
#include "scorer_mlp_tc.cuh"

namespace nann {

static nann_status scorer_prepare_users(nann_scorer* s, const float* users_dev, int B, float* hu,
                                        cudaStream_t st) {
  if (B <= 0) return NANN_OK;
  if (s->kind == 0) {
    NANN_LAUNCH(mlp_hoist_kernel, B, MLP_H, 0, st, users_dev, s->W1uT, s->b1, hu, B);
    return NANN_OK;
  }
  return attn_prepare_users(s, users_dev, B, hu, st);
}

static nann_status scorer_score(nann_scorer* s, const ScoreCall& c, cudaStream_t st) {
  if (c.B <= 0 || c.max_n <= 0) return NANN_OK;
  if (s->kind == 0) {
    if (s->precision == NANN_SCORER_TENSOR) return mlp_tc_score(s, c, st);
    MlpExactArgs a{};
    a.table = c.table; a.ids = c.ids; a.ids_stride = c.ids_stride; a.rows_stride = c.rows_stride;
    a.n_ptr = c.n_ptr; a.n_fixed = c.n_fixed; a.hu = c.hu;
    a.W1xT = s->W1xT; a.W2T = s->W2T; a.b2 = s->b2; a.w3 = s->w3;
    a.out = c.out; a.out_stride = c.out_stride; a.status = c.status;
    static bool attr_set[64] = {false};
    if (!attr_set[s->device & 63]) {
      NANN_CUDA(cudaFuncSetAttribute(mlp_exact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, EX_SMEM_BYTES));
      attr_set[s->device & 63] = true;
    }
    dim3 grid((unsigned)ceil_div(c.max_n, EX_TM), (unsigned)c.B);
    NANN_LAUNCH(mlp_exact_kernel, grid, EX_THREADS, EX_SMEM_BYTES, st, a);
    return NANN_OK;
  }
  return attn_score(s, c, st);
}

static size_t scorer_user_state_floats(const nann_scorer* s) {
  return s->kind == 0 ? (size_t)MLP_H : (size_t)ATT_L * ATT_QK;
}

}  // namespace nann

extern "C" {

nann_status nann_scorer_create_mlp(int d, int H, const float* W1, const float* b1, const float* W2,
                                   const float* b2, const float* w3, int device, nann_scorer_t** out) {
  if (!out) return fail(NANN_INVALID_ARGUMENT, "null out");
  *out = nullptr;
  NANN_TRY(require_device());
  if (d != MLP_D || H != MLP_H)
    return fail(NANN_UNIMPLEMENTED, "mlp scorer is built for d=%d H=%d (got d=%d H=%d)", MLP_D, MLP_H, d, H);
  NANN_CUDA(cudaSetDevice(device));
  cudaStream_t st = 0;
  auto* s = new nann_scorer();
  s->kind = 0; s->device = device; s->d = d; s->H = H;
  DevIn<float> dW1, dW2;
  nann_status rc = NANN_OK;
  auto guard = [&](nann_status r) { if (r != NANN_OK && rc == NANN_OK) rc = r; return r == NANN_OK; };
  if (guard(dW1.init(W1, (int64_t)H * 2 * d, st)) && guard(dW2.init(W2, (int64_t)H * H, st)) &&
      guard(to_device_copy(b1, H, &s->b1, st)) && guard(to_device_copy(b2, H, &s->b2, st)) &&
      guard(to_device_copy(w3, H, &s->w3, st))) {
    if (cudaMalloc(&s->W1uT, (size_t)d * H * 4) != cudaSuccess || cudaMalloc(&s->W1xT, (size_t)d * H * 4) != cudaSuccess ||
        cudaMalloc(&s->W2T, (size_t)H * H * 4) != cudaSuccess) {
      rc = fail(NANN_RESOURCE_EXHAUSTED, "OOM for scorer weights");
    } else {
      NANN_LAUNCH(transpose_kernel, (unsigned)ceil_div((int64_t)H * d, 256), 256, 0, st, dW1.d, H, d, 2 * d, 0, s->W1uT);
      NANN_LAUNCH(transpose_kernel, (unsigned)ceil_div((int64_t)H * d, 256), 256, 0, st, dW1.d, H, d, 2 * d, d, s->W1xT);
      NANN_LAUNCH(transpose_kernel, (unsigned)ceil_div((int64_t)H * H, 256), 256, 0, st, dW2.d, H, H, H, 0, s->W2T);
      if (cudaStreamSynchronize(st) != cudaSuccess) rc = fail(NANN_INTERNAL, "scorer weight upload failed");
    }
  }
  if (rc != NANN_OK) { nann_scorer_destroy(s); return rc; }
  *out = s;
  return NANN_OK;
}

nann_status nann_scorer_set_precision(nann_scorer_t* s, int precision) {
  if (!s) return fail(NANN_INVALID_ARGUMENT, "null scorer");
  if (precision != NANN_SCORER_EXACT && precision != NANN_SCORER_TENSOR)
    return fail(NANN_INVALID_ARGUMENT, "precision %d", precision);
  if (precision == NANN_SCORER_TENSOR) {
    if (s->kind != 0) return fail(NANN_UNIMPLEMENTED, "tensor-core path exists for the mlp scorer only");
    NANN_TRY(mlp_tc_prepare(s));
  }
  s->precision = precision;
  return NANN_OK;
}
int nann_scorer_user_floats(const nann_scorer_t* s) { return s ? (s->kind == 0 ? s->d : ATT_L * ATT_E) : 0; }
int nann_scorer_item_dim(const nann_scorer_t* s) { return s ? (s->kind == 0 ? s->d : ATT_E) : 0; }
void nann_scorer_destroy(nann_scorer_t* s) {
  if (!s) return;
  cudaFree(s->W1uT); cudaFree(s->W1xT); cudaFree(s->W2T); cudaFree(s->b1); cudaFree(s->b2); cudaFree(s->w3);
  cudaFree(s->blob);
  mlp_tc_release(s);
  delete s;
}

static nann_status scorer_run_common(nann_scorer_t* s, const float* user, const float* table, int64_t n_rows,
                                     const int32_t* ids, int64_t n, float* logits, void* stream) {
  NANN_TRY(require_device());
  if (!s) return fail(NANN_INVALID_ARGUMENT, "null scorer");
  if (n <= 0) return NANN_OK;
  if (n > 0x7fffffff) return fail(NANN_UNIMPLEMENTED, "n > 2^31-1");
  cudaStream_t st = (cudaStream_t)stream;
  NANN_CUDA(cudaSetDevice(s->device));
  const int dim = nann_scorer_item_dim(s);
  DevIn<float> d_user, d_tab;
  DevIn<int32_t> d_ids;
  DevOut<float> d_out;
  DevBuf<float> hu;
  NANN_TRY(d_user.init(user, nann_scorer_user_floats(s), st));
  NANN_TRY(d_tab.init(table, (ids ? n_rows : n) * dim, st));
  NANN_TRY(d_ids.init(ids, ids ? n : 0, st));
  NANN_TRY(d_out.init(logits, n, st, false));
  NANN_TRY(hu.alloc((int64_t)scorer_user_state_floats(s)));
  if (ids) {  // GatherV2 bounds (InvalidArgument in TF)
    DevBuf<int> bad;
    NANN_TRY(bad.alloc(1));
    NANN_CUDA(cudaMemsetAsync(bad.d, 0, sizeof(int), st));
    NANN_LAUNCH(check_i32_range_kernel, (unsigned)std::min<int64_t>(ceil_div(n, 256), 1184), 256, 0, st, d_ids.d, n,
                n_rows, bad.d);
    int b = 0;
    NANN_CUDA(cudaMemcpyAsync(&b, bad.d, sizeof(int), cudaMemcpyDeviceToHost, st));
    NANN_CUDA(cudaStreamSynchronize(st));
    if (b) return fail(NANN_INVALID_ARGUMENT, "indices out of range [0, %lld)", (long long)n_rows);
  }
  NANN_TRY(scorer_prepare_users(s, d_user.d, 1, hu.d, st));
  ScoreCall c{};
  c.table = d_tab.d; c.ids = d_ids.d; c.ids_stride = 0; c.rows_stride = 0;
  c.n_ptr = nullptr; c.n_fixed = (int)n; c.max_n = (int)n; c.B = 1;
  c.hu = hu.d; c.users = d_user.d; c.out = d_out.d; c.out_stride = 0; c.status = nullptr;
  TcWorkspace ws;   // op-level call: private scratch for the duration of the call
  c.ws = &ws;
  nann_status rc_score = scorer_score(s, c, st);
  if (rc_score == NANN_OK) rc_score = d_out.finish(st);
  cudaError_t sync_err = cudaStreamSynchronize(st);
  tc_ws_free(&ws);
  NANN_TRY(rc_score);
  NANN_CUDA(sync_err);
  return NANN_OK;
}

nann_status nann_blaze_xla_run(nann_scorer_t* s, const float* user, const float* item_emb, int64_t n,
                               float* logits, void* stream) {
  if (!s) return fail(NANN_INVALID_ARGUMENT, "null scorer");
  NANN_TRY(s->gate.enter());
  const nann_status rc = scorer_run_common(s, user, item_emb, n, nullptr, n, logits, stream);
  s->gate.leave();
  return rc;
}
nann_status nann_scorer_run_ids(nann_scorer_t* s, const float* user, const float* table, int64_t n_rows,
                                const int32_t* ids, int64_t n, float* logits, void* stream) {
  if (!ids) return fail(NANN_INVALID_ARGUMENT, "ids is NULL");
  if (!s) return fail(NANN_INVALID_ARGUMENT, "null scorer");
  NANN_TRY(s->gate.enter());
  const nann_status rc = scorer_run_common(s, user, table, n_rows, ids, n, logits, stream);
  s->gate.leave();
  return rc;
}
nann_status nann_scorer_set_admission(nann_scorer_t* s, int running_max, int max_waiting, int wait_ms) {
  if (!s) return fail(NANN_INVALID_ARGUMENT, "null scorer");
  std::lock_guard<std::mutex> lk(s->gate.mu);
  if (running_max >= 1) s->gate.running_max = running_max;
  if (max_waiting >= 0) s->gate.max_waiting = max_waiting;
  if (wait_ms >= 0) s->gate.wait_ns = (long long)wait_ms * 1000000ll;
  return NANN_OK;
}
nann_status nann_scorer_admission_state(nann_scorer_t* s, int* running, int* waiting, int* running_max, int* max_waiting,
                                        int* wait_ms) {
  if (!s) return fail(NANN_INVALID_ARGUMENT, "null scorer");
  std::lock_guard<std::mutex> lk(s->gate.mu);
  if (running) *running = s->gate.running;
  if (waiting) *waiting = s->gate.waiting;
  if (running_max) *running_max = s->gate.running_max;
  if (max_waiting) *max_waiting = s->gate.max_waiting;
  if (wait_ms) *wait_ms = (int)(s->gate.wait_ns / 1000000ll);
  return NANN_OK;
}

// ---- index -----------------------------------------------------------------------------------
nann_status nann_index_create(int64_t n_items, int dim, const void* emb, int emb_dtype,
                              const int64_t* item_ids, const void* enter_points, int ep_dtype,
                              int64_t n_ep, const void* const nbr_values[2], const int64_t n_nbr_values[2],
                              int nbr_dtype, const int64_t* const nbr_row_splits[2], int device, nann_index_t** out) {
  return nann_index_create_sharded(n_items, dim, emb, emb_dtype, 0, n_items, item_ids, enter_points, ep_dtype, n_ep,
                                   nbr_values, n_nbr_values, nbr_dtype, nbr_row_splits, device, out);
}

nann_status nann_index_create_sharded(int64_t n_items, int dim, const void* emb, int emb_dtype, int64_t row_lo, int64_t n_local,
                                      const int64_t* item_ids, const void* enter_points, int ep_dtype,
                                      int64_t n_ep, const void* const nbr_values[2], const int64_t n_nbr_values[2],
                                      int nbr_dtype, const int64_t* const nbr_row_splits[2], int device, nann_index_t** out) {
  if (!out) return fail(NANN_INVALID_ARGUMENT, "null out");
  *out = nullptr;
  NANN_TRY(require_device());
  if (n_items <= 0 || dim <= 0 || !emb || !item_ids || !enter_points || !nbr_values || !n_nbr_values || !nbr_row_splits)
    return fail(NANN_INVALID_ARGUMENT, "nann_index_create: null or empty input");
  if (n_items > 0x7fffffffll) return fail(NANN_UNIMPLEMENTED, "n_items > 2^31-1 per shard");
  if (row_lo < 0 || n_local <= 0 || row_lo + n_local > n_items)
    return fail(NANN_INVALID_ARGUMENT, "table rows [%lld, %lld) outside [0, %lld)", (long long)row_lo, (long long)(row_lo + n_local), (long long)n_items);
  NANN_CUDA(cudaSetDevice(device));
  cudaStream_t st = 0;
  auto* ix = new nann_index();
  ix->device = device; ix->n_items = n_items; ix->dim = dim; ix->n_ep = n_ep;
  ix->row_lo = row_lo; ix->n_local = n_local;
  DevBuf<int> flags;  // [0]=bad ids, [1]=bad csr, [2]=ep order, [3..4]=max degree
  nann_status rc = flags.alloc(8);
  auto step = [&](nann_status r) { if (rc == NANN_OK) rc = r; return rc == NANN_OK; };
  if (rc == NANN_OK) cudaMemsetAsync(flags.d, 0, 8 * sizeof(int), st);
  if (rc == NANN_OK) {
    if (emb_dtype == NANN_F32) step(to_device_copy((const float*)emb, n_local * dim, &ix->emb, st));
    else if (emb_dtype == NANN_F16) {  // widened once in memory (never on disk)
      DevIn<__half> h;
      if (step(h.init((const __half*)emb, n_local * dim, st)) &&
          step(cudaMalloc(&ix->emb, (size_t)n_local * dim * 4) == cudaSuccess ? NANN_OK
                   : fail(NANN_RESOURCE_EXHAUSTED, "OOM for item_embs"))) {
        NANN_LAUNCH(f16_to_f32_kernel, 148 * 8, 256, 0, st, h.d, ix->emb, n_local * dim);
        cudaStreamSynchronize(st);
      }
    } else rc = fail(NANN_INVALID_ARGUMENT, "item_embs must be f16 or f32");
  }
  step(to_device_copy(item_ids, n_items, &ix->item_ids, st));
  step(ids_to_device_i32(enter_points, ep_dtype, n_ep, n_items, &ix->ep, flags.d + 0, st));
  if (rc == NANN_OK && n_ep > 1)
    NANN_LAUNCH(ascending_check_kernel, (unsigned)std::min<int64_t>(ceil_div(n_ep, 256), 1184), 256, 0, st, ix->ep, n_ep, flags.d + 2);
  for (int l = 0; l < 2 && rc == NANN_OK; ++l) {
    if (!step(to_device_copy(nbr_row_splits[l], n_items + 1, &ix->nbr_rs[l], st))) break;
    int64_t nv = 0;
    if (cudaMemcpyAsync(&nv, ix->nbr_rs[l] + n_items, 8, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess) { rc = fail(NANN_INTERNAL, "row_splits readback failed"); break; }
    if (nv != n_nbr_values[l]) {   // ValidateRaggedTensor code 3 (GroupGather_kernel.cc:14)
      rc = fail(NANN_INVALID_ARGUMENT, "Invalid RaggedTensor neighbors level %d, code: 3 (row_splits[-1]=%lld, %lld values)", l,
                (long long)nv, (long long)n_nbr_values[l]);
      break;
    }
    ix->n_nbr[l] = nv;
    if (!step(ids_to_device_i32(nbr_values[l], nbr_dtype, nv, n_items, &ix->nbr_values[l], flags.d + 0, st))) break;
    NANN_LAUNCH(csr_check_kernel, (unsigned)std::min<int64_t>(ceil_div(n_items, 256), 1184), 256, 0, st, ix->nbr_rs[l], n_items,
                nv, flags.d + 3 + l, flags.d + 1);
  }
  int hf[8] = {0};
  if (rc == NANN_OK) {
    if (cudaMemcpyAsync(hf, flags.d, sizeof(hf), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess)
      rc = fail(NANN_INTERNAL, "index validation failed: %s", cudaGetErrorString(cudaGetLastError()));
    else if (hf[0]) rc = fail(NANN_INVALID_ARGUMENT, "node id outside [0, n_items) in enter_points or neighbors");
    else if (hf[1]) rc = fail(NANN_INVALID_ARGUMENT, "Invalid RaggedTensor neighbors, code: %d", hf[1]);
    else if (hf[2]) rc = fail(NANN_INVALID_ARGUMENT, "enter_points must be strictly ascending (np.nonzero order)");
  }
  if (rc != NANN_OK) { nann_index_destroy(ix); return rc; }
  ix->max_deg[0] = hf[3]; ix->max_deg[1] = hf[4];
  *out = ix;
  return NANN_OK;
}

static nann_status load_npy_any(const std::string& path, int want_a, int want_b, int device,
                                nann_huge_const_t** h, int* dtype, std::vector<int64_t>* shape) {
  int rank = 0;
  int64_t shp[8];
  NANN_TRY(nann_npy_peek(path.c_str(), dtype, &rank, shp));
  if (*dtype != want_a && *dtype != want_b)
    return fail(NANN_INTERNAL, "DataType mismatch in %s: dtype code %d", path.c_str(), *dtype);
  shape->assign(shp, shp + rank);
  return nann_huge_const_create(path.c_str(), *dtype, shp, rank, -1, h);
  (void)device;
}

nann_status nann_index_load(const char* embs_dir, const char* index_dir, int device, nann_index_t** out) {
  if (!out || !embs_dir || !index_dir) return fail(NANN_INVALID_ARGUMENT, "null argument");
  *out = nullptr;
  NANN_TRY(require_device());
  std::string ed(embs_dir), id(index_dir);
  nann_huge_const_t *h_emb = nullptr, *h_ids = nullptr, *h_ep = nullptr, *h_v[2] = {nullptr, nullptr}, *h_rs[2] = {nullptr, nullptr};
  int dt_emb, dt_ids, dt_ep, dt_v[2], dt_rs[2];
  std::vector<int64_t> s_emb, s_ids, s_ep, s_v[2], s_rs[2];
  nann_status rc = NANN_OK;
  auto step = [&](nann_status r) { if (rc == NANN_OK) rc = r; return rc == NANN_OK; };
  step(load_npy_any(ed + "/item_embs.npy", NANN_F32, NANN_F16, device, &h_emb, &dt_emb, &s_emb));
  step(load_npy_any(ed + "/item_ids.npy", NANN_I64, NANN_I64, device, &h_ids, &dt_ids, &s_ids));
  step(load_npy_any(id + "/enter_points.npy", NANN_I64, NANN_I32, device, &h_ep, &dt_ep, &s_ep));
  for (int l = 0; l < 2; ++l) {
    step(load_npy_any(id + "/neighbors_level_" + std::to_string(l) + "_values.npy", NANN_I64, NANN_I32, device, &h_v[l], &dt_v[l], &s_v[l]));
    step(load_npy_any(id + "/neighbors_level_" + std::to_string(l) + "_row_splits.npy", NANN_I64, NANN_I64, device, &h_rs[l], &dt_rs[l], &s_rs[l]));
  }
  if (rc == NANN_OK) {
    if (s_emb.size() != 2 || s_ids.size() != 1 || s_ids[0] != s_emb[0] || s_ep.size() != 1 ||
        s_rs[0].size() != 1 || s_rs[0][0] != s_emb[0] + 1 || s_rs[1].size() != 1 || s_rs[1][0] != s_emb[0] + 1 ||
        dt_v[0] != dt_v[1])
      rc = fail(NANN_INTERNAL, "attr_shape and np_shape NOT match (index files inconsistent)");
  }
  if (rc == NANN_OK) {
    const void* vals[2] = {nann_huge_const_host(h_v[0]), nann_huge_const_host(h_v[1])};
    const int64_t* rss[2] = {(const int64_t*)nann_huge_const_host(h_rs[0]), (const int64_t*)nann_huge_const_host(h_rs[1])};
    const int64_t nvs[2] = {s_v[0].empty() ? 0 : s_v[0][0], s_v[1].empty() ? 0 : s_v[1][0]};
    rc = nann_index_create(s_emb[0], (int)s_emb[1], nann_huge_const_host(h_emb), dt_emb,
                           (const int64_t*)nann_huge_const_host(h_ids), nann_huge_const_host(h_ep), dt_ep, s_ep[0],
                           vals, nvs, dt_v[0], rss, device, out);
  }
  nann_huge_const_destroy(h_emb); nann_huge_const_destroy(h_ids); nann_huge_const_destroy(h_ep);
  for (int l = 0; l < 2; ++l) { nann_huge_const_destroy(h_v[l]); nann_huge_const_destroy(h_rs[l]); }
  return rc;
}

int64_t nann_index_n_items(const nann_index_t* ix) { return ix ? ix->n_items : 0; }
int nann_index_dim(const nann_index_t* ix) { return ix ? ix->dim : 0; }
int64_t nann_index_n_enter_points(const nann_index_t* ix) { return ix ? ix->n_ep : 0; }
const float* nann_index_emb_device(const nann_index_t* ix) { return ix ? ix->emb : nullptr; }
void nann_index_destroy(nann_index_t* ix) {
  if (!ix) return;
  cudaFree(ix->emb); cudaFree(ix->item_ids); cudaFree(ix->ep);
  for (int l = 0; l < 2; ++l) { cudaFree(ix->nbr_values[l]); cudaFree(ix->nbr_rs[l]); }
  delete ix;
}

}  // extern "C"

// ---- searcher ------------------------------------------------------------------------------------
struct nann_dist_group;   // lib_dist.inl
struct nann_searcher {
  nann_dist_group* dist = nullptr;   // set for the duration of a nann_search_distributed call
  const nann_index* ix = nullptr;
  nann_scorer* sc = nullptr;
  int max_batch = 0;
  int maxT[6] = {0};
  int64_t n_words = 0;   // bitmap words per query = ceil(n_items/32)  (build_opt_graph.py:115)
  int64_t maxc = 0;      // candidate capacity per query per round
  int64_t maxr = 0;      // result list capacity T1+T2+T3+T4
  float* users = nullptr;      // [maxB][user_floats]
  float* ustate = nullptr;     // [maxB][user_state]
  uint32_t* bitmap = nullptr;  // [maxB][n_words]
  int32_t* cand_ids = nullptr; float* cand_sc = nullptr;   // [maxB][maxc]
  int32_t* round_n = nullptr;  int32_t* round_exp = nullptr;  // [5][maxB]
  int32_t* r0_ids = nullptr; float* r0_sc = nullptr;       // [maxB][T0]
  int32_t* res_ids = nullptr; float* res_sc = nullptr;     // [maxB][maxr]
  int32_t* out_nodes = nullptr; float* out_sc = nullptr; int64_t* out_item = nullptr;  // [maxB][T5]
  int32_t* status = nullptr;   // [maxB]
  nann::TcWorkspace* tcws = nullptr;                        // tensor-core scorer scratch (lazy)
  bool trace = false;
  int32_t* tr_ids = nullptr; float* tr_sc = nullptr;       // [5][maxB][maxc] when tracing
  // CUDA graphs of the launch sequence for small batches, keyed by (B, level_topn, scorer precision)
  struct GraphKey {
    int B, precision, T[6];
    bool operator<(const GraphKey& o) const { return memcmp(this, &o, sizeof(GraphKey)) < 0; }
  };
  struct GraphVal { cudaGraphExec_t exec = nullptr; uint64_t kernels = 0; };   // kernels per replay (for the launch counter)
  std::map<GraphKey, GraphVal> graphs;
  cudaStream_t own_stream = nullptr;   // for host-in / host-out calls on the NULL stream (which cannot be captured), lazily
  // host mirrors of the last call
  std::vector<int32_t> h_round_n, h_round_exp, h_status;
  int last_B = 0, last_k = 0;
  // stage timing (CUDA events on the launching stream), accumulated over calls while enabled
  bool profile = false;
  std::vector<cudaEvent_t> ev;       // pairs (start, stop)
  std::vector<int> ev_stage;         // stage id of each pair
  int ev_used = 0;
  double stage_ms[4] = {0, 0, 0, 0}; // 0 score (gather+model), 1 expand+filter, 2 top-k, 3 other
  int64_t stage_launches[4] = {0, 0, 0, 0};
  int64_t prof_rows = 0, prof_calls = 0;
};

extern "C" {

void nann_searcher_destroy(nann_searcher_t* s) {
  if (!s) return;
  cudaFree(s->users); cudaFree(s->ustate); cudaFree(s->bitmap); cudaFree(s->cand_ids); cudaFree(s->cand_sc);
  cudaFree(s->round_n); cudaFree(s->round_exp); cudaFree(s->r0_ids); cudaFree(s->r0_sc);
  cudaFree(s->res_ids); cudaFree(s->res_sc); cudaFree(s->out_nodes); cudaFree(s->out_sc);
  cudaFree(s->out_item); cudaFree(s->status); cudaFree(s->tr_ids); cudaFree(s->tr_sc);
  for (auto e : s->ev) cudaEventDestroy(e);
  for (auto& kv : s->graphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  if (s->own_stream) cudaStreamDestroy(s->own_stream);
  if (s->tcws) { nann::tc_ws_free(s->tcws); delete s->tcws; }
  delete s;
}

nann_status nann_searcher_create(const nann_index_t* ix, nann_scorer_t* scorer, int max_batch,
                                 const int32_t maxT[6], nann_searcher_t** out) {
  if (!out) return fail(NANN_INVALID_ARGUMENT, "null out");
  *out = nullptr;
  NANN_TRY(require_device());
  if (!ix || !scorer || max_batch <= 0 || !maxT) return fail(NANN_INVALID_ARGUMENT, "nann_searcher_create: bad argument");
  if (max_batch > 65535) return fail(NANN_UNIMPLEMENTED, "max_batch > 65535");
  if (nann_scorer_item_dim(scorer) != ix->dim)
    return fail(NANN_INVALID_ARGUMENT, "scorer item dim %d != index dim %d", nann_scorer_item_dim(scorer), ix->dim);
  for (int i = 0; i < 6; ++i)
    if (maxT[i] < 0 || maxT[i] > TOPK_MAX_K) return fail(NANN_UNIMPLEMENTED, "level_topn[%d]=%d outside [0,%d]", i, maxT[i], TOPK_MAX_K);
  NANN_CUDA(cudaSetDevice(ix->device));
  auto* s = new nann_searcher();
  s->ix = ix; s->sc = scorer; s->max_batch = max_batch;
  for (int i = 0; i < 6; ++i) s->maxT[i] = maxT[i];
  s->n_words = (ix->n_items + 31) / 32;
  const int64_t front0 = std::max<int64_t>(std::max(maxT[1], maxT[2]), maxT[3]);
  s->maxc = std::max<int64_t>(std::max<int64_t>(ix->n_ep, (int64_t)ix->max_deg[1] * maxT[0]), (int64_t)ix->max_deg[0] * front0);
  s->maxc = std::max<int64_t>((s->maxc + 63) / 64 * 64, 64);
  s->maxr = std::max<int64_t>((int64_t)maxT[1] + maxT[2] + maxT[3] + maxT[4], 1);
  const int64_t B = max_batch;
  const int uf = nann_scorer_user_floats(scorer);
  bool ok = true;
  auto A = [&](void** p, size_t bytes) { if (ok && cudaMalloc(p, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); ok = false; } };
  A((void**)&s->users, (size_t)B * uf * 4);
  A((void**)&s->ustate, (size_t)B * scorer_user_state_floats(scorer) * 4);
  A((void**)&s->bitmap, (size_t)B * s->n_words * 4);
  A((void**)&s->cand_ids, (size_t)B * s->maxc * 4);
  A((void**)&s->cand_sc, (size_t)B * s->maxc * 4);
  A((void**)&s->round_n, (size_t)5 * B * 4);
  A((void**)&s->round_exp, (size_t)5 * B * 4);
  A((void**)&s->r0_ids, (size_t)B * std::max(maxT[0], 1) * 4);
  A((void**)&s->r0_sc, (size_t)B * std::max(maxT[0], 1) * 4);
  A((void**)&s->res_ids, (size_t)B * s->maxr * 4);
  A((void**)&s->res_sc, (size_t)B * s->maxr * 4);
  A((void**)&s->out_nodes, (size_t)B * std::max(maxT[5], 1) * 4);
  A((void**)&s->out_sc, (size_t)B * std::max(maxT[5], 1) * 4);
  A((void**)&s->out_item, (size_t)B * std::max(maxT[5], 1) * 8);
  A((void**)&s->status, (size_t)B * 4);
  if (!ok) {
    nann_searcher_destroy(s);
    return fail(NANN_RESOURCE_EXHAUSTED, "OOM for the search workspace (batch %d, %lld bitmap words/query, %lld candidates/query)",
                max_batch, (long long)s->n_words, (long long)s->maxc);
  }
  *out = s;
  return NANN_OK;
}

nann_status nann_searcher_set_trace(nann_searcher_t* s, int enable) {
  if (!s) return fail(NANN_INVALID_ARGUMENT, "null searcher");
  if (enable && !s->tr_ids) {
    const size_t bytes = (size_t)5 * s->max_batch * s->maxc * 4;
    if (cudaMalloc(&s->tr_ids, bytes) != cudaSuccess || cudaMalloc(&s->tr_sc, bytes) != cudaSuccess) {
      cudaGetLastError();
      cudaFree(s->tr_ids); s->tr_ids = nullptr;
      return fail(NANN_RESOURCE_EXHAUSTED, "OOM for trace buffers (%zu bytes x2)", bytes);
    }
  }
  s->trace = enable != 0;
  return NANN_OK;
}

}  // extern "C"

namespace nann {

// Enqueues the whole dataflow for B queries on `st` and returns without synchronising: per-shard results end up in
// the searcher's out_sc / out_nodes / out_item ([B][max(k,1)]), per-query status and the per-round counters in
// s->status / s->round_n / s->round_exp.  `push` (optional) makes the final top-k deliver its records to the shard
// group's windows as well (lib_shard.inl).
static nann_status dist_share_user_state(nann_dist_group* g, const float* ustate, const float* users, int B, cudaStream_t st);
static nann_status dist_score_round(nann_searcher* s, nann_dist_group* g, const int32_t* ids, int64_t ids_stride, const int32_t* n_ptr,
                                    int n_fixed, int64_t bound, int B, cudaStream_t st);

__global__ void copy_floats_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

// everything after the users are in s->users: a fixed launch sequence whose arguments depend only on
// (searcher, B, level_topn, scorer precision) -- which is what makes it capturable as a CUDA graph
static nann_status search_core(nann_searcher* s, int B, const int32_t T[6], cudaStream_t st, const ShardPush* push) {
  const nann_index* ix = s->ix;
  const int64_t mb = s->max_batch;
  const int k = T[5];
  {   // per-call state in ONE kernel (no memset nodes: see fill_words_kernel)
    SearchInitArgs ia{};
    ia.status = s->status; ia.B = B;
    ia.round_n = s->round_n; ia.round_exp = s->round_exp; ia.n_round = 5 * mb;
    ia.out_item_w = (uint32_t*)s->out_item; ia.out_sc_w = (uint32_t*)s->out_sc; ia.out_nodes_w = (uint32_t*)s->out_nodes;
    ia.n_out = (int64_t)B * k;
    ia.users_src = nullptr; ia.users_dst = s->users; ia.n_users = 0;
    const int64_t work = std::max<int64_t>(std::max<int64_t>(ia.n_round, ia.n_out), 1);
    NANN_LAUNCH(search_init_kernel, (unsigned)std::min<int64_t>(ceil_div(work, 256), 148 * 4), 256, 0, st, ia);
  }
  NANN_TRY(scorer_prepare_users(s->sc, s->users, B, s->ustate, st));
  if (s->dist) NANN_TRY(dist_share_user_state(s->dist, s->ustate, s->users, B, st));

  s->ev_used = 0;
  auto t_begin = [&](int stage) {
    if (!s->profile) return;
    if ((size_t)s->ev_used + 2 > s->ev.size()) {
      cudaEvent_t a, b;
      cudaEventCreate(&a); cudaEventCreate(&b);
      s->ev.push_back(a); s->ev.push_back(b); s->ev_stage.push_back(stage);
    }
    s->ev_stage[s->ev_used / 2] = stage;
    cudaEventRecord(s->ev[s->ev_used], st);
  };
  auto t_end = [&]() {
    if (!s->profile) return;
    cudaEventRecord(s->ev[s->ev_used + 1], st);
    s->ev_used += 2;
  };
  auto score_round = [&](int r, const int32_t* ids, int64_t ids_stride, const int32_t* n_ptr, int n_fixed,
                         int64_t bound) -> nann_status {
    ScoreCall c{};
    c.table = ix->emb; c.ids = ids; c.ids_stride = ids_stride; c.rows_stride = 0;
    c.n_ptr = n_ptr; c.n_fixed = n_fixed; c.max_n = (int)std::min<int64_t>(bound, s->maxc); c.B = B;
    c.hu = s->ustate; c.users = s->users; c.out = s->cand_sc; c.out_stride = s->maxc; c.status = s->status;
    if (!s->tcws) s->tcws = new nann::TcWorkspace();
    c.ws = s->tcws;
    t_begin(0);
    nann_status rc_score = s->dist ? dist_score_round(s, s->dist, ids, ids_stride, n_ptr, n_fixed, bound, B, st)
                                   : scorer_score(s->sc, c, st);
    t_end();
    NANN_TRY(rc_score);
    if (s->trace) {
      const size_t bytes = (size_t)B * s->maxc * 4;
      if (ids_stride != 0)
        NANN_CUDA(cudaMemcpyAsync(s->tr_ids + (size_t)r * mb * s->maxc, ids, bytes, cudaMemcpyDeviceToDevice, st));
      NANN_CUDA(cudaMemcpyAsync(s->tr_sc + (size_t)r * mb * s->maxc, s->cand_sc, bytes, cudaMemcpyDeviceToDevice, st));
    }
    return NANN_OK;
  };
  auto expand_round = [&](int r, int level, const int32_t* frontier, int64_t f_stride, int f_n) -> nann_status {
    const int threads = 128;
    // CTA-per-query kernel when the round's expansion (f_n * max_deg ids) and its hash table fit in shared memory
    static const bool force_warp = [] { const char* e = std::getenv("NANN_EXPAND_WARP"); return e && atoi(e) != 0; }();
    const int64_t cap = (int64_t)f_n * ix->max_deg[level];
    int64_t tsize = 1024;
    while (tsize < cap) tsize <<= 1;
    if (cap * 4 + tsize * 4 <= 200 * 1024) tsize <<= 1;          // halve the load factor when there is room
    const int64_t smem = cap * 4 + tsize * 2;
    t_begin(1);
    if (!force_warp && f_n > 0 && f_n <= EFC_MAX_FRONTIER && cap > 0 && cap < 65535 && smem <= 200 * 1024) {
      // once per device (a process may hold searchers on several GPUs).  NOT on every call: the driver call can wait for
      // launches of the function that are still pending, and a pending launch may sit behind a kernel that waits for work
      // this host thread has yet to enqueue (two group members driven from one thread)
      static std::atomic<bool> efc_attr[64];
      if (!efc_attr[ix->device & 63].load(std::memory_order_acquire)) {
        NANN_CUDA(cudaFuncSetAttribute(expand_filter_cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        efc_attr[ix->device & 63].store(true, std::memory_order_release);
      }
      NANN_LAUNCH(expand_filter_cta_kernel, (unsigned)B, EFC_THREADS, (size_t)smem, st,
                  ix->nbr_values[level], ix->nbr_rs[level], frontier, f_stride, f_n, s->bitmap, s->n_words,
                  s->cand_ids, s->maxc, s->round_n + r * mb, s->round_exp + r * mb, s->status, (int)cap, (int)tsize);
    } else {
      NANN_LAUNCH(expand_filter_kernel, (unsigned)ceil_div((int64_t)B * 32, threads), threads, 0, st,
                  ix->nbr_values[level], ix->nbr_rs[level], frontier, f_stride, f_n, s->bitmap, s->n_words,
                  s->cand_ids, s->maxc, s->round_n + r * mb, s->round_exp + r * mb, s->status, B);
    }
    t_end();
    return NANN_OK;
  };
  auto mark = [&](const int32_t* list, int64_t stride, int n) -> nann_status {
    t_begin(3);
    NANN_LAUNCH(fill_words_kernel, 148 * 8, 256, 0, st, s->bitmap, (int64_t)B * s->n_words, 0u);  // Assign zeros :118,:131
    if (n > 0)
      NANN_LAUNCH(mark_kernel, (unsigned)ceil_div((int64_t)B * n, 256), 256, 0, st, list, stride, n, s->bitmap,
                  s->n_words, s->status, B);
    t_end();
    return NANN_OK;
  };
  auto topk_timed = [&](const TopkArgs& a) -> nann_status {
    t_begin(2);
    nann_status rc_topk = launch_topk(a, B, st);
    t_end();
    return rc_topk;
  };

  // ---- level 2 (:109-112): score every enter point, top T[0]
  NANN_TRY(score_round(0, ix->ep, 0, nullptr, (int)ix->n_ep, ix->n_ep));
  {
    TopkArgs a{};
    a.b_sc = s->cand_sc; a.b_sc_stride = s->maxc; a.b_ids = ix->ep; a.b_ids_stride = 0; a.b_n_fixed = (int)ix->n_ep;
    a.k = T[0]; a.out_sc = s->r0_sc; a.out_ids = s->r0_ids; a.out_stride = std::max(s->maxT[0], 1);
    a.status = s->status; a.reject_single = 1;
    NANN_TRY(topk_timed(a));
  }
  // ---- level 1 (:114-127)
  NANN_TRY(mark(s->r0_ids, std::max(s->maxT[0], 1), T[0]));
  NANN_TRY(expand_round(1, 1, s->r0_ids, std::max(s->maxT[0], 1), T[0]));
  NANN_TRY(score_round(1, s->cand_ids, s->maxc, s->round_n + 1 * mb, 0, (int64_t)ix->max_deg[1] * T[0]));
  {
    TopkArgs a{};
    a.a_sc = s->r0_sc; a.a_ids = s->r0_ids; a.a_stride = std::max(s->maxT[0], 1); a.a_n = T[0];
    a.b_sc = s->cand_sc; a.b_sc_stride = s->maxc; a.b_ids = s->cand_ids; a.b_ids_stride = s->maxc;
    a.b_n_ptr = s->round_n + 1 * mb;
    a.k = T[1]; a.out_sc = s->res_sc; a.out_ids = s->res_ids; a.out_stride = s->maxr; a.out_offset = 0;
    a.status = s->status; a.reject_single = 1;
    NANN_TRY(topk_timed(a));
  }
  // ---- level 0 (:128-141)
  NANN_TRY(mark(s->res_ids, s->maxr, T[1]));
  int64_t off = T[1];
  int64_t f_off = 0;
  int f_n = T[1];
  for (int i = 0; i < 3; ++i) {
    const int r = 2 + i;
    NANN_TRY(expand_round(r, 0, s->res_ids + f_off, s->maxr, f_n));
    NANN_TRY(score_round(r, s->cand_ids, s->maxc, s->round_n + r * mb, 0, (int64_t)ix->max_deg[0] * f_n));
    TopkArgs a{};
    a.b_sc = s->cand_sc; a.b_sc_stride = s->maxc; a.b_ids = s->cand_ids; a.b_ids_stride = s->maxc;
    a.b_n_ptr = s->round_n + r * mb;
    a.k = T[i + 2]; a.out_sc = s->res_sc; a.out_ids = s->res_ids; a.out_stride = s->maxr; a.out_offset = off;
    a.status = s->status; a.reject_single = 1;
    NANN_TRY(topk_timed(a));
    f_off = off; f_n = T[i + 2];
    off += T[i + 2];
  }
  // ---- final (:143-144): top T[5] of the concatenated result list, then item_ids gather
  {
    TopkArgs a{};
    a.a_sc = s->res_sc; a.a_ids = s->res_ids; a.a_stride = s->maxr; a.a_n = (int)off;
    a.b_n_fixed = 0;
    a.k = k; a.out_sc = s->out_sc; a.out_ids = s->out_nodes; a.out_stride = std::max(k, 1);
    a.item_ids = ix->item_ids; a.out_item_ids = s->out_item; a.out_item_stride = std::max(k, 1);
    a.status = s->status;
    if (push) {
      a.push = *push;
      if (push->need_done > 0)
        NANN_LAUNCH(shard_backpressure_kernel, 1, 32, 0, st, push->my_done, push->world, push->need_done, push->error);
    }
    NANN_TRY(topk_timed(a));
  }
  return NANN_OK;
}

// Small batches are launch-bound (batch 1: ~31 launches for ~0.2 ms of GPU work), so the core sequence of a
// (B, level_topn, precision) key is captured into a CUDA graph the second time the key is seen (the first call runs
// eagerly and sizes every lazily grown buffer) and replayed with ONE launch afterwards.  Not used while tracing or
// profiling (those add per-call work) or for the sharded push (its arguments carry the sequence number).
static int search_graph_max_batch() {
  static const int v = [] { const char* e = std::getenv("NANN_GRAPH_MAX_BATCH"); return e ? atoi(e) : 32; }();
  return v;
}

static nann_status search_enqueue(nann_searcher* s, const float* users, int B, const int32_t T[6], cudaStream_t st,
                                  const ShardPush* push) {
  const int uf = nann_scorer_user_floats(s->sc);
  if (is_device_ptr(users))
    NANN_LAUNCH(copy_floats_kernel, (unsigned)std::min<int64_t>(ceil_div((int64_t)B * uf, 256), 148 * 4), 256, 0, st, users, s->users, (int64_t)B * uf);
  else
    NANN_CUDA(cudaMemcpyAsync(s->users, users, (size_t)B * uf * 4, cudaMemcpyHostToDevice, st));
  s->last_B = B; s->last_k = T[5];
  const int graph_max_b = search_graph_max_batch();
  // (the legacy default stream cannot be captured: eager there, without provoking the error on every call)
  if (push || s->dist || s->trace || s->profile || B > graph_max_b || st == nullptr || st == cudaStreamLegacy)
    return search_core(s, B, T, st, push);
  nann_searcher::GraphKey key{};
  key.B = B; key.precision = s->sc->precision;
  for (int i = 0; i < 6; ++i) key.T[i] = T[i];
  auto it = s->graphs.find(key);
  if (it == s->graphs.end()) {                    // first sight: eager run, remember the key
    s->graphs.emplace(key, nann_searcher::GraphVal());
    return search_core(s, B, T, st, push);
  }
  if (it->second.exec == nullptr) {               // second sight: capture
    cudaGraph_t graph = nullptr;
    const uint64_t l0 = g_launches.load(std::memory_order_relaxed);
    if (cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed) != cudaSuccess) {
      cudaGetLastError();
      s->graphs.erase(it);                         // this stream cannot be captured: do not try again on every call
      return search_core(s, B, T, st, push);
    }
    const nann_status rc = search_core(s, B, T, st, push);
    const cudaError_t e = cudaStreamEndCapture(st, &graph);
    if (rc != NANN_OK || e != cudaSuccess || !graph) {
      cudaGetLastError();
      if (graph) cudaGraphDestroy(graph);
      s->graphs.erase(it);
      NANN_TRY(rc);
      return search_core(s, B, T, st, push);
    }
    cudaGraphExec_t exec = nullptr;
    const cudaError_t ei = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ei != cudaSuccess || !exec) {
      cudaGetLastError();
      s->graphs.erase(it);
      return search_core(s, B, T, st, push);
    }
    it->second.exec = exec;
    it->second.kernels = g_launches.load(std::memory_order_relaxed) - l0;   // counted while capturing: this replay is paid for
    NANN_CUDA(cudaGraphLaunch(exec, st));
    return NANN_OK;
  }
  NANN_CUDA(cudaGraphLaunch(it->second.exec, st));
  g_launches.fetch_add(it->second.kernels, std::memory_order_relaxed);      // nann_kernel_launch_count() counts kernels, not graphs
  return NANN_OK;
}

// after a stream synchronisation: fold the stage events of the last enqueue into the profile
static void search_collect_profile(nann_searcher* s, int B) {
  if (!s->profile) return;
  const int64_t mb = s->max_batch;
  for (int i = 0; i < s->ev_used; i += 2) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, s->ev[i], s->ev[i + 1]) == cudaSuccess) {
      s->stage_ms[s->ev_stage[i / 2]] += ms;
      s->stage_launches[s->ev_stage[i / 2]] += 1;
    }
  }
  s->prof_calls += 1;
  s->prof_rows += (int64_t)B * s->ix->n_ep;
  for (int q = 0; q < B; ++q)
    for (int r = 1; r < 5; ++r) s->prof_rows += s->h_round_n[(size_t)r * mb + q];
}

// results of the last enqueue -> caller.  All-device outputs and nothing the host has to look at (stats, profile):
// asynchronous, the caller orders later work on `st`; otherwise the call synchronises.
static nann_status search_deliver(nann_searcher* s, int B, int k, int64_t* out_item_ids, float* out_scores, int32_t* out_status,
                                  nann_search_stats_t* stats, cudaStream_t st) {
  const nann_index* ix = s->ix;
  const int64_t mb = s->max_batch;
  const bool dev_ids = !out_item_ids || is_device_ptr(out_item_ids), dev_sc = !out_scores || is_device_ptr(out_scores);
  const bool dev_st = !out_status || is_device_ptr(out_status);
  // device outputs are written by a KERNEL: a device-to-device cudaMemcpyAsync is an implicit synchronisation point
  // between streams ("a memory copy between two addresses to the same device memory", CUDA programming guide), which
  // deadlocks two group members driven from one process: the copy of member 0 waits behind its flag-waiting kernel and
  // member 1's kernels, issued after the copy, wait behind the copy
  auto copy_words = [&](void* dst, const void* src, int64_t n_words) -> nann_status {
    NANN_LAUNCH(copy_floats_kernel, (unsigned)std::min<int64_t>(ceil_div(n_words, 256), 148 * 4), 256, 0, st, (const float*)src,
                (float*)dst, n_words);
    return NANN_OK;
  };
  if (k > 0) {
    if (out_item_ids) {
      if (dev_ids) NANN_TRY(copy_words(out_item_ids, s->out_item, (int64_t)B * k * 2));
      else NANN_CUDA(cudaMemcpyAsync(out_item_ids, s->out_item, (size_t)B * k * 8, cudaMemcpyDeviceToHost, st));
    }
    if (out_scores) {
      if (dev_sc) NANN_TRY(copy_words(out_scores, s->out_sc, (int64_t)B * k));
      else NANN_CUDA(cudaMemcpyAsync(out_scores, s->out_sc, (size_t)B * k * 4, cudaMemcpyDeviceToHost, st));
    }
  }
  if (out_status && dev_st) NANN_TRY(copy_words(out_status, s->status, B));
  if (dev_ids && dev_sc && dev_st && !stats && !s->profile) return NANN_OK;

  s->h_round_n.resize((size_t)5 * mb); s->h_round_exp.resize((size_t)5 * mb); s->h_status.resize(B);
  NANN_CUDA(cudaMemcpyAsync(s->h_round_n.data(), s->round_n, (size_t)5 * mb * 4, cudaMemcpyDeviceToHost, st));
  NANN_CUDA(cudaMemcpyAsync(s->h_round_exp.data(), s->round_exp, (size_t)5 * mb * 4, cudaMemcpyDeviceToHost, st));
  NANN_CUDA(cudaMemcpyAsync(s->h_status.data(), s->status, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
  NANN_CUDA(cudaStreamSynchronize(st));
  search_collect_profile(s, B);
  if (out_status && !dev_st) memcpy(out_status, s->h_status.data(), (size_t)B * 4);
  if (stats) {
    for (int q = 0; q < B; ++q) {
      if (s->h_status[q] != 0) stats->n_failed++;
      for (int r = 1; r < 5; ++r) {
        stats->n_scored[r] += s->h_round_n[(size_t)r * mb + q];
        stats->n_expanded[r] += s->h_round_exp[(size_t)r * mb + q];
      }
    }
    stats->n_scored[0] = (int64_t)B * ix->n_ep;
    stats->n_expanded[0] = (int64_t)B * ix->n_ep;
  }
  return NANN_OK;
}

static nann_status search_check_args(nann_searcher* s, const float* users, int B, const int32_t T[6]) {
  if (!s || !users || !T) return fail(NANN_INVALID_ARGUMENT, "nann_search_batch: null argument");
  if (B < 0 || B > s->max_batch) return fail(NANN_INVALID_ARGUMENT, "batch %d outside [0, %d]", B, s->max_batch);
  for (int i = 0; i < 6; ++i) {
    if (T[i] < 0) return fail(NANN_INVALID_ARGUMENT, "Need k >= 0, got %d", T[i]);  // topk_op.cc:60-61
    if (T[i] > s->maxT[i]) return fail(NANN_INVALID_ARGUMENT, "level_topn[%d]=%d exceeds the searcher's maximum %d", i, T[i], s->maxT[i]);
  }
  return NANN_OK;
}

}  // namespace nann

extern "C" {

nann_status nann_search_batch(nann_searcher_t* s, const float* users, int B, const int32_t T[6],
                              int64_t* out_item_ids, float* out_scores, int32_t* out_status,
                              nann_search_stats_t* stats, void* stream) {
  NANN_TRY(require_device());
  NANN_TRY(search_check_args(s, users, B, T));
  if (s->ix->n_local != s->ix->n_items)
    return fail(NANN_FAILED_PRECONDITION, "this index holds rows [%lld, %lld) of the table only: use nann_search_distributed",
                (long long)s->ix->row_lo, (long long)(s->ix->row_lo + s->ix->n_local));
  if (stats) memset(stats, 0, sizeof(*stats));
  if (B == 0) return NANN_OK;
  const nann_index* ix = s->ix;
  cudaStream_t st = (cudaStream_t)stream;
  NANN_CUDA(cudaSetDevice(ix->device));
  // The reference's operating mode -- a small batch, host tensors in and out, no stream given (what the TF shim and the
  // serving front-end do): such a call synchronises before it returns and depends on no earlier GPU work, so it may run on
  // a stream of the searcher's own, where the launch sequence replays as a CUDA graph (the NULL stream cannot be captured).
  if ((st == nullptr || st == cudaStreamLegacy) && B <= search_graph_max_batch() && !is_device_ptr(users)) {
    const bool any_dev = (out_item_ids && is_device_ptr(out_item_ids)) || (out_scores && is_device_ptr(out_scores)) ||
                         (out_status && is_device_ptr(out_status));
    const bool will_sync = out_item_ids || out_scores || out_status || stats;
    if (!any_dev && will_sync) {
      if (!s->own_stream) NANN_CUDA(cudaStreamCreateWithFlags(&s->own_stream, cudaStreamNonBlocking));
      st = s->own_stream;
    }
  }
  NANN_TRY(search_enqueue(s, users, B, T, st, nullptr));
  return search_deliver(s, B, T[5], out_item_ids, out_scores, out_status, stats, st);
}

nann_status nann_searcher_get_trace(nann_searcher_t* s, int q, int round, int32_t* ids, float* scores,
                                    int64_t cap, int64_t* n) {
  if (!s || !s->trace || !s->tr_ids) return fail(NANN_FAILED_PRECONDITION, "tracing is off");
  if (q < 0 || q >= s->last_B || round < 0 || round > 4) return fail(NANN_INVALID_ARGUMENT, "bad q/round");
  const int64_t mb = s->max_batch;
  int64_t cnt = round == 0 ? s->ix->n_ep : s->h_round_n[(size_t)round * mb + q];
  if (n) *n = cnt;
  cnt = std::min(cnt, cap);
  if (cnt <= 0) return NANN_OK;
  const size_t o = ((size_t)round * mb + q) * s->maxc;
  if (ids) NANN_CUDA(cudaMemcpy(ids, round == 0 ? s->ix->ep : s->tr_ids + o, (size_t)cnt * 4, cudaMemcpyDeviceToHost));
  if (scores) NANN_CUDA(cudaMemcpy(scores, s->tr_sc + o, (size_t)cnt * 4, cudaMemcpyDeviceToHost));
  return NANN_OK;
}

nann_status nann_searcher_set_profile(nann_searcher_t* s, int enable) {
  if (!s) return fail(NANN_INVALID_ARGUMENT, "null searcher");
  s->profile = enable != 0;
  for (int i = 0; i < 4; ++i) { s->stage_ms[i] = 0; s->stage_launches[i] = 0; }
  s->prof_rows = 0; s->prof_calls = 0;
  return NANN_OK;
}
nann_status nann_searcher_get_profile(nann_searcher_t* s, double stage_ms[4], int64_t stage_launches[4],
                                      int64_t* rows_scored, int64_t* calls) {
  if (!s) return fail(NANN_INVALID_ARGUMENT, "null searcher");
  for (int i = 0; i < 4; ++i) {
    if (stage_ms) stage_ms[i] = s->stage_ms[i];
    if (stage_launches) stage_launches[i] = s->stage_launches[i];
  }
  if (rows_scored) *rows_scored = s->prof_rows;
  if (calls) *calls = s->prof_calls;
  return NANN_OK;
}

nann_status nann_searcher_get_nodes(nann_searcher_t* s, int32_t* out_nodes, int64_t cap) {
  if (!s || !out_nodes) return fail(NANN_INVALID_ARGUMENT, "null argument");
  const int64_t cnt = std::min<int64_t>((int64_t)s->last_B * s->last_k, cap);
  if (cnt > 0) NANN_CUDA(cudaMemcpy(out_nodes, s->out_nodes, (size_t)cnt * 4, cudaMemcpyDeviceToHost));
  return NANN_OK;
}

}  // extern "C"
