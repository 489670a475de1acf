// lib_ragged.inl -- the reference's ragged-batch helper ops, same C-ABI style:
//   BatchGatherOnRT  UO/beam_search_op/BatchGatherOnRT_kernel.cc:17-97   per-group gather by LOCAL offset
//   BatchConcatOnRT  UO/beam_search_op/BatchConcatOnRT_kernel.cc:18-115  per-group concatenation
//   SplitsGather     UO/beam_search_op/SplitsGather_kernel.cc:20-105     CSR range expansion
//   BitmapInit       UO/bitmap_op/bitmap_ops.cc:28-75                    bitmap of a list of ids
//   BitmapDifference UO/bitmap_op/bitmap_ops.cc:83-143                   value-semantics visited filter
// They are not emitted by build_opt_graph.py (exec.pb) -- they are the reference's route to batch>1 inside
// TF graphs -- so they are small grid-stride kernels, not tuned.

namespace nann {

template <typename T>
__global__ void batch_gather_on_rt_kernel(const T* __restrict__ pv, const int64_t* __restrict__ prs,
                                          const int64_t* __restrict__ iv, const int64_t* __restrict__ irs,
                                          int64_t n_groups, int64_t n_pv, T* __restrict__ out, int* __restrict__ bad) {
  const int64_t g = blockIdx.x;
  if (g >= n_groups) return;
  const int64_t base = prs[g], lim = prs[g + 1];
  for (int64_t j = irs[g] + threadIdx.x; j < irs[g + 1]; j += blockDim.x) {
    const int64_t idx = base + iv[j];                       // BatchGatherOnRT_kernel.cc:87-88
    if (iv[j] < 0 || idx >= lim || idx >= n_pv) { *bad = 1; continue; }
    out[j] = pv[idx];
  }
}

template <typename T>
__global__ void batch_concat_on_rt_kernel(const T* __restrict__ lv, const int64_t* __restrict__ lrs,
                                          const T* __restrict__ rv, const int64_t* __restrict__ rrs,
                                          int64_t n_groups, T* __restrict__ out, int64_t* __restrict__ out_rs) {
  const int64_t g = blockIdx.x;
  if (g >= n_groups) return;
  const int64_t lb = lrs[g], le = lrs[g + 1], rb = rrs[g], re = rrs[g + 1];
  for (int64_t i = lb + threadIdx.x; i < le; i += blockDim.x) out[rb + i] = lv[i];          // :99
  for (int64_t i = rb + threadIdx.x; i < re; i += blockDim.x) out[le + i] = rv[i];          // :100
  if (threadIdx.x == 0) { out_rs[g + 1] = le + re; if (g == 0) out_rs[0] = 0; }             // :94,:101
}

template <typename T>
__global__ void splits_gather_count_kernel(const T* __restrict__ splits, int64_t n_splits, const int64_t* __restrict__ iv,
                                           const int64_t* __restrict__ irs, int64_t n_groups,
                                           int64_t* __restrict__ len, int* __restrict__ bad) {
  const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (g >= n_groups) return;
  int64_t sum = 0;
  for (int64_t j = irs[g]; j < irs[g + 1]; ++j) {
    const int64_t idx = iv[j];
    if (idx < 0 || idx + 1 >= n_splits) { *bad = 1; continue; }
    sum += (int64_t)(splits[idx + 1] - splits[idx]);
  }
  len[g] = sum;
}
template <typename T>
__global__ void splits_gather_fill_kernel(const T* __restrict__ splits, const int64_t* __restrict__ iv,
                                          const int64_t* __restrict__ irs, int64_t n_groups,
                                          const int64_t* __restrict__ out_rs, T* __restrict__ out) {
  const int64_t g = blockIdx.x;
  if (g >= n_groups) return;
  int64_t o = out_rs[g];
  for (int64_t j = irs[g]; j < irs[g + 1]; ++j) {
    const int64_t b = (int64_t)splits[iv[j]], e = (int64_t)splits[iv[j] + 1];
    for (int64_t k = b + threadIdx.x; k < e; k += blockDim.x) out[o + (k - b)] = (T)k;      // :88-96
    o += e - b;
  }
}

template <typename T>
__global__ void bitmap_init_kernel(const T* __restrict__ idx, int64_t n, uint32_t* __restrict__ bm, int64_t n_words,
                                   int* __restrict__ bad) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const T v = idx[i];
    if (v < 0 || (int64_t)(v >> 5) >= n_words) { *bad = 1; continue; }
    atomicOr(bm + (int64_t)(v >> 5), 1u << (unsigned)(v & 31));
  }
}

template <typename T>
static nann_status batch_gather_on_rt_impl(const T* pv, int64_t n_pv, const int64_t* prs, int64_t n_prs, const int64_t* iv,
                                           int64_t n_iv, const int64_t* irs, int64_t n_irs, nann_alloc_fn alloc, void* ctx,
                                           void* stream) {
  NANN_TRY(require_device());
  if (!alloc) return fail(NANN_INVALID_ARGUMENT, "alloc callback is NULL");
  cudaStream_t st = (cudaStream_t)stream;
  DevIn<T> d_pv; DevIn<int64_t> d_prs, d_iv, d_irs;
  NANN_TRY(d_pv.init(pv, n_pv, st)); NANN_TRY(d_prs.init(prs, n_prs, st));
  NANN_TRY(d_iv.init(iv, n_iv, st)); NANN_TRY(d_irs.init(irs, n_irs, st));
  int code = 0;
  NANN_TRY(validate_ragged(n_pv, d_prs.d, n_prs, st, &code));
  if (code) return fail(NANN_INVALID_ARGUMENT, "Invalid RaggedTensor input0 params, code: %d", code);
  NANN_TRY(validate_ragged(n_iv, d_irs.d, n_irs, st, &code));
  if (code) return fail(NANN_INVALID_ARGUMENT, "Invalid RaggedTensor input1 indices, code: %d", code);
  if (n_prs == 1 || n_irs == 1) return deliver_void(alloc, ctx);                       // :59-67
  if (n_prs != n_irs)                                                                    // :69-71
    return fail(NANN_INVALID_ARGUMENT, "row_splits of two inputs do NOT match: %lld!=%lld", (long long)n_prs, (long long)n_irs);
  const int64_t G = n_irs - 1;
  DevBuf<T> d_out; DevBuf<int> d_bad;
  NANN_TRY(d_out.alloc(std::max<int64_t>(n_iv, 1))); NANN_TRY(d_bad.alloc(1));
  NANN_CUDA(cudaMemsetAsync(d_bad.d, 0, sizeof(int), st));
  NANN_LAUNCH(batch_gather_on_rt_kernel<T>, (unsigned)G, 128, 0, st, d_pv.d, d_prs.d, d_iv.d, d_irs.d, G, n_pv, d_out.d, d_bad.d);
  int bad = 0;
  NANN_CUDA(cudaMemcpyAsync(&bad, d_bad.d, sizeof(int), cudaMemcpyDeviceToHost, st));
  NANN_CUDA(cudaStreamSynchronize(st));
  if (bad) return fail(NANN_INVALID_ARGUMENT, "local offset outside its group");
  NANN_TRY(deliver<T>(alloc, ctx, 0, d_out.d, n_iv, st));
  NANN_TRY(deliver<int64_t>(alloc, ctx, 1, d_irs.d, n_irs, st));                        // set_output(1, input(3)) :79
  NANN_CUDA(cudaStreamSynchronize(st));
  return NANN_OK;
}

template <typename T>
static nann_status batch_concat_on_rt_impl(const T* lv, int64_t n_lv, const int64_t* lrs, int64_t n_lrs, const T* rv,
                                           int64_t n_rv, const int64_t* rrs, int64_t n_rrs, nann_alloc_fn alloc, void* ctx,
                                           void* stream) {
  NANN_TRY(require_device());
  if (!alloc) return fail(NANN_INVALID_ARGUMENT, "alloc callback is NULL");
  cudaStream_t st = (cudaStream_t)stream;
  DevIn<T> d_lv, d_rv; DevIn<int64_t> d_lrs, d_rrs;
  NANN_TRY(d_lv.init(lv, n_lv, st)); NANN_TRY(d_lrs.init(lrs, n_lrs, st));
  NANN_TRY(d_rv.init(rv, n_rv, st)); NANN_TRY(d_rrs.init(rrs, n_rrs, st));
  int code = 0;
  NANN_TRY(validate_ragged(n_lv, d_lrs.d, n_lrs, st, &code));
  if (code) return fail(NANN_INVALID_ARGUMENT, "Invalid RaggedTensor input0 left, code: %d", code);
  NANN_TRY(validate_ragged(n_rv, d_rrs.d, n_rrs, st, &code));
  if (code) return fail(NANN_INVALID_ARGUMENT, "Invalid RaggedTensor input1 right, code: %d", code);
  if (n_lrs == 1) {                                                                      // void left -> right (:62-67)
    NANN_TRY(deliver<T>(alloc, ctx, 0, d_rv.d, n_rv, st));
    NANN_TRY(deliver<int64_t>(alloc, ctx, 1, d_rrs.d, n_rrs, st));
    NANN_CUDA(cudaStreamSynchronize(st));
    return NANN_OK;
  }
  if (n_rrs == 1) {                                                                      // void right -> left (:68-73)
    NANN_TRY(deliver<T>(alloc, ctx, 0, d_lv.d, n_lv, st));
    NANN_TRY(deliver<int64_t>(alloc, ctx, 1, d_lrs.d, n_lrs, st));
    NANN_CUDA(cudaStreamSynchronize(st));
    return NANN_OK;
  }
  if (n_lrs != n_rrs)
    return fail(NANN_INVALID_ARGUMENT, "row_splits of two inputs do NOT match: %lld!=%lld", (long long)n_lrs, (long long)n_rrs);
  const int64_t G = n_lrs - 1;
  DevBuf<T> d_out; DevBuf<int64_t> d_ors;
  NANN_TRY(d_out.alloc(std::max<int64_t>(n_lv + n_rv, 1))); NANN_TRY(d_ors.alloc(n_lrs));
  NANN_LAUNCH(batch_concat_on_rt_kernel<T>, (unsigned)G, 128, 0, st, d_lv.d, d_lrs.d, d_rv.d, d_rrs.d, G, d_out.d, d_ors.d);
  NANN_TRY(deliver<T>(alloc, ctx, 0, d_out.d, n_lv + n_rv, st));
  NANN_TRY(deliver<int64_t>(alloc, ctx, 1, d_ors.d, n_lrs, st));
  NANN_CUDA(cudaStreamSynchronize(st));
  return NANN_OK;
}

template <typename T>
static nann_status splits_gather_impl(const T* splits, int64_t n_splits, const int64_t* iv, int64_t n_iv, const int64_t* irs,
                                      int64_t n_irs, nann_alloc_fn alloc, void* ctx, void* stream) {
  NANN_TRY(require_device());
  if (!alloc) return fail(NANN_INVALID_ARGUMENT, "alloc callback is NULL");
  cudaStream_t st = (cudaStream_t)stream;
  DevIn<T> d_sp; DevIn<int64_t> d_iv, d_irs;
  NANN_TRY(d_sp.init(splits, n_splits, st)); NANN_TRY(d_iv.init(iv, n_iv, st)); NANN_TRY(d_irs.init(irs, n_irs, st));
  int code = 0;
  NANN_TRY(validate_ragged(n_iv, d_irs.d, n_irs, st, &code));
  if (code) return fail(NANN_INVALID_ARGUMENT, "Invalid RaggedTensor input1 indices, code: %d", code);
  if (n_splits > 0) {                                                                    // :45-46
    T first;
    NANN_CUDA(cudaMemcpyAsync(&first, d_sp.d, sizeof(T), cudaMemcpyDeviceToHost, st));
    NANN_CUDA(cudaStreamSynchronize(st));
    if (first != 0) return fail(NANN_INVALID_ARGUMENT, "input splits should NOT contain less than ONE element.");
  }
  if (n_splits <= 1 || n_irs == 1) return deliver_void(alloc, ctx);                    // :47-55
  const int64_t G = n_irs - 1;
  DevBuf<int64_t> d_len, d_rs; DevBuf<int> d_bad;
  NANN_TRY(d_len.alloc(G)); NANN_TRY(d_rs.alloc(G + 1)); NANN_TRY(d_bad.alloc(1));
  NANN_CUDA(cudaMemsetAsync(d_bad.d, 0, sizeof(int), st));
  NANN_LAUNCH(splits_gather_count_kernel<T>, (unsigned)ceil_div(G, 128), 128, 0, st, d_sp.d, n_splits, d_iv.d, d_irs.d, G, d_len.d, d_bad.d);
  std::vector<int64_t> len(G), rs(G + 1);
  int bad = 0;
  NANN_CUDA(cudaMemcpyAsync(len.data(), d_len.d, G * 8, cudaMemcpyDeviceToHost, st));
  NANN_CUDA(cudaMemcpyAsync(&bad, d_bad.d, sizeof(int), cudaMemcpyDeviceToHost, st));
  NANN_CUDA(cudaStreamSynchronize(st));
  if (bad) return fail(NANN_INVALID_ARGUMENT, "indices_values out of range of splits");
  rs[0] = 0;
  for (int64_t g = 0; g < G; ++g) rs[g + 1] = rs[g] + len[g];
  NANN_CUDA(cudaMemcpyAsync(d_rs.d, rs.data(), (G + 1) * 8, cudaMemcpyHostToDevice, st));
  DevBuf<T> d_out;
  NANN_TRY(d_out.alloc(std::max<int64_t>(rs[G], 1)));
  if (rs[G] > 0) NANN_LAUNCH(splits_gather_fill_kernel<T>, (unsigned)G, 128, 0, st, d_sp.d, d_iv.d, d_irs.d, G, d_rs.d, d_out.d);
  NANN_TRY(deliver<T>(alloc, ctx, 0, d_out.d, rs[G], st));
  NANN_TRY(deliver<int64_t>(alloc, ctx, 1, d_rs.d, G + 1, st));
  NANN_CUDA(cudaStreamSynchronize(st));
  return NANN_OK;
}

template <typename T>
static nann_status bitmap_init_impl(const T* idx, int64_t n, int32_t length, int32_t* bitmap, void* stream) {
  NANN_TRY(require_device());
  if (length < 0 || n > length)                                                          // bitmap_ops.cc:54-55
    return fail(NANN_INVALID_ARGUMENT, "require: length >= idx.size() and length >=0 but length:%d idx.size():%lld", length, (long long)n);
  cudaStream_t st = (cudaStream_t)stream;
  DevIn<T> d_idx;
  DevOut<int32_t> d_bm;
  NANN_TRY(d_idx.init(idx, n, st));
  NANN_TRY(d_bm.init(bitmap, length, st, false));
  if (length > 0) NANN_CUDA(cudaMemsetAsync(d_bm.d, 0, (size_t)length * 4, st));        // :60
  DevBuf<int> d_bad;
  NANN_TRY(d_bad.alloc(1));
  NANN_CUDA(cudaMemsetAsync(d_bad.d, 0, sizeof(int), st));
  if (n > 0) NANN_LAUNCH(bitmap_init_kernel<T>, (unsigned)std::min<int64_t>(ceil_div(n, 256), 1184), 256, 0, st, d_idx.d, n, (uint32_t*)d_bm.d, (int64_t)length, d_bad.d);
  int bad = 0;
  NANN_CUDA(cudaMemcpyAsync(&bad, d_bad.d, sizeof(int), cudaMemcpyDeviceToHost, st));
  NANN_TRY(d_bm.finish(st));
  NANN_CUDA(cudaStreamSynchronize(st));
  if (bad) return fail(NANN_INVALID_ARGUMENT, "node id outside the bitmap (>= 32*%d or negative)", length);
  return NANN_OK;
}

}  // namespace nann

extern "C" {
#define NANN_RAGGED_WRAP(T, SFX)                                                                                          \
  nann_status nann_batch_gather_on_rt_##SFX(const T* pv, int64_t n_pv, const int64_t* prs, int64_t n_prs, const int64_t* iv, \
                                            int64_t n_iv, const int64_t* irs, int64_t n_irs, nann_alloc_fn alloc, void* ctx, \
                                            void* stream) {                                                                \
    return batch_gather_on_rt_impl<T>(pv, n_pv, prs, n_prs, iv, n_iv, irs, n_irs, alloc, ctx, stream);                    \
  }                                                                                                                        \
  nann_status nann_batch_concat_on_rt_##SFX(const T* lv, int64_t n_lv, const int64_t* lrs, int64_t n_lrs, const T* rv,     \
                                            int64_t n_rv, const int64_t* rrs, int64_t n_rrs, nann_alloc_fn alloc, void* ctx, \
                                            void* stream) {                                                                \
    return batch_concat_on_rt_impl<T>(lv, n_lv, lrs, n_lrs, rv, n_rv, rrs, n_rrs, alloc, ctx, stream);                    \
  }                                                                                                                        \
  nann_status nann_splits_gather_##SFX(const T* splits, int64_t n_splits, const int64_t* iv, int64_t n_iv,                 \
                                       const int64_t* irs, int64_t n_irs, nann_alloc_fn alloc, void* ctx, void* stream) { \
    return splits_gather_impl<T>(splits, n_splits, iv, n_iv, irs, n_irs, alloc, ctx, stream);                             \
  }                                                                                                                        \
  nann_status nann_bitmap_init_##SFX(const T* idx, int64_t n, int32_t length, int32_t* bitmap, void* stream) {             \
    return bitmap_init_impl<T>(idx, n, length, bitmap, stream);                                                           \
  }                                                                                                                        \
  /* BitmapDifference (bitmap_ops.cc:83-143): copy idx_flag, then the same ordered test-and-set; outputs 0 = */           \
  /* idx_next_new (T), idx_flag_new written to the caller's buffer */                                                     \
  nann_status nann_bitmap_difference_##SFX(const T* idx_next, int64_t n, const int32_t* idx_flag, int64_t n_flags,         \
                                           int32_t* idx_flag_new, nann_alloc_fn alloc, void* ctx, void* stream) {          \
    NANN_TRY(require_device());                                                                                            \
    if (!alloc || !idx_flag_new) return fail(NANN_INVALID_ARGUMENT, "null argument");                                      \
    cudaStream_t st = (cudaStream_t)stream;                                                                                \
    DevOut<int32_t> d_new;                                                                                                 \
    NANN_TRY(d_new.init(idx_flag_new, n_flags, st, false));                                                                \
    if (n_flags > 0)                                                                                                       \
      NANN_CUDA(cudaMemcpyAsync(d_new.d, idx_flag, (size_t)n_flags * 4,                                                    \
                                is_device_ptr(idx_flag) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));        \
    struct Fwd { nann_alloc_fn alloc; void* ctx; } fwd{alloc, ctx};                                                        \
    auto only_values = [](void* c, int idx, int64_t cnt) -> void* {                                                        \
      static thread_local int64_t sink[2];                                                                                 \
      auto* f = (Fwd*)c;                                                                                                   \
      return idx == 0 ? f->alloc(f->ctx, 0, cnt) : (void*)sink;  /* the op has no row_splits output */                     \
    };                                                                                                                     \
    const int64_t rs[2] = {0, n};                                                                                          \
    nann_status rc = bitmap_diff_impl<T>(idx_next, n, rs, 2, d_new.d, n_flags, only_values, &fwd, stream);                 \
    if (rc != NANN_OK) return rc;                                                                                          \
    NANN_TRY(d_new.finish(st));                                                                                            \
    NANN_CUDA(cudaStreamSynchronize(st));                                                                                  \
    return NANN_OK;                                                                                                        \
  }
NANN_RAGGED_WRAP(int32_t, i32)
NANN_RAGGED_WRAP(int64_t, i64)
#undef NANN_RAGGED_WRAP
}  // extern "C"
