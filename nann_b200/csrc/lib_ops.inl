// lib_ops.inl -- op-level entry points (one call = one OpKernel::Compute of the reference).
// Host pointers are staged over the device; the kernels are in traverse_kernels.cuh.

namespace nann {

// ragged validation, GroupGather_kernel.cc:9-16; row_splits may live on the device
static nann_status validate_ragged(int64_t n_values, const int64_t* rs_dev, int64_t n_rs,
                                   cudaStream_t st, int* code) {
  *code = 0;
  if (n_rs == 0) { *code = 1; return NANN_OK; }
  int64_t ends[2] = {0, 0};
  NANN_CUDA(cudaMemcpyAsync(&ends[0], rs_dev, 8, cudaMemcpyDeviceToHost, st));
  NANN_CUDA(cudaMemcpyAsync(&ends[1], rs_dev + (n_rs - 1), 8, cudaMemcpyDeviceToHost, st));
  NANN_CUDA(cudaStreamSynchronize(st));
  if (ends[0] != 0) *code = 2;
  else if (ends[1] != n_values) *code = 3;
  return NANN_OK;
}

// copy a device result into whatever the allocator returned (host or device)
template <typename T>
static nann_status deliver(nann_alloc_fn alloc, void* ctx, int idx, const T* dev_src, int64_t n,
                           cudaStream_t st) {
  T* dst = (T*)alloc(ctx, idx, n);
  if (n == 0) return NANN_OK;
  if (!dst) return fail(NANN_RESOURCE_EXHAUSTED, "allocator returned NULL for output %d", idx);
  NANN_CUDA(cudaMemcpyAsync(dst, dev_src, (size_t)n * sizeof(T),
                            is_device_ptr(dst) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
  return NANN_OK;
}

static nann_status deliver_void(nann_alloc_fn alloc, void* ctx) {  // values=[], row_splits=[0]
  alloc(ctx, 0, 0);
  int64_t* rs = (int64_t*)alloc(ctx, 1, 1);
  if (!rs) return fail(NANN_RESOURCE_EXHAUSTED, "allocator returned NULL for output 1");
  const int64_t zero = 0;
  if (is_device_ptr(rs)) NANN_CUDA(cudaMemcpy(rs, &zero, 8, cudaMemcpyHostToDevice));
  else *rs = 0;
  return NANN_OK;
}

// exclusive scan of small int64 arrays on the host side of the op (group counts are tiny:
// one entry per ragged group); keeps the op-level path simple.
template <typename T>
static nann_status group_gather_impl(const T* pv, int64_t n_pv, const int64_t* prs, int64_t n_prs,
                                     const int64_t* iv, int64_t n_iv, const int64_t* irs,
                                     int64_t n_irs, int unique, nann_alloc_fn alloc, void* ctx,
                                     void* stream) {
  NANN_TRY(require_device());
  if (!alloc) return fail(NANN_INVALID_ARGUMENT, "alloc callback is NULL");
  cudaStream_t st = (cudaStream_t)stream;
  DevIn<T> d_pv;
  DevIn<int64_t> d_prs, d_iv, d_irs;
  NANN_TRY(d_pv.init(pv, n_pv, st));
  NANN_TRY(d_prs.init(prs, n_prs, st));
  NANN_TRY(d_iv.init(iv, n_iv, st));
  NANN_TRY(d_irs.init(irs, n_irs, st));
  int code = 0;
  NANN_TRY(validate_ragged(n_pv, d_prs.d, n_prs, st, &code));
  if (code) return fail(NANN_INVALID_ARGUMENT, "Invalid RaggedTensor input0 params, code: %d", code);
  NANN_TRY(validate_ragged(n_iv, d_irs.d, n_irs, st, &code));
  if (code) return fail(NANN_INVALID_ARGUMENT, "Invalid RaggedTensor input1 indices, code: %d", code);
  if (n_prs == 1 || n_irs == 1) return deliver_void(alloc, ctx);  // GroupGather_kernel.cc:69-77

  const int64_t G = n_irs - 1;
  DevBuf<int64_t> d_len, d_rs;
  DevBuf<int> d_bad;
  NANN_TRY(d_len.alloc(G));
  NANN_TRY(d_rs.alloc(G + 1));
  NANN_TRY(d_bad.alloc(1));
  NANN_CUDA(cudaMemsetAsync(d_bad.d, 0, sizeof(int), st));
  const int threads = 128;
  const int blocks = (int)ceil_div(G * 32, threads);
  NANN_LAUNCH(group_gather_count_kernel<T>, blocks, threads, 0, st, d_prs.d, n_prs, d_iv.d, d_irs.d, G,
              d_len.d, d_bad.d);
  std::vector<int64_t> len(G), rs(G + 1);
  int bad = 0;
  NANN_CUDA(cudaMemcpyAsync(len.data(), d_len.d, G * 8, cudaMemcpyDeviceToHost, st));
  NANN_CUDA(cudaMemcpyAsync(&bad, d_bad.d, sizeof(int), cudaMemcpyDeviceToHost, st));
  NANN_CUDA(cudaStreamSynchronize(st));
  if (bad) return fail(NANN_INVALID_ARGUMENT, "indices_values out of range of params_row_splits");
  rs[0] = 0;
  for (int64_t g = 0; g < G; ++g) rs[g + 1] = rs[g] + len[g];
  NANN_CUDA(cudaMemcpyAsync(d_rs.d, rs.data(), (G + 1) * 8, cudaMemcpyHostToDevice, st));
  DevBuf<T> d_out;
  NANN_TRY(d_out.alloc(rs[G]));
  if (!unique) {
    if (rs[G] > 0)
      NANN_LAUNCH(group_gather_fill_kernel<T>, blocks, threads, 0, st, d_pv.d, d_prs.d, d_iv.d, d_irs.d, G,
                  d_rs.d, d_out.d);
    NANN_TRY(deliver<T>(alloc, ctx, 0, d_out.d, rs[G], st));
    NANN_TRY(deliver<int64_t>(alloc, ctx, 1, d_rs.d, G + 1, st));
    NANN_CUDA(cudaStreamSynchronize(st));
    return NANN_OK;
  }
  // unique: dedup into the capacity layout, then compact group by group
  DevBuf<int64_t> d_ulen;
  NANN_TRY(d_ulen.alloc(G));
  NANN_LAUNCH(group_gather_unique_kernel<T>, blocks, threads, 0, st, d_pv.d, d_prs.d, d_iv.d, d_irs.d, G,
              d_rs.d, d_out.d, d_ulen.d);
  std::vector<int64_t> ulen(G), urs(G + 1);
  NANN_CUDA(cudaMemcpyAsync(ulen.data(), d_ulen.d, G * 8, cudaMemcpyDeviceToHost, st));
  NANN_CUDA(cudaStreamSynchronize(st));
  urs[0] = 0;
  for (int64_t g = 0; g < G; ++g) urs[g + 1] = urs[g] + ulen[g];
  T* dst = (T*)alloc(ctx, 0, urs[G]);
  if (urs[G] > 0 && !dst) return fail(NANN_RESOURCE_EXHAUSTED, "allocator returned NULL for output 0");
  const cudaMemcpyKind kind = is_device_ptr(dst) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
  for (int64_t g = 0; g < G; ++g)
    if (ulen[g] > 0)
      NANN_CUDA(cudaMemcpyAsync(dst + urs[g], d_out.d + rs[g], (size_t)ulen[g] * sizeof(T), kind, st));
  int64_t* rs_out = (int64_t*)alloc(ctx, 1, G + 1);
  if (!rs_out) return fail(NANN_RESOURCE_EXHAUSTED, "allocator returned NULL for output 1");
  NANN_CUDA(cudaMemcpyAsync(rs_out, urs.data(), (G + 1) * 8,
                            is_device_ptr(rs_out) ? cudaMemcpyHostToDevice : cudaMemcpyHostToHost, st));
  NANN_CUDA(cudaStreamSynchronize(st));
  return NANN_OK;
}

template <typename T>
static nann_status bitmap_diff_impl(const T* vals, int64_t n_v, const int64_t* rs, int64_t n_rs,
                                    int32_t* flags, int64_t n_flags, nann_alloc_fn alloc, void* ctx,
                                    void* stream) {
  NANN_TRY(require_device());
  if (!alloc) return fail(NANN_INVALID_ARGUMENT, "alloc callback is NULL");
  cudaStream_t st = (cudaStream_t)stream;
  DevIn<T> d_v;
  DevIn<int64_t> d_rs;
  NANN_TRY(d_v.init(vals, n_v, st));
  NANN_TRY(d_rs.init(rs, n_rs, st));
  int code = 0;
  NANN_TRY(validate_ragged(n_v, d_rs.d, n_rs, st, &code));
  if (code) return fail(NANN_INVALID_ARGUMENT, "Invalid RaggedTensor input0 a, code: %d", code);
  if (n_rs == 1) return deliver_void(alloc, ctx);  // bitmap_ops.cc:187-196
  const int64_t G = n_rs - 1;
  DevOut<int32_t> d_flags;
  NANN_TRY(d_flags.init(flags, n_flags, st, /*copy_in=*/true));
  DevBuf<T> d_out;
  DevBuf<int64_t> d_ors;
  DevBuf<int> d_bad;
  NANN_TRY(d_out.alloc(n_v > 0 ? n_v : 1));
  NANN_TRY(d_ors.alloc(G + 1));
  NANN_TRY(d_bad.alloc(1));
  NANN_CUDA(cudaMemsetAsync(d_bad.d, 0, sizeof(int), st));
  NANN_LAUNCH(bitmap_diff_op_kernel<T>, 1, 32, 0, st, d_v.d, d_rs.d, G, (uint32_t*)d_flags.d, n_flags,
              d_out.d, d_ors.d, d_bad.d);
  int bad = 0;
  int64_t total = 0;
  NANN_CUDA(cudaMemcpyAsync(&bad, d_bad.d, sizeof(int), cudaMemcpyDeviceToHost, st));
  NANN_CUDA(cudaMemcpyAsync(&total, d_ors.d + G, 8, cudaMemcpyDeviceToHost, st));
  NANN_CUDA(cudaStreamSynchronize(st));
  if (bad) return fail(NANN_INVALID_ARGUMENT, "node id outside the bitmap (>= 32*%lld or negative)", (long long)n_flags);
  NANN_TRY(deliver<T>(alloc, ctx, 0, d_out.d, total, st));
  NANN_TRY(deliver<int64_t>(alloc, ctx, 1, d_ors.d, G + 1, st));
  NANN_TRY(d_flags.finish(st));
  NANN_CUDA(cudaStreamSynchronize(st));
  return NANN_OK;
}

static size_t topk_smem_bytes(int k) {
  int kpad = 1;
  while (kpad < k) kpad <<= 1;
  return (size_t)kpad * 8;
}

static nann_status launch_topk(const TopkArgs& a, int64_t rows, cudaStream_t st) {
  if (a.k > TOPK_MAX_K) return fail(NANN_UNIMPLEMENTED, "k=%d > %d", a.k, TOPK_MAX_K);
  if (rows <= 0) return NANN_OK;
  NANN_LAUNCH(topk_kernel, (unsigned)rows, TOPK_THREADS, topk_smem_bytes(a.k > 0 ? a.k : 1), st, a);
  return NANN_OK;
}

}  // namespace nann

extern "C" {

nann_status nann_group_gather_i32(const int32_t* pv, int64_t n_pv, const int64_t* prs, int64_t n_prs,
                                  const int64_t* iv, int64_t n_iv, const int64_t* irs, int64_t n_irs,
                                  int unique, nann_alloc_fn alloc, void* ctx, void* stream) {
  return group_gather_impl<int32_t>(pv, n_pv, prs, n_prs, iv, n_iv, irs, n_irs, unique, alloc, ctx, stream);
}
nann_status nann_group_gather_i64(const int64_t* pv, int64_t n_pv, const int64_t* prs, int64_t n_prs,
                                  const int64_t* iv, int64_t n_iv, const int64_t* irs, int64_t n_irs,
                                  int unique, nann_alloc_fn alloc, void* ctx, void* stream) {
  return group_gather_impl<int64_t>(pv, n_pv, prs, n_prs, iv, n_iv, irs, n_irs, unique, alloc, ctx, stream);
}
nann_status nann_bitmap_ref_difference_i32(const int32_t* v, int64_t n_v, const int64_t* rs, int64_t n_rs,
                                           int32_t* flags, int64_t n_flags, nann_alloc_fn alloc,
                                           void* ctx, void* stream) {
  return bitmap_diff_impl<int32_t>(v, n_v, rs, n_rs, flags, n_flags, alloc, ctx, stream);
}
nann_status nann_bitmap_ref_difference_i64(const int64_t* v, int64_t n_v, const int64_t* rs, int64_t n_rs,
                                           int32_t* flags, int64_t n_flags, nann_alloc_fn alloc,
                                           void* ctx, void* stream) {
  return bitmap_diff_impl<int64_t>(v, n_v, rs, n_rs, flags, n_flags, alloc, ctx, stream);
}

nann_status nann_topk_v2_f32(const float* input, int64_t rows, int64_t cols, int32_t k, int sorted,
                             float* values, int32_t* indices, void* stream) {
  (void)sorted;
  NANN_TRY(require_device());
  if (k < 0) return fail(NANN_INVALID_ARGUMENT, "Need k >= 0, got %d", k);                 // topk_op.cc:60-61
  if (cols < k)                                                                              // :66-69
    return fail(NANN_INVALID_ARGUMENT, "input must have at least k columns. Had %lld, needed %d",
                (long long)cols, k);
  if (k == 0 || rows == 0) return NANN_OK;                                                   // :84-85
  if (cols > 0x7fffffff) return fail(NANN_UNIMPLEMENTED, "cols > 2^31-1");
  cudaStream_t st = (cudaStream_t)stream;
  DevIn<float> d_in;
  DevOut<float> d_val;
  DevOut<int32_t> d_idx;
  NANN_TRY(d_in.init(input, rows * cols, st));
  NANN_TRY(d_val.init(values, rows * k, st, false));
  NANN_TRY(d_idx.init(indices, rows * k, st, false));
  TopkArgs a{};
  a.a_n = 0;
  a.b_sc = d_in.d; a.b_sc_stride = cols; a.b_n_fixed = (int)cols;
  a.k = k;
  a.out_sc = d_val.d; a.out_pos = d_idx.d; a.out_stride = k;
  NANN_TRY(launch_topk(a, rows, st));
  NANN_TRY(d_val.finish(st));
  NANN_TRY(d_idx.finish(st));
  if (d_val.staged() || d_idx.staged() || d_in.owned) NANN_CUDA(cudaStreamSynchronize(st));
  return NANN_OK;
}

nann_status nann_batch_topk_on_rt_f32(const float* values_in, int64_t n_values, const int64_t* row_splits_in,
                                      int64_t n_row_splits, const int64_t* k, int64_t n_k, int ascending,
                                      nann_alloc_fn alloc, void* ctx, void* stream) {
  NANN_TRY(require_device());
  if (!alloc || !k) return fail(NANN_INVALID_ARGUMENT, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  DevIn<float> d_v;
  DevIn<int64_t> d_rs;
  NANN_TRY(d_v.init(values_in, n_values, st));
  NANN_TRY(d_rs.init(row_splits_in, n_row_splits, st));
  int code = 0;
  NANN_TRY(validate_ragged(n_values, d_rs.d, n_row_splits, st, &code));
  if (code) return fail(NANN_INVALID_ARGUMENT, "Invalid RaggedTensor input, code: %d", code);
  const int64_t G = n_row_splits - 1;
  if (G == 0) {                                                    // BatchTopKOnRT_kernel.cc:88-96
    alloc(ctx, 0, 0); alloc(ctx, 1, 0);
    int64_t* rs = (int64_t*)alloc(ctx, 2, 1);
    if (!rs) return fail(NANN_RESOURCE_EXHAUSTED, "allocator returned NULL for output 2");
    const int64_t zero = 0;
    if (is_device_ptr(rs)) NANN_CUDA(cudaMemcpy(rs, &zero, 8, cudaMemcpyHostToDevice)); else *rs = 0;
    return NANN_OK;
  }
  if (n_k != 1 && n_k != G)                                        // :103-105
    return fail(NANN_INVALID_ARGUMENT, "Size of k vector does NOT match number of groups: %lld!=%lld", (long long)n_k, (long long)G);
  std::vector<int64_t> rs(n_row_splits), kk(G), ors(G + 1);
  NANN_CUDA(cudaMemcpyAsync(rs.data(), d_rs.d, n_row_splits * 8, cudaMemcpyDeviceToHost, st));
  std::vector<int64_t> kh(n_k);
  NANN_CUDA(cudaMemcpyAsync(kh.data(), k, n_k * 8, is_device_ptr(k) ? cudaMemcpyDeviceToHost : cudaMemcpyHostToHost, st));
  NANN_CUDA(cudaStreamSynchronize(st));
  ors[0] = 0;
  int64_t kmax = 0;
  for (int64_t g = 0; g < G; ++g) {
    const int64_t len = rs[g + 1] - rs[g];
    if (len < 0 || len > 0x7fffffff) return fail(NANN_INVALID_ARGUMENT, "Invalid RaggedTensor input: group %lld", (long long)g);
    kk[g] = std::max<int64_t>(0, std::min(len, n_k == 1 ? kh[0] : kh[g]));   // :117-121
    kmax = std::max(kmax, kk[g]);
    ors[g + 1] = ors[g] + kk[g];
  }
  if (kmax > TOPK_MAX_K) return fail(NANN_UNIMPLEMENTED, "k=%lld > %d", (long long)kmax, TOPK_MAX_K);
  const int64_t total = ors[G];
  DevBuf<int64_t> d_k, d_ors, d_idx;
  DevBuf<float> d_out;
  NANN_TRY(d_k.alloc(G)); NANN_TRY(d_ors.alloc(G + 1)); NANN_TRY(d_idx.alloc(std::max<int64_t>(total, 1)));
  NANN_TRY(d_out.alloc(std::max<int64_t>(total, 1)));
  NANN_CUDA(cudaMemcpyAsync(d_k.d, kk.data(), G * 8, cudaMemcpyHostToDevice, st));
  NANN_CUDA(cudaMemcpyAsync(d_ors.d, ors.data(), (G + 1) * 8, cudaMemcpyHostToDevice, st));
  if (total > 0) {
    TopkArgs a{};
    a.b_sc = d_v.d; a.b_row_off = d_rs.d; a.k_ptr = d_k.d; a.k = (int)kmax; a.out_row_off = d_ors.d;
    a.out_sc = d_out.d; a.out_pos64 = d_idx.d; a.ascending = ascending ? 1 : 0;
    NANN_LAUNCH(topk_kernel, (unsigned)G, TOPK_THREADS, topk_smem_bytes((int)std::max<int64_t>(kmax, 1)), st, a);
  }
  NANN_TRY(deliver<float>(alloc, ctx, 0, d_out.d, total, st));
  NANN_TRY(deliver<int64_t>(alloc, ctx, 1, d_idx.d, total, st));
  NANN_TRY(deliver<int64_t>(alloc, ctx, 2, d_ors.d, G + 1, st));
  NANN_CUDA(cudaStreamSynchronize(st));
  return NANN_OK;
}

nann_status nann_gather_rows(const void* table, int64_t n_rows, int64_t row_bytes, const int32_t* ids,
                             int64_t n, void* out, void* stream) {
  NANN_TRY(require_device());
  if (n <= 0) return NANN_OK;
  if (row_bytes <= 0) return fail(NANN_INVALID_ARGUMENT, "row_bytes=%lld", (long long)row_bytes);
  cudaStream_t st = (cudaStream_t)stream;
  DevIn<uint8_t> d_tab;
  DevIn<int32_t> d_ids;
  DevOut<uint8_t> d_out;
  NANN_TRY(d_tab.init((const uint8_t*)table, n_rows * row_bytes, st));
  NANN_TRY(d_ids.init(ids, n, st));
  NANN_TRY(d_out.init((uint8_t*)out, n * row_bytes, st, false));
  DevBuf<int> d_bad;
  NANN_TRY(d_bad.alloc(1));
  NANN_CUDA(cudaMemsetAsync(d_bad.d, 0, sizeof(int), st));
  const bool vec = (row_bytes % 16 == 0) && (((uintptr_t)d_tab.d | (uintptr_t)d_out.d) % 16 == 0);
  const int threads = 256;
  if (vec) {
    const int vpr = (int)(row_bytes / 16);
    const int64_t total = n * vpr;
    const int blocks = (int)std::min<int64_t>(ceil_div(total, threads), 148 * 16);
    NANN_LAUNCH(gather_rows_vec_kernel, blocks, threads, 0, st, (const uint4*)d_tab.d, n_rows, vpr, d_ids.d,
                n, (uint4*)d_out.d, d_bad.d);
  } else {
    const int64_t total = n * row_bytes;
    const int blocks = (int)std::min<int64_t>(ceil_div(total, threads), 148 * 16);
    NANN_LAUNCH(gather_rows_bytes_kernel, blocks, threads, 0, st, d_tab.d, n_rows, row_bytes, d_ids.d, n,
                d_out.d, d_bad.d);
  }
  int bad = 0;
  NANN_CUDA(cudaMemcpyAsync(&bad, d_bad.d, sizeof(int), cudaMemcpyDeviceToHost, st));
  NANN_TRY(d_out.finish(st));
  NANN_CUDA(cudaStreamSynchronize(st));
  if (bad) return fail(NANN_INVALID_ARGUMENT, "indices out of range [0, %lld)", (long long)n_rows);
  return NANN_OK;
}

nann_status nann_merge_topk(const float* scores, const int64_t* ids, int G, int B, int k_in, int k_out,
                            float* out_scores, int64_t* out_ids, void* stream) {
  NANN_TRY(require_device());
  if (G <= 0 || B < 0 || k_in < 0 || k_out < 0) return fail(NANN_INVALID_ARGUMENT, "bad merge shape");
  if ((int64_t)G * k_in < k_out)
    return fail(NANN_INVALID_ARGUMENT, "input must have at least k columns. Had %lld, needed %d",
                (long long)G * k_in, k_out);
  if (B == 0 || k_out == 0) return NANN_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t tot = (int64_t)G * B * k_in;
  DevIn<float> d_sc;
  DevIn<int64_t> d_ids;
  DevOut<float> d_osc;
  DevOut<int64_t> d_oid;
  NANN_TRY(d_sc.init(scores, tot, st));
  NANN_TRY(d_ids.init(ids, tot, st));
  NANN_TRY(d_osc.init(out_scores, (int64_t)B * k_out, st, false));
  NANN_TRY(d_oid.init(out_ids, (int64_t)B * k_out, st, false));
  DevBuf<float> packed;
  DevBuf<int32_t> pos;
  NANN_TRY(packed.alloc(tot));
  NANN_TRY(pos.alloc((int64_t)B * k_out));
  NANN_LAUNCH(merge_pack_kernel, (unsigned)ceil_div(tot, 256), 256, 0, st, d_sc.d, G, B, k_in, packed.d);
  TopkArgs a{};
  a.b_sc = packed.d; a.b_sc_stride = (int64_t)G * k_in; a.b_n_fixed = G * k_in;
  a.k = k_out;
  a.out_sc = d_osc.d; a.out_pos = pos.d; a.out_stride = k_out;
  NANN_TRY(launch_topk(a, B, st));
  NANN_LAUNCH(merge_emit_kernel, (unsigned)ceil_div((int64_t)B * k_out, 256), 256, 0, st, pos.d, d_ids.d, G, B,
              k_in, k_out, d_oid.d);
  NANN_TRY(d_osc.finish(st));
  NANN_TRY(d_oid.finish(st));
  NANN_CUDA(cudaStreamSynchronize(st));  // temporaries die with this scope
  return NANN_OK;
}

}  // extern "C"
