// lib_builder.inl -- nann_hnsw_build: the index files of build_hnsw_index.py:33-67 built on the GPU
// (kernels and the algorithm in builder_kernels.cuh).

namespace nann {

struct DevMem {           // plain cudaMalloc with RAII (multi-GB buffers: not from the small caching pool)
  void* p = nullptr;
  ~DevMem() { if (p) cudaFree(p); }
  nann_status alloc(size_t bytes) {
    if (p) { cudaFree(p); p = nullptr; }
    if (cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess) {
      cudaGetLastError();
      return fail(NANN_RESOURCE_EXHAUSTED, "OOM in the index builder (%zu bytes)", bytes);
    }
    return NANN_OK;
  }
  template <typename T> T* as() const { return (T*)p; }
};

// out[0..n] = exclusive scan of in[0..n), out[n] = total
static nann_status device_scan(const int32_t* in, int64_t n, int64_t* out, cudaStream_t st) {
  if (n <= 0) { NANN_CUDA(cudaMemsetAsync(out, 0, 8, st)); return NANN_OK; }
  const int64_t nb = ceil_div(n, SCAN_BLOCK);
  DevMem sums;
  NANN_TRY(sums.alloc((size_t)(nb + 1) * 8));
  NANN_LAUNCH(scan_block_kernel, (unsigned)nb, SCAN_BLOCK, 0, st, in, n, out, sums.as<int64_t>());
  NANN_LAUNCH(scan_sums_kernel, 1, SCAN_BLOCK, 0, st, sums.as<int64_t>(), nb);
  NANN_LAUNCH(scan_add_kernel, (unsigned)ceil_div(n + 1, SCAN_BLOCK), SCAN_BLOCK, 0, st, out, n, sums.as<int64_t>(), nb);
  NANN_CUDA(cudaStreamSynchronize(st));   // `sums` dies with this scope
  return NANN_OK;
}

static double now_s() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

// one level: members `nodes_h` (ascending global ids) -> links [s][cap] (level-local ids), out_cnt[s]
static nann_status build_level(const float* emb_d, int64_t n, const std::vector<int32_t>& nodes_h, int M, int cap, int n_sm,
                               DevMem& links, DevMem& out_cnt, DevMem& nodes_d, nann_hnsw_build_stats_t* stats, cudaStream_t st) {
  const int64_t s = (int64_t)nodes_h.size();
  const bool identity = s == n;
  NANN_TRY(nodes_d.alloc((size_t)s * 4));
  NANN_CUDA(cudaMemcpyAsync(nodes_d.p, nodes_h.data(), (size_t)s * 4, cudaMemcpyHostToDevice, st));
  DevMem Xg;
  const float* X = emb_d;
  if (!identity) {
    NANN_TRY(Xg.alloc((size_t)s * KB_D * 4));
    NANN_LAUNCH(gather_members_kernel, (unsigned)ceil_div(s * (KB_D / 4), 256), 256, 0, st, emb_d, nodes_d.as<int32_t>(), s, Xg.as<float>());
    X = Xg.as<float>();
  }
  DevMem sq, fwd, fwd_d, fwd_cnt;
  NANN_TRY(sq.alloc((size_t)s * 4));
  NANN_LAUNCH(row_sq_kernel, (unsigned)ceil_div(s, 256), 256, 0, st, X, s, sq.as<float>());
  NANN_TRY(fwd.alloc((size_t)s * M * 4));
  NANN_TRY(fwd_d.alloc((size_t)s * M * 4));
  NANN_TRY(fwd_cnt.alloc((size_t)s * 4));
  const int n_cand = std::min<int64_t>(cap + M, s - 1);
  const double t0 = now_s();

  // ---- candidates + refinement, in row ranges of at most 1M rows (the append buffers are 4 KB per row)
  const int64_t s_pad = ceil_div(s, 256) * 256;
  const bool all_pairs = s - 1 <= KB_KC;
  DevMem img, hj;
  if (!all_pairs) {
    NANN_TRY(img.alloc((size_t)s_pad / 256 * KB_BLK_BYTES));
    NANN_TRY(hj.alloc((size_t)s_pad * 4));
    NANN_LAUNCH(knn_image_kernel, (unsigned)ceil_div(s_pad * 16, 256), 256, 0, st, X, sq.as<float>(), s, s_pad, img.as<uint8_t>(), hj.as<float>());
    NANN_CUDA(cudaFuncSetAttribute(knn_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, KB_SMEM_BYTES));
  }
  NANN_CUDA(cudaFuncSetAttribute(knn_refine_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, KR_SMEM_BYTES));
  const int64_t range = std::min<int64_t>(ceil_div(s, 128) * 128, (int64_t)1 << 20);
  DevMem tau, cnt, buf, ovf;
  NANN_TRY(tau.alloc((size_t)range * 4));
  NANN_TRY(cnt.alloc((size_t)range * 4));
  NANN_TRY(buf.alloc((size_t)range * KB_CAP * 8));
  NANN_TRY(ovf.alloc(8));
  NANN_CUDA(cudaMemsetAsync(ovf.p, 0, 8, st));
  for (int64_t row0 = 0; row0 < s; row0 += range) {
    const int64_t rows = std::min<int64_t>(range, ceil_div(s - row0, 128) * 128);
    if (all_pairs) {
      NANN_LAUNCH(knn_all_pairs_kernel, (unsigned)s, 128, 0, st, cnt.as<int>(), buf.as<uint2>(), s);
    } else {
      NANN_LAUNCH(fill_words_kernel, 148 * 4, 256, 0, st, tau.as<uint32_t>(), rows, 0x7f800000u);
      NANN_LAUNCH(fill_words_kernel, 148 * 4, 256, 0, st, cnt.as<uint32_t>(), rows, 0u);
      int64_t col0 = 0;
      while (col0 < s_pad) {
        // rounds of doubling width: the first 512 columns fill the buffers (tau = inf), afterwards a round is as wide as
        // everything seen before it, so it appends ~KC pairs per row
        const int64_t width = std::min<int64_t>(col0 == 0 ? 512 : col0, s_pad - col0);
        KnnArgs a{};
        a.img = img.as<uint8_t>(); a.hj = hj.as<float>(); a.sq = sq.as<float>();
        a.tau = tau.as<float>(); a.cnt = cnt.as<int>(); a.buf = buf.as<uint2>();
        a.row0 = row0; a.n_rb = (int)(rows / 128); a.s = s; a.col0 = col0; a.n_ct = (int)(width / KB_TN);
        a.overflow = ovf.as<unsigned long long>();
        const int64_t items = (int64_t)ceil_div(a.n_ct, KB_STRIP) * a.n_rb;
        NANN_LAUNCH(knn_filter_kernel, (unsigned)std::min<int64_t>(items, n_sm), KB_THREADS, KB_SMEM_BYTES, st, a);
        NANN_LAUNCH(knn_compact_kernel, (unsigned)ceil_div(rows * 32, 256), 256, 0, st, tau.as<float>(), cnt.as<int>(), buf.as<uint2>(), rows, KB_KC);
        col0 += width;
      }
    }
    RefineArgs r{};
    r.X = X; r.sq = sq.as<float>(); r.cnt = cnt.as<int>(); r.buf = buf.as<uint2>();
    r.row0 = row0; r.n_rows = rows; r.s = s; r.n_cand = n_cand; r.M = M;
    r.fwd = fwd.as<int32_t>(); r.fwd_d = fwd_d.as<float>(); r.fwd_cnt = fwd_cnt.as<int32_t>();
    NANN_LAUNCH(knn_refine_kernel, (unsigned)std::min<int64_t>(rows, s - row0), KR_THREADS, KR_SMEM_BYTES, st, r);
  }
  NANN_CUDA(cudaStreamSynchronize(st));
  const double t1 = now_s();
  if (stats) {
    unsigned long long o = 0;
    NANN_CUDA(cudaMemcpy(&o, ovf.p, 8, cudaMemcpyDeviceToHost));
    stats->n_overflow += (int64_t)o;
    stats->seconds_knn += t1 - t0;
  }
  tau.alloc(0); cnt.alloc(0); buf.alloc(0); img.alloc(0); hj.alloc(0);

  // ---- reverse links, de-duplication, closest `cap`
  DevMem rev_cnt, rev_fill, rev_off, rev;
  NANN_TRY(rev_cnt.alloc((size_t)s * 4));
  NANN_TRY(rev_fill.alloc((size_t)s * 4));
  NANN_TRY(rev_off.alloc((size_t)(s + 1) * 8));
  NANN_CUDA(cudaMemsetAsync(rev_cnt.p, 0, (size_t)s * 4, st));
  NANN_CUDA(cudaMemsetAsync(rev_fill.p, 0, (size_t)s * 4, st));
  NANN_LAUNCH(link_count_rev_kernel, (unsigned)ceil_div(s * M, 256), 256, 0, st, fwd.as<int32_t>(), fwd_cnt.as<int32_t>(), s, M, rev_cnt.as<int>());
  NANN_TRY(device_scan(rev_cnt.as<int32_t>(), s, rev_off.as<int64_t>(), st));
  int64_t n_fwd = 0;
  NANN_CUDA(cudaMemcpy(&n_fwd, rev_off.as<int64_t>() + s, 8, cudaMemcpyDeviceToHost));
  NANN_TRY(rev.alloc((size_t)std::max<int64_t>(n_fwd, 1) * 8));
  NANN_LAUNCH(link_fill_rev_kernel, (unsigned)ceil_div(s * M, 256), 256, 0, st, fwd.as<int32_t>(), fwd_d.as<float>(), fwd_cnt.as<int32_t>(), s, M,
              rev_off.as<int64_t>(), rev_fill.as<int>(), rev.as<uint2>());
  NANN_TRY(links.alloc((size_t)s * cap * 4));
  NANN_TRY(out_cnt.alloc((size_t)s * 4));
  NANN_LAUNCH(link_finalize_kernel, (unsigned)ceil_div(s, 4), 128, 0, st, fwd.as<int32_t>(), fwd_d.as<float>(), fwd_cnt.as<int32_t>(),
              rev_off.as<int64_t>(), rev.as<uint2>(), s, M, cap, links.as<int32_t>(), out_cnt.as<int32_t>());
  NANN_CUDA(cudaStreamSynchronize(st));
  if (stats) { stats->seconds_links += now_s() - t1; stats->n_forward_links += n_fwd; }
  return NANN_OK;
}

}  // namespace nann

extern "C" nann_status nann_hnsw_build(const float* emb, int64_t n, int dim, const int32_t* levels, int M, int n_levels,
                                       int device, nann_alloc_fn alloc, void* alloc_ctx, nann_hnsw_build_stats_t* stats) {
  NANN_CUDA(cudaSetDevice(device));
  NANN_TRY(require_device());
  if (!emb || !levels || !alloc || n <= 0) return fail(NANN_INVALID_ARGUMENT, "nann_hnsw_build: null or empty input");
  if (dim != KB_D) return fail(NANN_UNIMPLEMENTED, "the index builder is built for dim=%d (got %d)", KB_D, dim);
  if (M < 2 || M > 32) return fail(NANN_UNIMPLEMENTED, "M=%d outside [2, 32]", M);
  if (n_levels < 1 || n_levels > 8) return fail(NANN_INVALID_ARGUMENT, "n_levels=%d outside [1, 8]", n_levels);
  if (n > 0x7fffffffll) return fail(NANN_UNIMPLEMENTED, "n > 2^31-1 rows per shard");
  if (stats) memset(stats, 0, sizeof(*stats));
  cudaStream_t st = 0;
  cudaDeviceProp pr;
  NANN_CUDA(cudaGetDeviceProperties(&pr, device));
  DevMem emb_own;
  const float* emb_d = emb;
  if (!is_device_ptr(emb)) {
    NANN_TRY(emb_own.alloc((size_t)n * KB_D * 4));
    NANN_CUDA(cudaMemcpy(emb_own.p, emb, (size_t)n * KB_D * 4, cudaMemcpyHostToDevice));
    emb_d = emb_own.as<float>();
  }
  std::vector<int32_t> lv(n);
  if (is_device_ptr(levels)) NANN_CUDA(cudaMemcpy(lv.data(), levels, (size_t)n * 4, cudaMemcpyDeviceToHost));
  else memcpy(lv.data(), levels, (size_t)n * 4);

  for (int l = 0; l < n_levels; ++l) {
    std::vector<int32_t> nodes;
    for (int64_t i = 0; i < n; ++i) if (lv[i] >= l) nodes.push_back((int32_t)i);
    const int64_t s = (int64_t)nodes.size();
    const int cap = l == 0 ? 2 * M : M;
    DevMem counts, rs, vals;
    NANN_TRY(counts.alloc((size_t)n * 4));
    NANN_TRY(rs.alloc((size_t)(n + 1) * 8));
    NANN_CUDA(cudaMemsetAsync(counts.p, 0, (size_t)n * 4, st));
    int64_t total = 0;
    DevMem links, out_cnt, nodes_d;
    if (s >= 2) {
      NANN_TRY(build_level(emb_d, n, nodes, M, cap, pr.multiProcessorCount, links, out_cnt, nodes_d, stats, st));
      NANN_LAUNCH(csr_counts_kernel, (unsigned)ceil_div(s, 256), 256, 0, st, nodes_d.as<int32_t>(), out_cnt.as<int32_t>(), s, counts.as<int32_t>());
    }
    NANN_TRY(device_scan(counts.as<int32_t>(), n, rs.as<int64_t>(), st));
    NANN_CUDA(cudaMemcpy(&total, rs.as<int64_t>() + n, 8, cudaMemcpyDeviceToHost));
    NANN_TRY(vals.alloc((size_t)std::max<int64_t>(total, 1) * 4));
    if (s >= 2 && total > 0)
      NANN_LAUNCH(csr_values_kernel, (unsigned)ceil_div(s * cap, 256), 256, 0, st, nodes_d.as<int32_t>(), links.as<int32_t>(), out_cnt.as<int32_t>(),
                  s, cap, rs.as<int64_t>(), vals.as<int32_t>());
    NANN_CUDA(cudaStreamSynchronize(st));
    void* o_vals = alloc(alloc_ctx, 2 * l, total);
    void* o_rs = alloc(alloc_ctx, 2 * l + 1, n + 1);
    if ((total > 0 && !o_vals) || !o_rs) return fail(NANN_RESOURCE_EXHAUSTED, "allocator returned NULL for level %d", l);
    if (total > 0)
      NANN_CUDA(cudaMemcpy(o_vals, vals.p, (size_t)total * 4, is_device_ptr(o_vals) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost));
    NANN_CUDA(cudaMemcpy(o_rs, rs.p, (size_t)(n + 1) * 8, is_device_ptr(o_rs) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost));
  }
  return NANN_OK;
}
