// lib_eval.inl -- the `main.py --job-type test` traversal (NANN_impls/nann/model/model.py:299-362)
// for a batch of queries, behind the C ABI.
//
// Differences from the exec.pb traversal (lib_search.inl): the candidates of a round are the UNIQUE unvisited
// neighbours of the frontier in ASCENDING id order (tf.unique + tf.sets.set_difference, :319-322), the running
// result is merged with the new scores under a clamped k (:268, :329-331), the next frontier is the new nodes
// whose score reaches the worst kept score (:333-334), and the visited set of a level starts as the level's entry
// results (:312).  "Unique, minus visited, ascending" is exactly what a bitmap gives for free:
//     eval_expand_kernel   every neighbour whose visited bit is clear sets its bit in a second ("cand") bitmap
//                          -- order-free, so the whole frontier expands in parallel (atomicOr)
//     eval_compact_kernel  one CTA per query scans the cand words in ascending order: popcount + block scan ->
//                          ascending id list; visited |= cand; cand = 0
//     scorer               the same fused gather + scorer kernels as the exec.pb path
//     topk_kernel          (R ++ Nx) -> clamped top-k, into the other half of a ping-pong result buffer
//     eval_frontier_kernel order-preserving compaction of Nx by score >= cut

namespace nann {

// ---- mark list[q][0 .. n[q]) in bitmap[q]
__global__ void eval_mark_kernel(const int32_t* __restrict__ list, int64_t stride, const int32_t* __restrict__ n_ptr, int cap,
                                 uint32_t* __restrict__ bitmap, int64_t n_words, const int32_t* __restrict__ status) {
  const int q = blockIdx.y;
  if (status[q] != 0) return;
  const int n = n_ptr[q];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n && i < cap; i += gridDim.x * blockDim.x) {
    const int32_t v = list[(int64_t)q * stride + i];
    atomicOr(bitmap + (int64_t)q * n_words + (v >> 5), 1u << (v & 31));
  }
}

// ---- one warp per frontier node: its neighbours that are not visited set their cand bit
__global__ void __launch_bounds__(256)
eval_expand_kernel(const int32_t* __restrict__ values, const int64_t* __restrict__ rs,
                   const int32_t* __restrict__ frontier, int64_t f_stride, const int32_t* __restrict__ f_n,
                   const uint32_t* __restrict__ visited, uint32_t* __restrict__ cand, int64_t n_words,
                   const int32_t* __restrict__ status) {
  const int q = blockIdx.y;
  if (status[q] != 0) return;
  const int n = f_n[q];
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t* vis = visited + (int64_t)q * n_words;
  uint32_t* cd = cand + (int64_t)q * n_words;
  for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += warps) {
    const int32_t node = frontier[(int64_t)q * f_stride + i];
    const int64_t b = rs[node], e = rs[node + 1];
    for (int64_t j = b + lane; j < e; j += 32) {
      const int32_t v = values[j];
      const uint32_t bit = 1u << (v & 31);
      if (!(vis[v >> 5] & bit)) atomicOr(cd + (v >> 5), bit);
    }
  }
}

constexpr int EVAL_THREADS = 1024;

// block-wide exclusive scan of one int per thread (EVAL_THREADS threads); returns the exclusive prefix, total in *tot
__device__ __forceinline__ int eval_block_scan(int v, int* s_warp, int* tot) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
  if (lane == 31) s_warp[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    int w = s_warp[lane];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, w, d); if (lane >= d) w += t; }
    s_warp[lane] = w;   // inclusive over warps
  }
  __syncthreads();
  const int before = wid > 0 ? s_warp[wid - 1] : 0;
  *tot = s_warp[31];
  __syncthreads();      // s_warp is reused by the next call
  return before + incl - v;
}

// ---- one CTA per query: cand bitmap -> ascending id list; visited |= cand; cand = 0
__global__ void __launch_bounds__(EVAL_THREADS)
eval_compact_kernel(uint32_t* __restrict__ cand, uint32_t* __restrict__ visited, int64_t n_words,
                    int32_t* __restrict__ out_ids, int64_t out_stride, int cap, int32_t* __restrict__ out_n,
                    int32_t* __restrict__ status) {
  __shared__ int s_warp[32];
  const int q = blockIdx.x;
  if (status[q] != 0) return;
  uint32_t* cd = cand + (int64_t)q * n_words;
  uint32_t* vis = visited + (int64_t)q * n_words;
  int32_t* out = out_ids + (int64_t)q * out_stride;
  int total = 0;
  bool overflow = false;
  for (int64_t base = 0; base < n_words; base += EVAL_THREADS) {
    const int64_t w = base + threadIdx.x;
    uint32_t bits = w < n_words ? cd[w] : 0u;
    if (!__syncthreads_or(bits != 0u)) continue;       // the frontier touches few words: most chunks are empty
    int chunk_tot;
    int pos = total + eval_block_scan(__popc(bits), s_warp, &chunk_tot);
    if (bits) {
      vis[w] |= bits;                                    // set_union :324
      cd[w] = 0u;
      while (bits) {
        const int b = __ffs(bits) - 1;
        bits &= bits - 1;
        if (pos < cap) out[pos] = (int32_t)(w * 32 + b);
        ++pos;
      }
    }
    total += chunk_tot;
    if (total > cap) overflow = true;
  }
  if (threadIdx.x == 0) {
    out_n[q] = total < cap ? total : cap;
    if (overflow) status[q] = NANN_RESOURCE_EXHAUSTED;
    else if (total == 1) status[q] = NANN_INVALID_ARGUMENT;   // tf.squeeze -> scalar scores cannot be concatenated (:250, :329)
  }
}

// ---- one CTA per query: frontier = Nx[sx >= R_sc[R_n - 1]] in Nx order
__global__ void __launch_bounds__(EVAL_THREADS)
eval_frontier_kernel(const int32_t* __restrict__ nx_ids, const float* __restrict__ nx_sc, int64_t nx_stride,
                     const int32_t* __restrict__ nx_n, const float* __restrict__ r_sc, int64_t r_stride,
                     const int32_t* __restrict__ r_n, int32_t* __restrict__ out_ids, int64_t out_stride,
                     int32_t* __restrict__ out_n, const int32_t* __restrict__ status) {
  __shared__ int s_warp[32];
  const int q = blockIdx.x;
  if (status[q] != 0) return;
  const int n = nx_n[q], rn = r_n[q];
  if (rn <= 0) { if (threadIdx.x == 0) out_n[q] = 0; return; }
  const float cut = r_sc[(int64_t)q * r_stride + rn - 1];
  int total = 0;
  for (int base = 0; base < n; base += EVAL_THREADS) {
    const int i = base + threadIdx.x;
    const bool keep = i < n && nx_sc[(int64_t)q * nx_stride + i] >= cut;
    int chunk_tot;
    const int pos = total + eval_block_scan(keep ? 1 : 0, s_warp, &chunk_tot);
    if (keep) out_ids[(int64_t)q * out_stride + pos] = nx_ids[(int64_t)q * nx_stride + i];
    total += chunk_tot;
  }
  if (threadIdx.x == 0) out_n[q] = total;
}

// ---- results[:topk_eval] -> item ids (:359-362)
__global__ void eval_emit_kernel(const int32_t* __restrict__ r_ids, const float* __restrict__ r_sc, int64_t r_stride,
                                 const int32_t* __restrict__ r_n, const int64_t* __restrict__ item_ids, int k, int B,
                                 int64_t* __restrict__ out_item, float* __restrict__ out_sc, int32_t* __restrict__ out_nodes,
                                 int32_t* __restrict__ out_n, const int32_t* __restrict__ status) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= (int64_t)B * k) return;
  const int q = (int)(t / k), i = (int)(t % k);
  const bool ok = status[q] == 0;
  const int n = ok ? min(r_n[q], k) : 0;
  if (i == 0) out_n[q] = n;
  if (i < n) {
    const int32_t node = r_ids[(int64_t)q * r_stride + i];
    out_item[t] = item_ids[node];
    out_sc[t] = r_sc[(int64_t)q * r_stride + i];
    out_nodes[t] = node;
  } else {
    out_item[t] = -1; out_sc[t] = 0.f; out_nodes[t] = -1;
  }
}

}  // namespace nann

struct nann_eval_searcher {
  const nann_index* ix = nullptr;
  nann_scorer* sc = nullptr;
  int max_batch = 0;
  int maxK[3] = {0, 0, 0};
  int kmax = 0;
  int64_t n_words = 0, maxc = 0;
  float* users = nullptr; float* ustate = nullptr;
  uint32_t* visited = nullptr; uint32_t* cand = nullptr;        // [maxB][n_words]
  int32_t* nx_ids = nullptr; float* nx_sc = nullptr; int32_t* nx_n = nullptr;     // [maxB][maxc], [maxB]
  int32_t* fr_ids = nullptr; int32_t* fr_n = nullptr;                             // frontier [maxB][maxc]
  int32_t* r_ids[2] = {nullptr, nullptr}; float* r_sc[2] = {nullptr, nullptr}; int32_t* r_n[2] = {nullptr, nullptr};  // [maxB][kmax]
  int32_t* status = nullptr; int32_t* out_n = nullptr;
  int64_t* out_item = nullptr; float* out_sc = nullptr; int32_t* out_nodes = nullptr;   // [maxB][topk cap]
  int out_cap = 0;
  unsigned long long* scored = nullptr;      // device counter of scored rows
  nann::TcWorkspace* tcws = nullptr;
};

namespace nann {
__global__ void eval_count_kernel(const int32_t* __restrict__ n, const int32_t* __restrict__ status, int B, int n_fixed,
                                  unsigned long long* __restrict__ total) {
  unsigned long long t = 0;
  for (int q = threadIdx.x; q < B; q += blockDim.x)
    if (status[q] == 0) t += (unsigned long long)(n ? n[q] : n_fixed);
  for (int d = 16; d > 0; d >>= 1) t += __shfl_down_sync(0xffffffffu, t, d);
  if ((threadIdx.x & 31) == 0 && t) atomicAdd(total, t);
}
}  // namespace nann

extern "C" {

void nann_eval_searcher_destroy(nann_eval_searcher_t* s) {
  if (!s) return;
  cudaFree(s->users); cudaFree(s->ustate); cudaFree(s->visited); cudaFree(s->cand);
  cudaFree(s->nx_ids); cudaFree(s->nx_sc); cudaFree(s->nx_n); cudaFree(s->fr_ids); cudaFree(s->fr_n);
  for (int i = 0; i < 2; ++i) { cudaFree(s->r_ids[i]); cudaFree(s->r_sc[i]); cudaFree(s->r_n[i]); }
  cudaFree(s->status); cudaFree(s->out_n); cudaFree(s->out_item); cudaFree(s->out_sc); cudaFree(s->out_nodes);
  cudaFree(s->scored);
  if (s->tcws) { nann::tc_ws_free(s->tcws); delete s->tcws; }
  delete s;
}

nann_status nann_eval_searcher_create(const nann_index_t* ix, nann_scorer_t* scorer, int max_batch,
                                      const int32_t max_top_k_per_level[3], int max_topk_eval,
                                      nann_eval_searcher_t** out) {
  using namespace nann;
  if (!out) return fail(NANN_INVALID_ARGUMENT, "null out");
  *out = nullptr;
  NANN_TRY(require_device());
  if (!ix || !scorer || max_batch <= 0 || !max_top_k_per_level || max_topk_eval <= 0)
    return fail(NANN_INVALID_ARGUMENT, "nann_eval_searcher_create: bad argument");
  if (max_batch > 65535) return fail(NANN_UNIMPLEMENTED, "max_batch > 65535");
  if (nann_scorer_item_dim(scorer) != ix->dim)
    return fail(NANN_INVALID_ARGUMENT, "scorer item dim %d != index dim %d", nann_scorer_item_dim(scorer), ix->dim);
  if (ix->n_local != ix->n_items) return fail(NANN_FAILED_PRECONDITION, "the index holds a slice of the table only");
  NANN_CUDA(cudaSetDevice(ix->device));
  auto* s = new nann_eval_searcher();
  s->ix = ix; s->sc = scorer; s->max_batch = max_batch;
  for (int i = 0; i < 3; ++i) {
    if (max_top_k_per_level[i] <= 0 || max_top_k_per_level[i] > TOPK_MAX_K) {
      delete s;
      return fail(NANN_UNIMPLEMENTED, "top_k_per_level[%d]=%d outside [1,%d]", i, max_top_k_per_level[i], TOPK_MAX_K);
    }
    s->maxK[i] = max_top_k_per_level[i];
    s->kmax = std::max(s->kmax, s->maxK[i]);
  }
  s->out_cap = max_topk_eval;
  s->n_words = (ix->n_items + 31) / 32;
  // a frontier holds the new nodes that reached the cut (<= K of them unless scores tie at the cut); its expansion
  // has at most max_deg neighbours per node.  Queries that exceed the capacity fail with ResourceExhausted.
  const int64_t front = std::max<int64_t>(s->kmax, 64);
  s->maxc = std::max<int64_t>(ix->n_ep, front * std::max(ix->max_deg[0], ix->max_deg[1]));
  s->maxc = std::max<int64_t>((s->maxc + 63) / 64 * 64, 64);
  const int64_t B = max_batch;
  const int uf = nann_scorer_user_floats(scorer);
  bool ok = true;
  auto A = [&](void** p, size_t bytes) { if (ok && cudaMalloc(p, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); ok = false; } };
  A((void**)&s->users, (size_t)B * uf * 4);
  A((void**)&s->ustate, (size_t)B * scorer_user_state_floats(scorer) * 4);
  A((void**)&s->visited, (size_t)B * s->n_words * 4);
  A((void**)&s->cand, (size_t)B * s->n_words * 4);
  A((void**)&s->nx_ids, (size_t)B * s->maxc * 4);
  A((void**)&s->nx_sc, (size_t)B * s->maxc * 4);
  A((void**)&s->fr_ids, (size_t)B * s->maxc * 4);
  A((void**)&s->nx_n, (size_t)B * 4);
  A((void**)&s->fr_n, (size_t)B * 4);
  for (int i = 0; i < 2; ++i) {
    A((void**)&s->r_ids[i], (size_t)B * s->kmax * 4);
    A((void**)&s->r_sc[i], (size_t)B * s->kmax * 4);
    A((void**)&s->r_n[i], (size_t)B * 4);
  }
  A((void**)&s->status, (size_t)B * 4);
  A((void**)&s->out_n, (size_t)B * 4);
  A((void**)&s->out_item, (size_t)B * s->out_cap * 8);
  A((void**)&s->out_sc, (size_t)B * s->out_cap * 4);
  A((void**)&s->out_nodes, (size_t)B * s->out_cap * 4);
  A((void**)&s->scored, 8);
  if (ok && cudaMemset(s->cand, 0, (size_t)B * s->n_words * 4) != cudaSuccess) ok = false;   // stays clean: compact clears what it reads
  if (!ok) {
    nann_eval_searcher_destroy(s);
    return fail(NANN_RESOURCE_EXHAUSTED, "OOM for the eval search workspace (batch %d)", max_batch);
  }
  *out = s;
  return NANN_OK;
}

nann_status nann_search_eval_batch(nann_eval_searcher_t* s, const float* users, int B,
                                   const int32_t num_scoring_per_level[3], const int32_t top_k_per_level[3],
                                   int topk_eval, int64_t* out_item_ids, float* out_scores, int32_t* out_nodes,
                                   int32_t* out_n, int32_t* out_status, int64_t* n_scored_total, void* stream) {
  using namespace nann;
  NANN_TRY(require_device());
  if (!s || !users || !num_scoring_per_level || !top_k_per_level)
    return fail(NANN_INVALID_ARGUMENT, "nann_search_eval_batch: null argument");
  if (B < 0 || B > s->max_batch) return fail(NANN_INVALID_ARGUMENT, "batch %d outside [0, %d]", B, s->max_batch);
  if (topk_eval <= 0 || topk_eval > s->out_cap)
    return fail(NANN_INVALID_ARGUMENT, "topk_eval %d outside [1, %d]", topk_eval, s->out_cap);
  for (int l = 0; l < 3; ++l) {
    if (top_k_per_level[l] <= 0 || top_k_per_level[l] > s->maxK[l])
      return fail(NANN_INVALID_ARGUMENT, "top_k_per_level[%d]=%d outside [1, %d]", l, top_k_per_level[l], s->maxK[l]);
    if (num_scoring_per_level[l] < 0) return fail(NANN_INVALID_ARGUMENT, "num_scoring_per_level[%d] < 0", l);
  }
  if (num_scoring_per_level[2] != 1)    // assert self.num_scoring_per_level[self.start_level] == 1  (model.py:347)
    return fail(NANN_INVALID_ARGUMENT, "num_scoring_per_level[start_level] must be 1, got %d", num_scoring_per_level[2]);
  if (n_scored_total) *n_scored_total = 0;
  if (B == 0) return NANN_OK;
  const nann_index* ix = s->ix;
  if (ix->n_ep == 0) return fail(NANN_INVALID_ARGUMENT, "the index has no enter points");
  cudaStream_t st = (cudaStream_t)stream;
  NANN_CUDA(cudaSetDevice(ix->device));
  const int uf = nann_scorer_user_floats(s->sc);
  NANN_CUDA(cudaMemcpyAsync(s->users, users, (size_t)B * uf * 4,
                            is_device_ptr(users) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
  NANN_CUDA(cudaMemsetAsync(s->status, 0, (size_t)B * 4, st));
  NANN_CUDA(cudaMemsetAsync(s->scored, 0, 8, st));
  NANN_TRY(scorer_prepare_users(s->sc, s->users, B, s->ustate, st));

  auto score = [&](const int32_t* ids, int64_t ids_stride, const int32_t* n_ptr, int n_fixed, int64_t bound) -> nann_status {
    ScoreCall c{};
    c.table = ix->emb; c.ids = ids; c.ids_stride = ids_stride; c.rows_stride = 0;
    c.n_ptr = n_ptr; c.n_fixed = n_fixed; c.max_n = (int)std::min<int64_t>(bound, s->maxc); c.B = B;
    c.hu = s->ustate; c.users = s->users; c.out = s->nx_sc; c.out_stride = s->maxc; c.status = s->status;
    if (!s->tcws) s->tcws = new nann::TcWorkspace();
    c.ws = s->tcws;
    NANN_TRY(scorer_score(s->sc, c, st));
    NANN_LAUNCH(eval_count_kernel, 1, 256, 0, st, n_ptr, s->status, B, n_fixed, s->scored);
    return NANN_OK;
  };
  int cur = 0;   // ping-pong half that holds the running result R

  // ---- start level (:350-354): every enter point is scored; clamped top K[2]
  if (ix->n_ep == 1) {   // tf.squeeze -> scalar
    std::vector<int32_t> bad((size_t)B, NANN_INVALID_ARGUMENT);
    NANN_CUDA(cudaMemcpyAsync(s->status, bad.data(), (size_t)B * 4, cudaMemcpyHostToDevice, st));
    NANN_CUDA(cudaStreamSynchronize(st));
  } else {
    NANN_TRY(score(ix->ep, 0, nullptr, (int)ix->n_ep, ix->n_ep));
    TopkArgs a{};
    a.b_sc = s->nx_sc; a.b_sc_stride = s->maxc; a.b_ids = ix->ep; a.b_ids_stride = 0; a.b_n_fixed = (int)ix->n_ep;
    a.k = top_k_per_level[2]; a.clamp = 1;
    a.out_sc = s->r_sc[cur]; a.out_ids = s->r_ids[cur]; a.out_stride = s->kmax; a.out_n = s->r_n[cur];
    a.status = s->status;
    NANN_TRY(launch_topk(a, B, st));
  }

  for (int level = 1; level >= 0; --level) {   // :356-357
    // visited_idx = idx_ep (:312); frontier = the level's entry results
    NANN_CUDA(cudaMemsetAsync(s->visited, 0, (size_t)B * s->n_words * 4, st));
    {
      dim3 grid((unsigned)ceil_div(s->kmax, 256), (unsigned)B);
      NANN_LAUNCH(eval_mark_kernel, grid, 256, 0, st, s->r_ids[cur], (int64_t)s->kmax, s->r_n[cur], s->kmax, s->visited,
                  s->n_words, s->status);
    }
    const int32_t* f_ids = s->r_ids[cur];
    int64_t f_stride = s->kmax;
    const int32_t* f_n = s->r_n[cur];
    int64_t f_cap = s->kmax;
    for (int it = 0; it < num_scoring_per_level[level]; ++it) {   // :317
      {
        dim3 grid((unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(f_cap, 8), 64)), (unsigned)B);
        NANN_LAUNCH(eval_expand_kernel, grid, 256, 0, st, ix->nbr_values[level], ix->nbr_rs[level], f_ids, f_stride, f_n,
                    s->visited, s->cand, s->n_words, s->status);
      }
      NANN_LAUNCH(eval_compact_kernel, (unsigned)B, EVAL_THREADS, 0, st, s->cand, s->visited, s->n_words, s->nx_ids,
                  s->maxc, (int)s->maxc, s->nx_n, s->status);
      NANN_TRY(score(s->nx_ids, s->maxc, s->nx_n, 0, s->maxc));
      {
        TopkArgs a{};
        a.a_sc = s->r_sc[cur]; a.a_ids = s->r_ids[cur]; a.a_stride = s->kmax; a.a_n_ptr = s->r_n[cur];
        a.b_sc = s->nx_sc; a.b_sc_stride = s->maxc; a.b_ids = s->nx_ids; a.b_ids_stride = s->maxc; a.b_n_ptr = s->nx_n;
        a.k = top_k_per_level[level]; a.clamp = 1;
        a.out_sc = s->r_sc[cur ^ 1]; a.out_ids = s->r_ids[cur ^ 1]; a.out_stride = s->kmax; a.out_n = s->r_n[cur ^ 1];
        a.status = s->status;
        NANN_TRY(launch_topk(a, B, st));
        cur ^= 1;
      }
      NANN_LAUNCH(eval_frontier_kernel, (unsigned)B, EVAL_THREADS, 0, st, s->nx_ids, s->nx_sc, s->maxc, s->nx_n, s->r_sc[cur],
                  (int64_t)s->kmax, s->r_n[cur], s->fr_ids, s->maxc, s->fr_n, s->status);
      f_ids = s->fr_ids; f_stride = s->maxc; f_n = s->fr_n; f_cap = s->maxc;
    }
  }
  NANN_LAUNCH(eval_emit_kernel, (unsigned)ceil_div((int64_t)B * topk_eval, 256), 256, 0, st, s->r_ids[cur], s->r_sc[cur],
              (int64_t)s->kmax, s->r_n[cur], ix->item_ids, topk_eval, B, s->out_item, s->out_sc, s->out_nodes, s->out_n, s->status);

  auto copy_out = [&](void* dst, const void* src, size_t bytes) -> nann_status {
    if (!dst) return NANN_OK;
    NANN_CUDA(cudaMemcpyAsync(dst, src, bytes, is_device_ptr(dst) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
    return NANN_OK;
  };
  NANN_TRY(copy_out(out_item_ids, s->out_item, (size_t)B * topk_eval * 8));
  NANN_TRY(copy_out(out_scores, s->out_sc, (size_t)B * topk_eval * 4));
  NANN_TRY(copy_out(out_nodes, s->out_nodes, (size_t)B * topk_eval * 4));
  NANN_TRY(copy_out(out_n, s->out_n, (size_t)B * 4));
  NANN_TRY(copy_out(out_status, s->status, (size_t)B * 4));
  unsigned long long h_scored = 0;
  NANN_CUDA(cudaMemcpyAsync(&h_scored, s->scored, 8, cudaMemcpyDeviceToHost, st));
  NANN_CUDA(cudaStreamSynchronize(st));
  if (n_scored_total) *n_scored_total = (int64_t)h_scored;
  return NANN_OK;
}

}  // extern "C"
