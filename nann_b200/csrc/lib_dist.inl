// lib_dist.inl -- the second multi-GPU form: ONE graph, the embedding table row-sharded, scoring distributed.
//
// nann_search_sharded (lib_shard.inl) gives every GPU its own HNSW over its own rows; holding recall at the one-GPU
// operating point then costs 1.2-1.4x the scored rows.  Here the traversal is the UNSHARDED one, so the
// result is bit-identical to a one-GPU search, and only the expensive part is spread out:
//   * the graph, enter points and item ids are replicated (27 % of the index bytes), the embedding table -- the part
//     that is gathered and scored -- is sharded by row: rank r holds rows [r*per, (r+1)*per);
//   * the global batch is partitioned: rank r runs the traversal (expand, visited filter, top-k) of ITS queries only;
//   * every scoring round, a query's candidates are bucketed by owner and sent to the owners (ids out), each rank scores
//     the candidates it owns for ALL queries of the box with the ordinary scorer kernel, and the scores travel back.
// Both exchanges are plain stores into IPC-mapped peer windows (NVLink) issued by the bucket / return kernels themselves,
// followed by a system-scope flag; waits are one-warp kernels on the same stream, so nothing synchronises the host.
// Per-GPU work = 1/N of the unsharded rows; cost = 1 + 2 x 5 flag round trips per step.
//
//   window (per rank): flags {hu[16], ids[16], sc[16]} | hu_all [G*B][ustate] | us_all [G*B][uraw] | req_ids [G*B][cap] |
//                      req_cnt [G*B] | resp_sc [G*B][cap]
//   (G = world, B = the call's batch per rank, cap = candidates per query; ustate = the per-query state the scorer
//   precomputes -- the hoisted layer-1 half of the mlp, the key projections of the attention scorer; uraw = the raw user
//   floats, shared only when the scoring kernel reads them as well: the attention scorer's 50 x 64 sequence)

struct nann_dist_group {
  int device = 0, rank = 0, world = 1, max_batch = 0, ustate = 0, uraw = 0;
  int64_t cap = 0;
  size_t off_hu = 0, off_us = 0, off_req_ids = 0, off_req_cnt = 0, off_resp = 0, window_bytes = 0;
  uint8_t* window = nullptr;
  uint8_t* peer[nann::NANN_MAX_SHARDS] = {nullptr};
  bool peer_ipc[nann::NANN_MAX_SHARDS] = {false};
  bool connected = false;
  float* svc_sc = nullptr;       // [G*max_batch][cap] scores of the candidates this rank owns, before they travel back
  int32_t* perm = nullptr;       // [max_batch][cap]   (owner << 24 | slot) of every own candidate
  unsigned int* counters = nullptr;
  int* error = nullptr;
  unsigned long long epoch_hu = 0, epoch_ids = 0, epoch_sc = 0;
  nann::TcWorkspace svc_ws;
};

namespace nann {

constexpr size_t DIST_FLAG_BYTES = 4096;
struct DistPeers { int world, rank; uint8_t* win[NANN_MAX_SHARDS]; };

// copy `n` floats to offset dst_off (and `n2` floats to dst_off2) of every rank's window, then publish
// flag[rank] = epoch everywhere (last CTA)
__global__ void dist_bcast_kernel(const float* __restrict__ src, int64_t n, size_t dst_off, const float* __restrict__ src2, int64_t n2,
                                  size_t dst_off2, DistPeers P, unsigned int* counter, size_t flag_off, unsigned long long epoch) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = src[i];
    for (int p = 0; p < P.world; ++p) reinterpret_cast<float*>(P.win[p] + dst_off)[i] = v;
  }
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n2; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = src2[i];
    for (int p = 0; p < P.world; ++p) reinterpret_cast<float*>(P.win[p] + dst_off2)[i] = v;
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int prev = atomicInc(counter, gridDim.x - 1);
    if (prev == gridDim.x - 1) {
      __threadfence_system();
      for (int p = 0; p < P.world; ++p) st_release_sys_u64(reinterpret_cast<unsigned long long*>(P.win[p] + flag_off) + P.rank, epoch);
    }
  }
}

struct DistBucketArgs {
  const int32_t* ids; int64_t ids_stride; const int32_t* n_ptr; int n_fixed; const int32_t* status;
  int B; int64_t per, cap;
  DistPeers P; size_t off_req_ids, off_req_cnt, flag_off;
  int32_t* perm; unsigned int* counter; unsigned long long epoch;
};
// one CTA per own query: candidate -> (owner, slot); the LOCAL row index goes straight into the owner's request window
__global__ void __launch_bounds__(256) dist_bucket_push_kernel(DistBucketArgs a) {
  __shared__ int cnt[NANN_MAX_SHARDS];
  const int q = blockIdx.x, tid = threadIdx.x;
  if (tid < NANN_MAX_SHARDS) cnt[tid] = 0;
  __syncthreads();
  const int n = (a.status && a.status[q] != 0) ? 0 : (a.n_ptr ? a.n_ptr[q] : a.n_fixed);
  const int64_t row = (int64_t)a.P.rank * a.B + q;              // this query's row in every owner's request window
  for (int i = tid; i < n; i += 256) {
    const int32_t id = a.ids[(int64_t)q * a.ids_stride + i];
    const int o = (int)(id / a.per);
    const int j = atomicAdd(&cnt[o], 1);
    a.perm[(int64_t)q * a.cap + i] = (o << 24) | j;
    reinterpret_cast<int32_t*>(a.P.win[o] + a.off_req_ids)[row * a.cap + j] = (int32_t)(id - (int64_t)o * a.per);
  }
  __syncthreads();
  if (tid < a.P.world) reinterpret_cast<int32_t*>(a.P.win[tid] + a.off_req_cnt)[row] = cnt[tid];
  __threadfence_system();
  __syncthreads();
  if (tid == 0) {
    const unsigned int prev = atomicInc(a.counter, (unsigned int)(a.B - 1));
    if (prev == (unsigned int)(a.B - 1)) {
      __threadfence_system();
      for (int p = 0; p < a.P.world; ++p) st_release_sys_u64(reinterpret_cast<unsigned long long*>(a.P.win[p] + a.flag_off) + a.P.rank, a.epoch);
    }
  }
}

struct DistReturnArgs {
  const float* svc_sc; const int32_t* req_cnt; int B; int64_t cap;
  DistPeers P; size_t off_resp, flag_off; unsigned int* counter; unsigned long long epoch;
};
// one CTA per (source rank, query): the scores of the candidates this rank owns go back into the source's window
__global__ void __launch_bounds__(256) dist_return_kernel(DistReturnArgs a) {
  const int pq = blockIdx.x;                                     // src * B + q
  const int src = pq / a.B, q = pq % a.B;
  const int n = a.req_cnt[pq];
  float* dst = reinterpret_cast<float*>(a.P.win[src] + a.off_resp) + ((int64_t)a.P.rank * a.B + q) * a.cap;
  const float* s = a.svc_sc + (int64_t)pq * a.cap;
  for (int i = threadIdx.x; i < n; i += 256) dst[i] = s[i];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int total = (unsigned int)(a.P.world * a.B);
    const unsigned int prev = atomicInc(a.counter, total - 1);
    if (prev == total - 1) {
      __threadfence_system();
      for (int p = 0; p < a.P.world; ++p) st_release_sys_u64(reinterpret_cast<unsigned long long*>(a.P.win[p] + a.flag_off) + a.P.rank, a.epoch);
    }
  }
}
// one CTA per own query: scores back into candidate order
__global__ void __launch_bounds__(256)
dist_unbucket_kernel(const int32_t* __restrict__ perm, const float* __restrict__ resp, const int32_t* __restrict__ n_ptr, int n_fixed,
                     const int32_t* __restrict__ status, int B, int64_t cap, float* __restrict__ out, int64_t out_stride) {
  const int q = blockIdx.x;
  const int n = (status && status[q] != 0) ? 0 : (n_ptr ? n_ptr[q] : n_fixed);
  for (int i = threadIdx.x; i < n; i += 256) {
    const int32_t pk = perm[(int64_t)q * cap + i];
    const int o = pk >> 24, j = pk & 0xffffff;
    out[(int64_t)q * out_stride + i] = resp[((int64_t)o * B + q) * cap + j];
  }
}

static DistPeers dist_peers(const nann_dist_group* g) {
  DistPeers P{};
  P.world = g->world; P.rank = g->rank;
  for (int p = 0; p < g->world; ++p) P.win[p] = g->peer[p];
  return P;
}

// the users' precomputed state (and, for the attention scorer, the raw sequences) -> every rank (once per call)
static nann_status dist_share_user_state(nann_dist_group* g, const float* ustate, const float* users, int B, cudaStream_t st) {
  const int64_t n = (int64_t)B * g->ustate, n2 = (int64_t)B * g->uraw;
  const unsigned long long epoch = ++g->epoch_hu;
  const unsigned grid = (unsigned)std::min<int64_t>(ceil_div(n, 256), 148 * 2);
  NANN_LAUNCH(dist_bcast_kernel, grid, 256, 0, st, ustate, n, g->off_hu + (size_t)g->rank * B * g->ustate * 4, users, n2,
              g->off_us + (size_t)g->rank * B * g->uraw * 4, dist_peers(g), g->counters + 0, (size_t)0, epoch);
  NANN_LAUNCH(shard_wait_kernel, 1, 32, 0, st, (const unsigned long long*)g->window, g->world, epoch - 1, g->error);
  return NANN_OK;
}

// one scoring round: ids out, score what this rank owns for the whole box, scores back
static nann_status dist_score_round(nann_searcher* s, nann_dist_group* g, const int32_t* ids, int64_t ids_stride, const int32_t* n_ptr,
                                    int n_fixed, int64_t bound, int B, cudaStream_t st) {
  const nann_index* ix = s->ix;
  const int G = g->world;
  const int64_t per = ceil_div(ix->n_items, G);
  DistBucketArgs b{};
  b.ids = ids; b.ids_stride = ids_stride; b.n_ptr = n_ptr; b.n_fixed = n_fixed; b.status = s->status;
  b.B = B; b.per = per; b.cap = g->cap; b.P = dist_peers(g);
  b.off_req_ids = g->off_req_ids; b.off_req_cnt = g->off_req_cnt; b.flag_off = NANN_MAX_SHARDS * 8;
  b.perm = g->perm; b.counter = g->counters + 1; b.epoch = ++g->epoch_ids;
  NANN_LAUNCH(dist_bucket_push_kernel, (unsigned)B, 256, 0, st, b);
  NANN_LAUNCH(shard_wait_kernel, 1, 32, 0, st, (const unsigned long long*)(g->window + NANN_MAX_SHARDS * 8), G, b.epoch - 1, g->error);
  // every candidate this rank owns, of every query of the box: the ordinary scorer on G*B pseudo-queries
  ScoreCall c{};
  c.table = ix->emb; c.ids = (const int32_t*)(g->window + g->off_req_ids); c.ids_stride = g->cap; c.rows_stride = 0;
  c.n_ptr = (const int32_t*)(g->window + g->off_req_cnt); c.n_fixed = 0;
  c.max_n = (int)std::min<int64_t>(bound, g->cap); c.B = G * B;
  c.hu = (const float*)(g->window + g->off_hu);
  c.users = g->uraw ? (const float*)(g->window + g->off_us) : nullptr;
  c.out = g->svc_sc; c.out_stride = g->cap; c.status = nullptr;
  c.ws = &g->svc_ws;
  NANN_TRY(scorer_score(s->sc, c, st));
  DistReturnArgs r{};
  r.svc_sc = g->svc_sc; r.req_cnt = c.n_ptr; r.B = B; r.cap = g->cap; r.P = dist_peers(g);
  r.off_resp = g->off_resp; r.flag_off = 2 * NANN_MAX_SHARDS * 8; r.counter = g->counters + 2; r.epoch = ++g->epoch_sc;
  NANN_LAUNCH(dist_return_kernel, (unsigned)(G * B), 256, 0, st, r);
  NANN_LAUNCH(shard_wait_kernel, 1, 32, 0, st, (const unsigned long long*)(g->window + 2 * NANN_MAX_SHARDS * 8), G, r.epoch - 1, g->error);
  NANN_LAUNCH(dist_unbucket_kernel, (unsigned)B, 256, 0, st, g->perm, (const float*)(g->window + g->off_resp), n_ptr, n_fixed, s->status, B,
              g->cap, s->cand_sc, s->maxc);
  return NANN_OK;
}

}  // namespace nann

extern "C" {

void nann_dist_group_destroy(nann_dist_group_t* g) {
  if (!g) return;
  cudaSetDevice(g->device);
  cudaDeviceSynchronize();
  for (int p = 0; p < g->world; ++p)
    if (p != g->rank && g->peer[p] && g->peer_ipc[p]) cudaIpcCloseMemHandle(g->peer[p]);
  cudaFree(g->window); cudaFree(g->svc_sc); cudaFree(g->perm); cudaFree(g->counters); cudaFree(g->error);
  nann::tc_ws_free(&g->svc_ws);
  delete g;
}

nann_status nann_dist_group_create(const nann_searcher_t* s, int rank, int world, nann_dist_group_t** out) {
  if (!out) return fail(NANN_INVALID_ARGUMENT, "null out");
  *out = nullptr;
  if (!s) return fail(NANN_INVALID_ARGUMENT, "null searcher");
  if (world < 1 || world > NANN_MAX_SHARDS || rank < 0 || rank >= world)
    return fail(NANN_INVALID_ARGUMENT, "rank %d / world %d (at most %d ranks)", rank, world, NANN_MAX_SHARDS);
  const nann_index* ix = s->ix;
  const int64_t per = ceil_div(ix->n_items, world);
  if (ix->row_lo != (int64_t)rank * per || ix->n_local != std::min<int64_t>(per, ix->n_items - ix->row_lo))
    return fail(NANN_INVALID_ARGUMENT, "rank %d of %d must hold table rows [%lld, %lld); the index holds [%lld, %lld)", rank, world,
                (long long)(rank * per), (long long)std::min<int64_t>((rank + 1) * per, ix->n_items), (long long)ix->row_lo,
                (long long)(ix->row_lo + ix->n_local));
  if (s->maxc >= (1 << 24)) return fail(NANN_UNIMPLEMENTED, "more than 2^24 candidates per query and round");
  NANN_CUDA(cudaSetDevice(ix->device));
  NANN_TRY(require_device());
  auto* g = new nann_dist_group();
  g->device = ix->device; g->rank = rank; g->world = world; g->max_batch = s->max_batch; g->cap = s->maxc;
  g->ustate = (int)scorer_user_state_floats(s->sc);
  g->uraw = s->sc->kind == 1 ? nann_scorer_user_floats(s->sc) : 0;   // the attention kernel also reads the raw sequence
  auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
  const size_t rows = (size_t)world * g->max_batch;
  g->off_hu = DIST_FLAG_BYTES;
  g->off_us = g->off_hu + al(rows * g->ustate * 4);
  g->off_req_ids = g->off_us + al(rows * g->uraw * 4);
  g->off_req_cnt = g->off_req_ids + al(rows * g->cap * 4);
  g->off_resp = g->off_req_cnt + al(rows * 4);
  g->window_bytes = g->off_resp + al(rows * g->cap * 4);
  nann_status rc = NANN_OK;
  if (cudaMalloc(&g->window, g->window_bytes) != cudaSuccess || cudaMalloc(&g->svc_sc, rows * g->cap * 4) != cudaSuccess ||
      cudaMalloc(&g->perm, (size_t)g->max_batch * g->cap * 4) != cudaSuccess || cudaMalloc(&g->counters, 16) != cudaSuccess ||
      cudaMalloc(&g->error, sizeof(int)) != cudaSuccess) {
    cudaGetLastError();
    rc = fail(NANN_RESOURCE_EXHAUSTED, "OOM for the distributed-scoring window (%zu bytes)", g->window_bytes);
  } else if (cudaMemset(g->window, 0, DIST_FLAG_BYTES) != cudaSuccess || cudaMemset(g->counters, 0, 16) != cudaSuccess ||
             cudaMemset(g->error, 0, sizeof(int)) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
    rc = fail(NANN_INTERNAL, "distributed-scoring group setup failed: %s", cudaGetErrorString(cudaGetLastError()));
  }
  // the tensor-core scorer's tile list for the G*B pseudo-queries, sized now: growing it inside a call would put a
  // cudaMalloc / cudaFree (device-wide synchronisation) between the enqueues of two members driven by one thread
  if (rc == NANN_OK) rc = nann::tc_ws_ensure(&g->svc_ws, (int)rows, (int64_t)rows * ceil_div(g->cap, nann::TC_M));
  if (rc != NANN_OK) { nann_dist_group_destroy(g); return rc; }
  g->peer[rank] = g->window;
  if (world == 1) g->connected = true;
  *out = g;
  return NANN_OK;
}

nann_status nann_dist_group_export(nann_dist_group_t* g, void* handle) {
  if (!g || !handle) return fail(NANN_INVALID_ARGUMENT, "null argument");
  NANN_CUDA(cudaSetDevice(g->device));
  cudaIpcMemHandle_t h;
  NANN_CUDA(cudaIpcGetMemHandle(&h, g->window));
  memcpy(handle, &h, sizeof(h));
  return NANN_OK;
}

nann_status nann_dist_group_connect(nann_dist_group_t* g, const void* handles) {
  if (!g || !handles) return fail(NANN_INVALID_ARGUMENT, "null argument");
  if (g->connected) return fail(NANN_FAILED_PRECONDITION, "group is already connected");
  NANN_CUDA(cudaSetDevice(g->device));
  for (int p = 0; p < g->world; ++p) {
    if (p == g->rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, (const uint8_t*)handles + (size_t)p * sizeof(h), sizeof(h));
    void* ptr = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      cudaGetLastError();
      return fail(NANN_FAILED_PRECONDITION, "cudaIpcOpenMemHandle for the window of rank %d failed: %s", p, cudaGetErrorString(e));
    }
    g->peer[p] = (uint8_t*)ptr;
    g->peer_ipc[p] = true;
  }
  g->connected = true;
  return NANN_OK;
}

nann_status nann_dist_group_connect_local(nann_dist_group_t* const* members, int world) {
  if (!members || world < 1 || world > NANN_MAX_SHARDS) return fail(NANN_INVALID_ARGUMENT, "bad member list");
  for (int r = 0; r < world; ++r) {
    nann_dist_group* g = members[r];
    if (!g || g->world != world || g->rank != r) return fail(NANN_INVALID_ARGUMENT, "members[%d] is not rank %d of %d", r, r, world);
    if (g->max_batch != members[0]->max_batch || g->cap != members[0]->cap || g->ustate != members[0]->ustate ||
        g->uraw != members[0]->uraw)
      return fail(NANN_INVALID_ARGUMENT, "members differ in batch / candidate capacity / scorer");
  }
  for (int r = 0; r < world; ++r) {
    nann_dist_group* g = members[r];
    NANN_CUDA(cudaSetDevice(g->device));
    for (int p = 0; p < world; ++p) {
      if (members[p]->device != g->device) {
        cudaError_t e = cudaDeviceEnablePeerAccess(members[p]->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) NANN_CUDA(e);
        cudaGetLastError();
      }
      g->peer[p] = members[p]->window;
    }
    g->connected = true;
  }
  return NANN_OK;
}

nann_status nann_dist_group_check(nann_dist_group_t* g) {
  if (!g) return fail(NANN_INVALID_ARGUMENT, "null group");
  NANN_CUDA(cudaSetDevice(g->device));
  int err = 0;
  NANN_CUDA(cudaMemcpy(&err, g->error, sizeof(int), cudaMemcpyDeviceToHost));     // synchronises with the work issued so far
  if (err) return fail(NANN_DEADLINE_EXCEEDED, "distributed scoring timed out: a peer did not deliver (ranks out of step, or a rank died)");
  return NANN_OK;
}

nann_status nann_search_distributed(nann_searcher_t* s, nann_dist_group_t* g, const float* users, int B, const int32_t T[6],
                                    int64_t* out_item_ids, float* out_scores, int32_t* out_status, nann_search_stats_t* stats,
                                    void* stream) {
  NANN_TRY(require_device());
  if (!g) return fail(NANN_INVALID_ARGUMENT, "null group");
  NANN_TRY(search_check_args(s, users, B, T));
  if (!g->connected) return fail(NANN_FAILED_PRECONDITION, "group is not connected (nann_dist_group_connect)");
  if (g->max_batch != s->max_batch || g->cap != s->maxc) return fail(NANN_INVALID_ARGUMENT, "group was created for another searcher");
  if (g->ustate != (int)scorer_user_state_floats(s->sc) || g->uraw != (s->sc->kind == 1 ? nann_scorer_user_floats(s->sc) : 0))
    return fail(NANN_INVALID_ARGUMENT, "group was created for another scorer");
  if (B == 0) return fail(NANN_INVALID_ARGUMENT, "every rank must bring the same, non-zero number of queries");
  cudaStream_t st = (cudaStream_t)stream;
  NANN_CUDA(cudaSetDevice(g->device));
  s->dist = g;
  const nann_status rc = search_enqueue(s, users, B, T, st, nullptr);
  s->dist = nullptr;
  NANN_TRY(rc);
  // the error flag is looked at when the call synchronises anyway (host outputs / stats / profile); otherwise by
  // nann_dist_group_check
  const bool will_sync = (out_item_ids && !is_device_ptr(out_item_ids)) || (out_scores && !is_device_ptr(out_scores)) ||
                         (out_status && !is_device_ptr(out_status)) || stats || s->profile;
  if (stats) memset(stats, 0, sizeof(*stats));
  NANN_TRY(search_deliver(s, B, T[5], out_item_ids, out_scores, out_status, stats, st));
  if (will_sync) {
    int err = 0;
    NANN_CUDA(cudaMemcpy(&err, g->error, sizeof(int), cudaMemcpyDeviceToHost));
    if (err) return fail(NANN_DEADLINE_EXCEEDED, "distributed scoring timed out: a peer did not deliver (ranks out of step, or a rank died)");
  }
  return NANN_OK;
}

}  // extern "C"
