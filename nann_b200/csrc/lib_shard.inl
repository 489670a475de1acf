// lib_shard.inl -- row-sharded retrieval across the GPUs of one box (the reference has replicas only,
// blaze-benchmark/benchmark/core/model.cc:192-234).
//
// Each rank (one process per GPU, or several members in one process) owns the rows of one shard with their item ids
// and an independent HNSW.  Every query visits every shard; the ONE exchange step is an all-gather of the per-shard
// (score f32, item id i64)[B][k_s] records, done WITHOUT a collective library call: every rank owns a receive window
// in its HBM, maps the windows of all peers (CUDA IPC between processes, peer access inside one process) and the
// final top-k kernel of a search stores its records straight into all windows over NVLink (topk_kernel's push
// epilogue, traverse_kernels.cuh).  A one-warp wait kernel and the merge kernel run on the group's own stream, so
// the exchange + merge of batch i overlaps the search of batch i+1 on the caller's stream; windows are
// double-buffered and a slot is rewritten only after every rank reported (done flags) that it merged the sequence
// that used it before: a one-warp back-pressure kernel in front of the final top-k waits for that (never the top-k
// CTAs themselves -- spinning CTAs that fill every SM keep this GPU's own merge kernel from being scheduled).  No host
// synchronisation anywhere unless the caller asked for host outputs.  The hand-off is modelled under skewed timings
// in tests/test_exchange_protocol_model.py.
//
//   window (per rank):  flags { arrive[depth][16], done[16] } | depth x { sc [B][G][k] | ids [B][G][k] | st [B][G] }

struct nann_shard_group {
  int device = 0, rank = 0, world = 1, depth = 2;
  int max_batch = 0, max_k = 0;
  size_t slot_bytes = 0, sc_off = 0, ids_off = 0, st_off = 0, window_bytes = 0;
  uint8_t* window = nullptr;                       // own receive window (cudaMalloc, IPC-exportable)
  uint8_t* peer[nann::NANN_MAX_SHARDS] = {nullptr};  // mapped windows, peer[rank] == window
  bool peer_ipc[nann::NANN_MAX_SHARDS] = {false};
  bool connected = false;
  unsigned int* counters = nullptr;                // [0] push CTAs, [1] merge CTAs
  int* error = nullptr;                            // device flag: a wait timed out
  unsigned long long seq = 0;
  bool pending = false;                            // a push whose merge has not been enqueued yet
  int pend_B = 0, pend_k = 0;
  cudaStream_t merge_stream = nullptr;
  cudaEvent_t ev_push = nullptr, ev_merged = nullptr;
  // device staging for host outputs
  float* o_sc = nullptr; int64_t* o_ids = nullptr; int32_t* o_st = nullptr;
  int o_k = 0;
};

namespace nann {
constexpr size_t SHARD_FLAG_BYTES = 4096;
static inline unsigned long long* shard_arrive(uint8_t* win, int slot, int r) {
  return (unsigned long long*)win + (size_t)slot * NANN_MAX_SHARDS + r;
}
static inline unsigned long long* shard_done(uint8_t* win, int depth, int r) {
  return (unsigned long long*)win + (size_t)depth * NANN_MAX_SHARDS + r;
}
}  // namespace nann

extern "C" {

void nann_shard_group_destroy(nann_shard_group_t* g) {
  if (!g) return;
  cudaSetDevice(g->device);
  if (g->merge_stream) cudaStreamSynchronize(g->merge_stream);
  for (int p = 0; p < g->world; ++p)
    if (p != g->rank && g->peer[p] && g->peer_ipc[p]) cudaIpcCloseMemHandle(g->peer[p]);
  cudaFree(g->window); cudaFree(g->counters); cudaFree(g->error);
  cudaFree(g->o_sc); cudaFree(g->o_ids); cudaFree(g->o_st);
  if (g->ev_push) cudaEventDestroy(g->ev_push);
  if (g->ev_merged) cudaEventDestroy(g->ev_merged);
  if (g->merge_stream) cudaStreamDestroy(g->merge_stream);
  delete g;
}

nann_status nann_shard_group_create(int device, int rank, int world, int max_batch, int max_k_shard,
                                    nann_shard_group_t** out) {
  if (!out) return fail(NANN_INVALID_ARGUMENT, "null out");
  *out = nullptr;
  if (world < 1 || world > NANN_MAX_SHARDS || rank < 0 || rank >= world)
    return fail(NANN_INVALID_ARGUMENT, "rank %d / world %d (at most %d shards)", rank, world, NANN_MAX_SHARDS);
  if (max_batch <= 0 || max_k_shard <= 0 || (int64_t)world * max_k_shard > 8192)
    return fail(NANN_INVALID_ARGUMENT, "max_batch %d, max_k_shard %d: need world*k <= 8192", max_batch, max_k_shard);
  NANN_CUDA(cudaSetDevice(device));
  NANN_TRY(require_device());
  auto* g = new nann_shard_group();
  g->device = device; g->rank = rank; g->world = world; g->max_batch = max_batch; g->max_k = max_k_shard;
  const size_t rec = (size_t)max_batch * world * max_k_shard;
  auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
  g->sc_off = 0; g->ids_off = al(rec * 4); g->st_off = g->ids_off + al(rec * 8);
  g->slot_bytes = g->st_off + al((size_t)max_batch * world * 4);
  g->window_bytes = SHARD_FLAG_BYTES + (size_t)g->depth * g->slot_bytes;
  nann_status rc = NANN_OK;
  if (cudaMalloc(&g->window, g->window_bytes) != cudaSuccess || cudaMalloc(&g->counters, 2 * sizeof(unsigned int)) != cudaSuccess ||
      cudaMalloc(&g->error, sizeof(int)) != cudaSuccess) {
    cudaGetLastError();
    rc = fail(NANN_RESOURCE_EXHAUSTED, "OOM for the shard window (%zu bytes)", g->window_bytes);
  } else if (cudaMemset(g->window, 0, SHARD_FLAG_BYTES) != cudaSuccess || cudaMemset(g->counters, 0, 2 * sizeof(unsigned int)) != cudaSuccess ||
             cudaMemset(g->error, 0, sizeof(int)) != cudaSuccess ||
             cudaStreamCreateWithFlags(&g->merge_stream, cudaStreamNonBlocking) != cudaSuccess ||
             cudaEventCreateWithFlags(&g->ev_push, cudaEventDisableTiming) != cudaSuccess ||
             cudaEventCreateWithFlags(&g->ev_merged, cudaEventDisableTiming) != cudaSuccess ||
             cudaDeviceSynchronize() != cudaSuccess) {
    rc = fail(NANN_INTERNAL, "shard group setup failed: %s", cudaGetErrorString(cudaGetLastError()));
  }
  if (rc != NANN_OK) { nann_shard_group_destroy(g); return rc; }
  g->peer[rank] = g->window;
  if (world == 1) g->connected = true;
  *out = g;
  return NANN_OK;
}

int nann_shard_group_handle_bytes(void) { return (int)sizeof(cudaIpcMemHandle_t); }

nann_status nann_shard_group_export(nann_shard_group_t* g, void* handle) {
  if (!g || !handle) return fail(NANN_INVALID_ARGUMENT, "null argument");
  NANN_CUDA(cudaSetDevice(g->device));
  cudaIpcMemHandle_t h;
  NANN_CUDA(cudaIpcGetMemHandle(&h, g->window));
  memcpy(handle, &h, sizeof(h));
  return NANN_OK;
}

nann_status nann_shard_group_connect(nann_shard_group_t* g, const void* handles) {
  if (!g || !handles) return fail(NANN_INVALID_ARGUMENT, "null argument");
  if (g->connected) return fail(NANN_FAILED_PRECONDITION, "shard group is already connected");
  NANN_CUDA(cudaSetDevice(g->device));
  for (int p = 0; p < g->world; ++p) {
    if (p == g->rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, (const uint8_t*)handles + (size_t)p * sizeof(h), sizeof(h));
    void* ptr = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      cudaGetLastError();
      return fail(NANN_FAILED_PRECONDITION, "cudaIpcOpenMemHandle for the window of rank %d failed: %s (GPUs without "
                  "peer access, or IPC disabled in this container)", p, cudaGetErrorString(e));
    }
    g->peer[p] = (uint8_t*)ptr;
    g->peer_ipc[p] = true;
  }
  g->connected = true;
  return NANN_OK;
}

nann_status nann_shard_group_connect_local(nann_shard_group_t* const* members, int world) {
  if (!members || world < 1 || world > NANN_MAX_SHARDS) return fail(NANN_INVALID_ARGUMENT, "bad member list");
  for (int r = 0; r < world; ++r) {
    nann_shard_group* g = members[r];
    if (!g || g->world != world || g->rank != r) return fail(NANN_INVALID_ARGUMENT, "members[%d] is not rank %d of %d", r, r, world);
    if (g->max_batch != members[0]->max_batch || g->max_k != members[0]->max_k)
      return fail(NANN_INVALID_ARGUMENT, "members differ in max_batch / max_k_shard");
    if (g->connected && world > 1) return fail(NANN_FAILED_PRECONDITION, "members[%d] is already connected", r);
  }
  for (int r = 0; r < world; ++r) {
    nann_shard_group* g = members[r];
    NANN_CUDA(cudaSetDevice(g->device));
    for (int p = 0; p < world; ++p) {
      if (members[p]->device != g->device) {
        int can = 0;
        NANN_CUDA(cudaDeviceCanAccessPeer(&can, g->device, members[p]->device));
        if (!can) return fail(NANN_FAILED_PRECONDITION, "device %d cannot access device %d", g->device, members[p]->device);
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

// phase 1: this shard's search on `stream`, its records pushed into every rank's window by the final top-k kernel
nann_status nann_search_sharded_push(nann_searcher_t* s, nann_shard_group_t* g, const float* users, int B,
                                     const int32_t T[6], void* stream) {
  NANN_TRY(require_device());
  if (!g) return fail(NANN_INVALID_ARGUMENT, "null shard group");
  NANN_TRY(search_check_args(s, users, B, T));
  if (!g->connected) return fail(NANN_FAILED_PRECONDITION, "shard group is not connected (nann_shard_group_connect)");
  if (g->pending) return fail(NANN_FAILED_PRECONDITION, "a pushed search is waiting for nann_search_sharded_merge");
  if (s->ix->device != g->device) return fail(NANN_INVALID_ARGUMENT, "searcher is on device %d, shard group on %d", s->ix->device, g->device);
  if (s->ix->n_local != s->ix->n_items) return fail(NANN_FAILED_PRECONDITION, "the index holds a slice of the table only: use nann_search_distributed");
  const int k_s = T[5];
  if (B > g->max_batch || k_s > g->max_k) return fail(NANN_INVALID_ARGUMENT, "batch %d / k %d exceed the group's window (%d / %d)", B, k_s, g->max_batch, g->max_k);
  if (B == 0) return NANN_OK;          // every rank sees the same B, so every rank skips the sequence
  cudaStream_t st = (cudaStream_t)stream;
  NANN_CUDA(cudaSetDevice(g->device));
  const unsigned long long seq = g->seq;
  const int slot = (int)(seq % (unsigned)g->depth);
  const size_t slot_off = SHARD_FLAG_BYTES + (size_t)slot * g->slot_bytes;
  ShardPush P{};
  P.world = g->world; P.rank = g->rank; P.k = k_s; P.B = B;
  for (int p = 0; p < g->world; ++p) {
    uint8_t* w = g->peer[p] + slot_off;
    P.sc[p] = (float*)(w + g->sc_off); P.ids[p] = (int64_t*)(w + g->ids_off); P.st[p] = (int32_t*)(w + g->st_off);
    P.arrive[p] = shard_arrive(g->peer[p], slot, g->rank);
  }
  P.my_done = shard_done(g->window, g->depth, 0);
  P.seq = seq; P.need_done = seq >= (unsigned)g->depth ? seq + 1 - g->depth : 0;
  P.counter = g->counters + 0; P.error = g->error;
  NANN_TRY(search_enqueue(s, users, B, T, st, &P));
  NANN_CUDA(cudaEventRecord(g->ev_push, st));
  g->seq = seq + 1;
  g->pending = true; g->pend_B = B; g->pend_k = k_s;
  return NANN_OK;
}

// phase 2, on the group's stream (the caller's stream is free for the next batch): wait for every shard's delivery,
// merge, tell the peers the slot is free
nann_status nann_search_sharded_merge(nann_shard_group_t* g, int k_out, int64_t* out_item_ids, float* out_scores,
                                      int32_t* out_status) {
  NANN_TRY(require_device());
  if (!g) return fail(NANN_INVALID_ARGUMENT, "null shard group");
  if (!g->pending) return NANN_OK;     // B == 0 push, or nothing pushed
  const int B = g->pend_B, k_s = g->pend_k;
  if (k_out < 0 || (int64_t)g->world * k_s < k_out)   // TopKV2 on the concatenation: topk_op.cc:66-69
    return fail(NANN_INVALID_ARGUMENT, "input must have at least k columns. Had %lld, needed %d", (long long)g->world * k_s, k_out);
  NANN_CUDA(cudaSetDevice(g->device));
  const unsigned long long seq = g->seq - 1;
  const int slot = (int)(seq % (unsigned)g->depth);
  const size_t slot_off = SHARD_FLAG_BYTES + (size_t)slot * g->slot_bytes;
  const bool host_out = (out_item_ids && !is_device_ptr(out_item_ids)) || (out_scores && !is_device_ptr(out_scores)) ||
                        (out_status && !is_device_ptr(out_status));
  const bool need_stage = host_out || !out_item_ids || !out_scores || !out_status;
  if (need_stage && (g->o_k < k_out || !g->o_sc)) {
    NANN_CUDA(cudaStreamSynchronize(g->merge_stream));
    cudaFree(g->o_sc); cudaFree(g->o_ids); cudaFree(g->o_st);
    g->o_sc = nullptr; g->o_ids = nullptr; g->o_st = nullptr; g->o_k = 0;
    NANN_CUDA(cudaMalloc(&g->o_sc, (size_t)g->max_batch * std::max(k_out, 1) * 4));
    NANN_CUDA(cudaMalloc(&g->o_ids, (size_t)g->max_batch * std::max(k_out, 1) * 8));
    NANN_CUDA(cudaMalloc(&g->o_st, (size_t)g->max_batch * 4));
    g->o_k = std::max(k_out, 1);
  }
  cudaStream_t ms = g->merge_stream;
  g->pending = false;
  NANN_CUDA(cudaStreamWaitEvent(ms, g->ev_push, 0));
  NANN_LAUNCH(shard_wait_kernel, 1, 32, 0, ms, shard_arrive(g->window, slot, 0), g->world, seq, g->error);
  ShardMergeArgs M{};
  M.world = g->world; M.rank = g->rank; M.k_in = k_s; M.k_out = k_out; M.B = B;
  uint8_t* w = g->window + slot_off;
  M.sc = (const float*)(w + g->sc_off); M.ids = (const int64_t*)(w + g->ids_off); M.st = (const int32_t*)(w + g->st_off);
  const bool d_sc = out_scores && is_device_ptr(out_scores), d_ids = out_item_ids && is_device_ptr(out_item_ids);
  const bool d_st = out_status && is_device_ptr(out_status);
  M.out_sc = d_sc ? out_scores : g->o_sc; M.out_ids = d_ids ? out_item_ids : g->o_ids; M.out_status = d_st ? out_status : g->o_st;
  for (int p = 0; p < g->world; ++p) M.done[p] = shard_done(g->peer[p], g->depth, g->rank);
  M.seq = seq; M.counter = g->counters + 1; M.error = g->error;
  int npad = 1;
  while (npad < g->world * k_s) npad <<= 1;
  NANN_LAUNCH(shard_merge_kernel, (unsigned)B, SHARD_MERGE_THREADS, (size_t)npad * 8, ms, M);
  if (host_out) {
    if (out_scores && !d_sc) NANN_CUDA(cudaMemcpyAsync(out_scores, g->o_sc, (size_t)B * k_out * 4, cudaMemcpyDeviceToHost, ms));
    if (out_item_ids && !d_ids) NANN_CUDA(cudaMemcpyAsync(out_item_ids, g->o_ids, (size_t)B * k_out * 8, cudaMemcpyDeviceToHost, ms));
    if (out_status && !d_st) NANN_CUDA(cudaMemcpyAsync(out_status, g->o_st, (size_t)B * 4, cudaMemcpyDeviceToHost, ms));
  }
  NANN_CUDA(cudaEventRecord(g->ev_merged, ms));
  if (host_out) {
    NANN_CUDA(cudaStreamSynchronize(ms));
    int err = 0;
    NANN_CUDA(cudaMemcpy(&err, g->error, sizeof(int), cudaMemcpyDeviceToHost));
    if (err) return fail(NANN_DEADLINE_EXCEEDED, "shard exchange timed out: a peer did not deliver (ranks out of step, or a rank died)");
  }
  return NANN_OK;
}

nann_status nann_search_sharded(nann_searcher_t* s, nann_shard_group_t* g, const float* users, int B,
                                const int32_t T[6], int k_out, int64_t* out_item_ids, float* out_scores,
                                int32_t* out_status, void* stream) {
  if (g && T && (k_out < 0 || (int64_t)g->world * T[5] < k_out))
    return fail(NANN_INVALID_ARGUMENT, "input must have at least k columns. Had %lld, needed %d", (long long)g->world * T[5], k_out);
  NANN_TRY(nann_search_sharded_push(s, g, users, B, T, stream));
  return nann_search_sharded_merge(g, k_out, out_item_ids, out_scores, out_status);
}

nann_status nann_shard_group_wait(nann_shard_group_t* g, void* stream, int host_block) {
  if (!g) return fail(NANN_INVALID_ARGUMENT, "null shard group");
  NANN_CUDA(cudaSetDevice(g->device));
  if (g->seq == 0) return NANN_OK;
  if (host_block) {
    NANN_CUDA(cudaEventSynchronize(g->ev_merged));
    int err = 0;
    NANN_CUDA(cudaMemcpy(&err, g->error, sizeof(int), cudaMemcpyDeviceToHost));
    if (err) return fail(NANN_DEADLINE_EXCEEDED, "shard exchange timed out: a peer did not deliver (ranks out of step, or a rank died)");
    return NANN_OK;
  }
  NANN_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, g->ev_merged, 0));
  return NANN_OK;
}

}  // extern "C"
