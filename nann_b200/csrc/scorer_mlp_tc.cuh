// scorer_mlp_tc.cuh -- host side of NANN_SCORER_TENSOR: fused row gather + "mlp2x512" scorer on the 5th-gen
// tensor cores (tcgen05.mma, accumulators in TMEM); kernel in scorer_mlp_tc8.cuh, numerics in scorer_tc_common.cuh.
// Earlier generations (bulk-synchronous cp.async kernel, TMA ring with an L2 h1 scratch, multicast weight stages)
// were measured slower and live in the git history only.
#pragma once
#include "scorer_tc_common.cuh"
#include "scorer_mlp_tc8.cuh"

namespace nann {

struct MlpTcState {
  float h_b2[MLP_H], h_w3[MLP_H];   // host copies for the by-value kernel parameters
  __half* W8img = nullptr;          // [rank 2][unit 10][hi 32 KB | lo 32 KB], 1.25 MB (scorer_mlp_tc8.cuh)
  int n_ctas = 0;
};

// Per-caller scratch of the tensor-core scorer (one per searcher / per op call, i.e. per stream: concurrent
// launches must not share it): the round's dense tile list.
struct TcWorkspace {
  int32_t* tile_start = nullptr;   // [b_cap + 1]
  int b_cap = 0;
  int2* tiles = nullptr;           // [tiles_cap]
  int64_t tiles_cap = 0;
};
static void tc_ws_free(TcWorkspace* ws) {
  if (!ws) return;
  cudaFree(ws->tile_start); cudaFree(ws->tiles);
  *ws = TcWorkspace();
}
static nann_status tc_ws_ensure(TcWorkspace* ws, int B, int64_t max_tiles) {
  if (ws->b_cap < B) {
    cudaFree(ws->tile_start); ws->tile_start = nullptr; ws->b_cap = 0;
    NANN_CUDA(cudaMalloc(&ws->tile_start, (size_t)(B + 1) * sizeof(int32_t)));
    ws->b_cap = B;
  }
  if (ws->tiles_cap < max_tiles) {
    cudaFree(ws->tiles); ws->tiles = nullptr; ws->tiles_cap = 0;
    NANN_CUDA(cudaMalloc(&ws->tiles, (size_t)max_tiles * sizeof(int2)));
    ws->tiles_cap = max_tiles;
  }
  return NANN_OK;
}

static nann_status mlp_tc_prepare(nann_scorer* s) {
  if (s->tc) return NANN_OK;
  NANN_CUDA(cudaSetDevice(s->device));
  cudaDeviceProp pr;
  NANN_CUDA(cudaGetDeviceProperties(&pr, s->device));
  if (pr.major != 10) return fail(NANN_FAILED_PRECONDITION, "tcgen05 path needs sm_100 (device is sm_%d%d)", pr.major, pr.minor);
  auto* st = new MlpTcState();
  st->n_ctas = pr.multiProcessorCount;
  // W1 (x half) / W2 row-major are recovered from the k-major copies the exact path keeps
  DevBuf<float> W1, W2, tmp;
  nann_status rc = W1.alloc(512 * 256);
  if (rc == NANN_OK) rc = W2.alloc(512 * 512);
  if (rc == NANN_OK) rc = tmp.alloc(512 * 128);
  if (rc == NANN_OK && cudaMalloc(&st->W8img, 2 * T8_IMG_BYTES_PER_RANK) != cudaSuccess) {
    cudaGetLastError();
    rc = fail(NANN_RESOURCE_EXHAUSTED, "OOM for tensor-core scorer state");
  }
  if (rc != NANN_OK) { delete st; return rc; }
  auto bail = [&](nann_status r) { cudaFree(st->W8img); delete st; return r; };
  if (cudaMemsetAsync(W1.d, 0, 512 * 256 * 4, 0) != cudaSuccess) return bail(fail(NANN_INTERNAL, "memset failed"));
  // transpose_kernel(in [rows][ld], out [cols][rows]): W1xT [128][512] -> temp [512][128] -> W1[:,128:256] (ld 256)
  NANN_LAUNCH_OR(bail, transpose_kernel, (unsigned)ceil_div(128 * 512, 256), 256, 0, 0, s->W1xT, 128, 512, 512, 0, tmp.d);
  if (cudaMemcpy2DAsync(W1.d + 128, 256 * 4, tmp.d, 128 * 4, 128 * 4, 512, cudaMemcpyDeviceToDevice, 0) != cudaSuccess)
    return bail(fail(NANN_INTERNAL, "weight staging failed"));
  NANN_LAUNCH_OR(bail, transpose_kernel, (unsigned)ceil_div(512 * 512, 256), 256, 0, 0, s->W2T, 512, 512, 512, 0, W2.d);
  NANN_LAUNCH_OR(bail, tc_build_w8_kernel, (2 * T8_UNITS * 256 * 64) / 256, 256, 0, 0, W1.d, W2.d, st->W8img);
  if (cudaFuncSetAttribute(mlp_tc8_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, T8_SMEM_BYTES) != cudaSuccess ||
      cudaFuncSetAttribute(mlp_tc8_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, T8_SMEM_BYTES) != cudaSuccess ||
      cudaMemcpy(st->h_b2, s->b2, sizeof(st->h_b2), cudaMemcpyDeviceToHost) != cudaSuccess ||
      cudaMemcpy(st->h_w3, s->w3, sizeof(st->h_w3), cudaMemcpyDeviceToHost) != cudaSuccess ||
      cudaDeviceSynchronize() != cudaSuccess)
    return bail(fail(NANN_INTERNAL, "tensor-core scorer setup failed: %s", cudaGetErrorString(cudaGetLastError())));
  s->tc = st;
  return NANN_OK;
}

static long long* g_tc_trace = nullptr;   // set by nann_debug_tc_trace (debug builds of the timeline only)

static nann_status mlp_tc_score(nann_scorer* s, const ScoreCall& c, cudaStream_t stm) {
  auto* st = (MlpTcState*)s->tc;
  if (!st) return fail(NANN_FAILED_PRECONDITION, "tensor-core scorer not prepared");
  if (!c.ws) return fail(NANN_INTERNAL, "tensor-core scorer needs a workspace");
  MlpTcArgs a{};
  a.table = c.table; a.ids = c.ids; a.ids_stride = c.ids_stride; a.rows_stride = c.rows_stride;
  a.n_ptr = c.n_ptr; a.n_fixed = c.n_fixed; a.B = c.B;
  a.hu = c.hu; a.W8img = st->W8img;
  a.out = c.out; a.out_stride = c.out_stride; a.status = c.status;
  memcpy(a.b2c, st->h_b2, sizeof(a.b2c));
  memcpy(a.w3c, st->h_w3, sizeof(a.w3c));
  const int64_t n_tiles = (int64_t)c.B * ceil_div(c.max_n, TC_M);
  NANN_TRY(tc_ws_ensure(c.ws, c.B, n_tiles));
  a.tiles = c.ws->tiles; a.tile_total = c.ws->tile_start + c.B;
  a.trace = g_tc_trace;
  static const int cta_cap = [] { const char* e = std::getenv("NANN_TC_CTAS"); return e ? atoi(e) : 1 << 30; }();   // debug
  // a dense tile list gives every cluster the same number of tiles however ragged the per-query counts are
  NANN_LAUNCH(tile_scan_kernel, 1, 1024, 0, stm, c.n_ptr, c.n_fixed, c.status, c.B, c.ws->tile_start);
  // cluster pairs, both CTAs on the same tile; the two CTAs red.add their partial scores into out (zeroed here)
  NANN_LAUNCH(tile_fill_kernel, c.B, 256, 0, stm, c.ws->tile_start, c.B, c.ws->tiles, c.n_ptr, c.n_fixed, c.out, c.out_stride);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(std::min(st->n_ctas, cta_cap) / 2 * 2));
  cfg.blockDim = dim3(T8_THREADS);
  cfg.dynamicSmemBytes = T8_SMEM_BYTES;
  cfg.stream = stm;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  if (a.trace) NANN_CUDA(cudaLaunchKernelEx(&cfg, mlp_tc8_kernel<true>, a));    // debug timeline (scripts/tc_timeline.py)
  else         NANN_CUDA(cudaLaunchKernelEx(&cfg, mlp_tc8_kernel<false>, a));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return NANN_OK;
}

static void mlp_tc_release(nann_scorer* s) {
  auto* st = (MlpTcState*)s->tc;
  if (!st) return;
  cudaFree(st->W8img);
  delete st;
  s->tc = nullptr;
}

}  // namespace nann
extern "C" nann_status nann_debug_tc_trace(long long* device_buffer_64x48) {
  nann::g_tc_trace = device_buffer_64x48;
  return NANN_OK;
}
