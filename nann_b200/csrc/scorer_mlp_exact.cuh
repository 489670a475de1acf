// scorer_mlp_exact.cuh -- fused GatherV2 + "mlp2x512" scorer in fp32 FFMA (NANN_SCORER_EXACT).
//
// Replaces, per scoring round, GatherV2 (build_opt_graph.py:92) + the H2D copy + BlazeXlaOp's
// nested XLA session + the D2H copy (blaze_xla_predictor.cc:179-225,425-426,317-358).
//
// Numerics: every output element is ONE sequential chain of fmaf in increasing k that starts
// from the bias -- the fp32 definition fixed in DESIGN.md (the CPU checker states the same one), so scores are
// bit-identical to that definition.  The per-query half of layer 1 (W1[:, :d] . u + b1) is the prefix of
// that chain and is hoisted into hu[q][H] by mlp_hoist_kernel.
//
// Tiling: one CTA = 64 candidates of one query; x tile [64][128] and h1 [64][512] stay in shared
// memory (h1 never touches HBM), weights stream from L2 in [16 k][128 neuron] chunks through a
// cp.async double buffer; 256 threads, each a 4-candidate x 8-neuron register tile.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nann {

constexpr int MLP_D = 128;
constexpr int MLP_H = 512;
constexpr int EX_TM = 64;            // candidates per CTA
constexpr int EX_TN = 128;           // neurons per pass
constexpr int EX_KC = 16;            // k rows per weight chunk
constexpr int EX_THREADS = 256;
constexpr int EX_XP = MLP_D + 4;     // x tile pitch (floats)
constexpr int EX_HP = MLP_H + 4;     // h1 tile pitch (floats)
constexpr int EX_SMEM_FLOATS = EX_TM * EX_XP + EX_TM * EX_HP + 2 * EX_KC * EX_TN;
constexpr int EX_SMEM_BYTES = EX_SMEM_FLOATS * 4;  // 182272

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

// hu[q][j] = b1[j] (+) sum_{k<d} W1uT[k][j] * u[q][k]      (sequential in k)
__global__ void __launch_bounds__(MLP_H)
mlp_hoist_kernel(const float* __restrict__ users, const float* __restrict__ W1uT,
                 const float* __restrict__ b1, float* __restrict__ hu, int B) {
  __shared__ float su[MLP_D];
  const int q = blockIdx.x, j = threadIdx.x;
  if (j < MLP_D) su[j] = users[(int64_t)q * MLP_D + j];
  __syncthreads();
  float a = b1[j];
#pragma unroll 8
  for (int k = 0; k < MLP_D; ++k) a = fmaf(W1uT[(int64_t)k * MLP_H + j], su[k], a);
  hu[(int64_t)q * MLP_H + j] = a;
}

// One pass: acc[4 cand][8 neurons] over K, A = smem [EX_TM][pitch] row-major, W = global WT[K][H]
// columns [n0, n0+128).  Thread (tm, tn): candidates tm*4..+3, neurons n0 + tn*4..+3 and
// n0 + 64 + tn*4..+3.
template <int K, int PITCH>
__device__ __forceinline__ void exact_pass(const float* __restrict__ A, const float* __restrict__ WT,
                                           int n0, float* __restrict__ wbuf, float (&acc)[4][8],
                                           int tid) {
  const int tm = tid >> 4, tn = tid & 15;
  constexpr int NCH = K / EX_KC;
  // chunk = 16 rows x 128 floats = 512 x 16 B -> 2 cp.async per thread
  auto load_chunk = [&](int c, int buf) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int v = tid + r * EX_THREADS;       // 0..511
      const int row = v >> 5, col4 = v & 31;    // 32 x 16 B per row
      cp_async16(wbuf + buf * (EX_KC * EX_TN) + row * EX_TN + col4 * 4,
                 WT + (int64_t)(c * EX_KC + row) * MLP_H + n0 + col4 * 4);
    }
    cp_async_commit();
  };
  load_chunk(0, 0);
  for (int c = 0; c < NCH; ++c) {
    if (c + 1 < NCH) { load_chunk(c + 1, (c + 1) & 1); cp_async_wait<1>(); }
    else             { cp_async_wait<0>(); }
    __syncthreads();
    const float* wb = wbuf + (c & 1) * (EX_KC * EX_TN);
    const float* a0 = A + (tm * 4) * PITCH + c * EX_KC;
#pragma unroll
    for (int kk = 0; kk < EX_KC; kk += 4) {
      float4 av[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = *reinterpret_cast<const float4*>(a0 + i * PITCH + kk);
#pragma unroll
      for (int k4 = 0; k4 < 4; ++k4) {
        const float4 b0 = *reinterpret_cast<const float4*>(wb + (kk + k4) * EX_TN + tn * 4);
        const float4 b1 = *reinterpret_cast<const float4*>(wb + (kk + k4) * EX_TN + 64 + tn * 4);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float a = k4 == 0 ? av[i].x : k4 == 1 ? av[i].y : k4 == 2 ? av[i].z : av[i].w;
          acc[i][0] = fmaf(b0.x, a, acc[i][0]);
          acc[i][1] = fmaf(b0.y, a, acc[i][1]);
          acc[i][2] = fmaf(b0.z, a, acc[i][2]);
          acc[i][3] = fmaf(b0.w, a, acc[i][3]);
          acc[i][4] = fmaf(b1.x, a, acc[i][4]);
          acc[i][5] = fmaf(b1.y, a, acc[i][5]);
          acc[i][6] = fmaf(b1.z, a, acc[i][6]);
          acc[i][7] = fmaf(b1.w, a, acc[i][7]);
        }
      }
    }
    __syncthreads();  // chunk buffer (c&1) is refilled by the load issued in iteration c+1
  }
}

struct MlpExactArgs {
  const float* table;       // [n_rows][128] f32
  const int32_t* ids;       // per query at ids + q*ids_stride (ids_stride 0 = shared list);
                            // NULL: rows are table + (q*rows_stride + i)*128 (dense item_emb input)
  int64_t ids_stride;
  int64_t rows_stride;
  const int32_t* n_ptr;     // per query count (nullable)
  int n_fixed;
  const float* hu;          // [B][512]
  const float* W1xT;        // [128][512]
  const float* W2T;         // [512][512]
  const float* b2;          // [512]
  const float* w3;          // [512]
  float* out;               // scores at out + q*out_stride
  int64_t out_stride;
  const int32_t* status;    // nullable
};

__global__ void __launch_bounds__(EX_THREADS, 1)
mlp_exact_kernel(MlpExactArgs p) {
  extern __shared__ __align__(16) float smem[];
  float* xs = smem;                        // [64][132]  (later: h2 block [64][132])
  float* h1 = smem + EX_TM * EX_XP;        // [64][516]
  float* wbuf = h1 + EX_TM * EX_HP;        // [2][16][128]
  const int q = blockIdx.y;
  const int tid = threadIdx.x;
  if (p.status && p.status[q] != 0) return;
  const int n = p.n_ptr ? p.n_ptr[q] : p.n_fixed;
  const int t0 = blockIdx.x * EX_TM;
  if (t0 >= n) return;
  const int nt = min(EX_TM, n - t0);

  // ---- gather: warp w owns rows w*8 .. w*8+7; one float4 per lane = one 512-B row per warp request;
  // ids come from one coalesced load and all 8 row loads are in flight before the first store
  {
    const int warp = tid >> 5, lane = tid & 31;
    const int my_r = warp * 8 + (lane & 7);
    const int my_cc = my_r < nt ? my_r : 0;   // pad with the tile's first row (scores not written)
    const long long my_row_idx = p.ids ? (long long)p.ids[(int64_t)q * p.ids_stride + t0 + my_cc]
                                       : ((long long)q * p.rows_stride + t0 + my_cc);
    float4 v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const long long ridx = __shfl_sync(0xffffffffu, my_row_idx, j);
      const float* src = p.table + ridx * MLP_D + lane * 4;
      asm("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
          : "=f"(v[j].x), "=f"(v[j].y), "=f"(v[j].z), "=f"(v[j].w) : "l"(src));
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) *reinterpret_cast<float4*>(xs + (warp * 8 + j) * EX_XP + lane * 4) = v[j];
  }
  __syncthreads();

  const int tm = tid >> 4, tn = tid & 15;
  float acc[4][8];
  // ---- layer 1: h1[c][j] = relu(hu[j] (+) sum_k W1xT[k][j] x[c][k])
  const float* huq = p.hu + (int64_t)q * MLP_H;
  for (int pass = 0; pass < MLP_H / EX_TN; ++pass) {
    const int n0 = pass * EX_TN;
    const float4 i0 = *reinterpret_cast<const float4*>(huq + n0 + tn * 4);
    const float4 i1 = *reinterpret_cast<const float4*>(huq + n0 + 64 + tn * 4);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      acc[i][0] = i0.x; acc[i][1] = i0.y; acc[i][2] = i0.z; acc[i][3] = i0.w;
      acc[i][4] = i1.x; acc[i][5] = i1.y; acc[i][6] = i1.z; acc[i][7] = i1.w;
    }
    exact_pass<MLP_D, EX_XP>(xs, p.W1xT, n0, wbuf, acc, tid);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float4 o0, o1;
      o0.x = acc[i][0] > 0.f ? acc[i][0] : 0.f; o0.y = acc[i][1] > 0.f ? acc[i][1] : 0.f;
      o0.z = acc[i][2] > 0.f ? acc[i][2] : 0.f; o0.w = acc[i][3] > 0.f ? acc[i][3] : 0.f;
      o1.x = acc[i][4] > 0.f ? acc[i][4] : 0.f; o1.y = acc[i][5] > 0.f ? acc[i][5] : 0.f;
      o1.z = acc[i][6] > 0.f ? acc[i][6] : 0.f; o1.w = acc[i][7] > 0.f ? acc[i][7] : 0.f;
      float* dst = h1 + (tm * 4 + i) * EX_HP + n0 + tn * 4;
      *reinterpret_cast<float4*>(dst) = o0;
      *reinterpret_cast<float4*>(dst + 64) = o1;
    }
  }
  __syncthreads();

  // ---- layer 2 + final dot: s[c] = sum_j w3[j] * relu(b2[j] (+) sum_k W2T[k][j] h1[c][k]), j ascending
  float s = 0.f;  // threads 0..63 own one candidate's chain
  float* h2 = xs; // [64][132]
  for (int pass = 0; pass < MLP_H / EX_TN; ++pass) {
    const int n0 = pass * EX_TN;
    const float4 i0 = *reinterpret_cast<const float4*>(p.b2 + n0 + tn * 4);
    const float4 i1 = *reinterpret_cast<const float4*>(p.b2 + n0 + 64 + tn * 4);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      acc[i][0] = i0.x; acc[i][1] = i0.y; acc[i][2] = i0.z; acc[i][3] = i0.w;
      acc[i][4] = i1.x; acc[i][5] = i1.y; acc[i][6] = i1.z; acc[i][7] = i1.w;
    }
    exact_pass<MLP_H, EX_HP>(h1, p.W2T, n0, wbuf, acc, tid);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float4 o0, o1;
      o0.x = acc[i][0] > 0.f ? acc[i][0] : 0.f; o0.y = acc[i][1] > 0.f ? acc[i][1] : 0.f;
      o0.z = acc[i][2] > 0.f ? acc[i][2] : 0.f; o0.w = acc[i][3] > 0.f ? acc[i][3] : 0.f;
      o1.x = acc[i][4] > 0.f ? acc[i][4] : 0.f; o1.y = acc[i][5] > 0.f ? acc[i][5] : 0.f;
      o1.z = acc[i][6] > 0.f ? acc[i][6] : 0.f; o1.w = acc[i][7] > 0.f ? acc[i][7] : 0.f;
      float* dst = h2 + (tm * 4 + i) * EX_XP + tn * 4;
      *reinterpret_cast<float4*>(dst) = o0;
      *reinterpret_cast<float4*>(dst + 64) = o1;
    }
    __syncthreads();
    if (tid < EX_TM) {
      const float* hrow = h2 + tid * EX_XP;
      const float* w = p.w3 + n0;
#pragma unroll 8
      for (int j = 0; j < EX_TN; ++j) s = fmaf(w[j], hrow[j], s);
    }
    __syncthreads();
  }
  if (tid < nt) p.out[(int64_t)q * p.out_stride + t0 + tid] = s;
}

}  // namespace nann
