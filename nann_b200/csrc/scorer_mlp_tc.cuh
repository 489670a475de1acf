// scorer_mlp_tc.cuh -- NANN_SCORER_TENSOR: fused row gather + "mlp2x512" scorer on the 5th-gen
// tensor cores (tcgen05.mma, accumulators in TMEM), fp32-grade accuracy from fp16 operands:
//   every fp32 operand v is split v = hi + lo (hi = fp16(v), lo = fp16(v - hi)) and every product is
//   issued as three kind::f16 MMAs  Ah*Bh + Ah*Bl + Al*Bh  into one fp32 TMEM accumulator
//   (the dropped Al*Bl term is 2^-22 relative).  |score - exact| <= 1e-5 is the contract.
//
// One persistent CTA per SM; a work item is a tile of 128 candidates of one query:
//   gather   rows -> registers -> (hi,lo) fp16 -> shared, K-major 128B-swizzled (UMMA canonical)
//   phase 1  h1 = relu(x.W1x + hu): 4 neuron chunks of 128, D1 double-buffered in TMEM (2x128 cols),
//            epilogue: tcgen05.ld -> +hu, relu, split -> this CTA's 256 KB h1 scratch (stays in L2)
//   phase 2  D2[128x512] (all 512 TMEM columns) += h1.W2^T over 8 K-slabs x 2 neuron halves,
//            epilogue: s = sum_j w3[j]*relu(D2[:,j]+b2[j])
// Operands stream with cp.async into a double buffer while the previous stage's MMAs run; MMA
// completion is tracked with tcgen05.commit -> mbarrier.  Weight planes are pre-swizzled images
// (built once by mlp_tc_prepare) so a stage is one contiguous 64 KB copy.
#pragma once
#include <cuda_fp16.h>

namespace nann {

constexpr int TC_M = 128;             // candidates per tile (UMMA M)
constexpr int TC_THREADS = 256;
constexpr int TC_A_BYTES = 65536;     // x tile (2 planes x 2 slabs x 16 KB) / phase-2 A double buffer (2 x 32 KB)
constexpr int TC_B_BYTES = 65536;     // one weight stage (2 planes)
constexpr int TC_SMEM_BYTES = TC_A_BYTES + 2 * TC_B_BYTES + 1024 /*align*/ + 1024 /*barriers, tmem slot, partials*/;
constexpr int TC_SLAB_BYTES = TC_M * 128;      // [128 rows][64 fp16] = 16 KB
constexpr int TC_SCRATCH_BYTES = 8 * 2 * TC_SLAB_BYTES;  // h1: 8 K-slabs x (hi,lo) = 256 KB per CTA

struct MlpTcState {
  float h_b2[MLP_H], h_w3[MLP_H];   // host copies for the by-value kernel parameters
  __half* W1img = nullptr;   // [4 chunks][2 planes][2 slabs][128 rows][64]   swizzled, 256 KB
  __half* W2img = nullptr;   // [8 slabs][2 halves][2 planes][256 rows][64]   swizzled, 1 MB
  __half* W8img = nullptr;    // mlp_tc8_kernel: [rank 2][unit 10][hi 32 KB | lo 32 KB], 1.25 MB (scorer_mlp_tc8.cuh)
  int n_ctas = 0;
};

// Per-caller scratch of the tensor-core scorer (one per searcher / per op call, i.e. per stream:
// concurrent launches must not share it): the CTAs' h1 slabs and the round's tile list.
struct TcWorkspace {
  uint8_t* scratch = nullptr;      // [n_ctas][TC_SCRATCH_BYTES]
  int n_ctas = 0;
  int32_t* tile_start = nullptr;   // [b_cap + 1]
  int b_cap = 0;
  int2* tiles = nullptr;           // [tiles_cap]
  int64_t tiles_cap = 0;
};
static void tc_ws_free(TcWorkspace* ws) {
  if (!ws) return;
  cudaFree(ws->scratch); cudaFree(ws->tile_start); cudaFree(ws->tiles);
  *ws = TcWorkspace();
}
static nann_status tc_ws_ensure(TcWorkspace* ws, int n_ctas, int B, int64_t max_tiles) {
  if (ws->n_ctas < n_ctas) {
    cudaFree(ws->scratch); ws->scratch = nullptr; ws->n_ctas = 0;
    NANN_CUDA(cudaMalloc(&ws->scratch, (size_t)n_ctas * TC_SCRATCH_BYTES));
    ws->n_ctas = n_ctas;
  }
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

// ---- PTX wrappers -------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Never hang the GPU: a wait that lasts longer than 2 s of wall clock (a lost arrival, i.e. a
// pipeline bug) turns into a trap -> a CUDA error on the host instead of a dead device.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  unsigned long long t0 = 0;
  for (uint32_t spin = 0; !done; ++spin) {
#ifdef NANN_MBAR_HINT_NS
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar), "r"(parity), "r"((uint32_t)NANN_MBAR_HINT_NS) : "memory");
#else
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
#endif
    if (!done && (spin & 63) == 63) {
      const unsigned long long now = global_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 2000000000ull) __trap();
    }
  }
}
// One lane of a converged warp (CUTLASS's elect_one_sync).  Guarding tcgen05.mma / commit with this inside
// warp-uniform control flow lets ptxas keep descriptors in uniform registers; an `if (lane == 0)` region made
// it wrap every UTCHMMA in an ELECT / R2UR.BROADCAST / BRA.U.ANY waterfall loop (~15 instructions per MMA).
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred;
}
// lane 0 polls, the warp reconverges: keeps the caller's control flow warp-uniform
__device__ __forceinline__ void mbar_wait_warp(uint32_t bar, uint32_t parity) {
  if ((threadIdx.x & 31) == 0) mbar_wait(bar, parity);
  __syncwarp();
}
// Same, for waiters that are not on the critical path: sleeps between polls so that a spinning warp does not
// take issue slots from the MMA issuer on its scheduler (the arbiter favours higher warp ids).
#ifndef NANN_MBAR_SLEEP_NS
#define NANN_MBAR_SLEEP_NS 64
#endif
__device__ __forceinline__ void mbar_wait_warp_relaxed(uint32_t bar, uint32_t parity) {
  if ((threadIdx.x & 31) == 0) {
    uint32_t done = 0;
    unsigned long long t0 = 0;
    for (uint32_t spin = 0;; ++spin) {
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(done) : "r"(bar), "r"(parity) : "memory");
      if (done) break;
      __nanosleep(NANN_MBAR_SLEEP_NS);
      if ((spin & 63) == 63) {
        const unsigned long long now = global_ns();
        if (t0 == 0) t0 = now;
        else if (now - t0 > 2000000000ull) __trap();
      }
    }
  }
  __syncwarp();
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], kind::f16, cta_group::1
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_ld_wait_dep(uint32_t (&v)[32]);
// 32 lanes x 32 columns of fp32: thread i of the warp gets row (lane base + i)
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr) : "memory");
  tc_ld_wait_dep(v);
}

// the same load without the wait: lets the next block's TMEM read overlap the math on the current one
__device__ __forceinline__ void tc_ld32_nowait(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// wait + a register dependency on the loaded block, so no use of v[] can be scheduled above the wait
__device__ __forceinline__ void tc_ld_wait_dep(uint32_t (&v)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                 "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]),
                 "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]),
                 "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
               :: "memory");
}

// UMMA shared-memory descriptor, K-major, SWIZZLE_128B: rows of 128 B, 8-row atoms 1024 B apart
// (cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48),
//  layout_type=2 (SWIZZLE_128B) [61,64)).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): c_format=F32 (1)<<4, a/b_format=F16 (0),
// a/b K-major (0), n_dim = N>>3 at [17,23), m_dim = M>>4 at [24,29)
__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// byte offset of the 16-B chunk `c` (0..7) of row r inside a [rows][64 fp16] SW128 K-major tile
__host__ __device__ __forceinline__ uint32_t sw128_chunk_off(int r, int c) {
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4));
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// 16-byte streaming load of a table row chunk; NOT volatile so several can be put in flight
__device__ __forceinline__ float4 ld_row16(const float* p) {
  float4 v;
  asm("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
      : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

__device__ __forceinline__ void split_f16(float a, __half& hi, __half& lo) {
  hi = __float2half_rn(a);
  lo = __float2half_rn(a - __half2float(hi));
}
// Blackwell packed fp32 pairs (FADD2 / FFMA2): two IEEE fp32 operations per issued instruction
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  float2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(*reinterpret_cast<unsigned long long*>(&d))
      : "l"(*reinterpret_cast<const unsigned long long*>(&a)), "l"(*reinterpret_cast<const unsigned long long*>(&b)));
  return d;
}
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(*reinterpret_cast<unsigned long long*>(&d))
      : "l"(*reinterpret_cast<const unsigned long long*>(&a)), "l"(*reinterpret_cast<const unsigned long long*>(&b)),
        "l"(*reinterpret_cast<const unsigned long long*>(&c)));
  return d;
}
// relu(x + bias) for a pair, then the (hi, lo) fp16 split of both values
__device__ __forceinline__ void bias_relu_split2(uint32_t r0, uint32_t r1, float2 bias, uint32_t& hi, uint32_t& lo) {
  float2 a = add2(make_float2(__uint_as_float(r0), __uint_as_float(r1)), bias);
  a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f);
  const __half2 h = __floats2half2_rn(a.x, a.y);
  const float2 hf = __half22float2(h);
  const float2 d = fma2(hf, make_float2(-1.f, -1.f), a);      // a - hf, exact
  const __half2 l = __floats2half2_rn(d.x, d.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

// Two values at once with the PACKED converts (cvt.rn.f16x2.f32 -> F2FP.PACK_AB, ALU rate); the scalar
// F2F.F16.F32 above runs on the 16-lane/SM conversion pipe and was THE bottleneck of every epilogue
// (profiles/r01_tc_timeline_v5.log: 4.3k cycles per 64-neuron chunk for 1.4k cycles of MMA).
// Same results bit for bit: both are round-to-nearest-even converts of the same fp32 values.
__device__ __forceinline__ void split2_f16(float a0, float a1, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(a0, a1);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(a0 - hf.x, a1 - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

// contiguous global -> shared copy of `bytes` (multiple of 16*TC_THREADS) with cp.async
__device__ __forceinline__ void tc_copy_async(uint8_t* dst, const uint8_t* __restrict__ src, int bytes, int tid) {
  for (int o = tid * 16; o < bytes; o += TC_THREADS * 16) cp_async16(dst + o, src + o);
}

struct MlpTcArgs {
  const float* table; const int32_t* ids; int64_t ids_stride; int64_t rows_stride;
  const int32_t* n_ptr; int n_fixed; int tiles_per_q; int B;
  const float* hu;        // [B][512]
  const __half* W1img; const __half* W2img; const __half* W8img;
  const float* b2; const float* w3;
  uint8_t* scratch;       // [gridDim.x][TC_SCRATCH_BYTES]
  float* out; int64_t out_stride; const int32_t* status;
  const int2* tiles; const int32_t* tile_total;   // dense tile list (mlp_tc3_kernel)
  long long* trace;                               // optional CTA-0 timeline [64 tiles][48 events] (debug)
  // b2 / w3 by value: kernel parameters live in the constant bank, so the layer-2 epilogue reads them with
  // uniform constant loads.  (With ~225 KB of shared memory per CTA the L1 data cache is a few KB: __ldg of
  // these vectors missed to L2 on almost every access and throttled the epilogues.)
  int mma_gap;                                    // debug (NANN_TC_GAP): cycles between phase-2 MMA issues
  alignas(16) float b2c[MLP_H];
  alignas(16) float w3c[MLP_H];
};

__global__ void __launch_bounds__(TC_THREADS, 1)
mlp_tc_kernel(MlpTcArgs p) {
  extern __shared__ uint8_t tc_smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)tc_smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                         // 64 KB
  uint8_t* sB = smem + TC_A_BYTES;            // 2 x 64 KB
  uint64_t* bars = (uint64_t*)(smem + TC_A_BYTES + 2 * TC_B_BYTES);   // 2 mbarriers
  uint32_t* tmem_slot = (uint32_t*)(bars + 2);
  float* part = (float*)(bars + 4);           // [128] partial sums of the upper column half

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar0 = smem_u32(&bars[0]), bar1 = smem_u32(&bars[1]);

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 32) {
    mbar_init(bar0, 1);
    mbar_init(bar1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  uint32_t ph0 = 0, ph1 = 0;  // parity of the next completion of bar0 / bar1 (uniform across threads)

  const uint32_t sA_u = smem_u32(sA), sB_u = smem_u32(sB);
  uint8_t* scratch = p.scratch + (size_t)blockIdx.x * TC_SCRATCH_BYTES;
  const int lane_q = warp & 3;      // TMEM lane quarter this warp may access
  const int col_half = warp >> 2;   // which half of the columns this warp's epilogue covers
  const int row = lane_q * 32 + lane;

  const int64_t n_tiles = (int64_t)p.B * p.tiles_per_q;
  for (int64_t g = blockIdx.x; g < n_tiles; g += gridDim.x) {
    const int q = (int)(g / p.tiles_per_q), t = (int)(g % p.tiles_per_q);
    if (p.status && p.status[q] != 0) continue;
    const int n = p.n_ptr ? p.n_ptr[q] : p.n_fixed;
    const int t0 = t * TC_M;
    if (t0 >= n) continue;
    const int nt = min(TC_M, n - t0);

    // ---- gather + split: x tile -> sA as [plane][slab][128 rows x 128 B swizzled]
    {  // warp w owns rows w*16 .. w*16+15: coalesced id load, 8 row loads in flight at a time
      const int my_r = warp * 16 + (lane & 15);
      const int my_cc = my_r < nt ? my_r : 0;
      const long long my_row_idx = p.ids ? (long long)p.ids[(int64_t)q * p.ids_stride + t0 + my_cc]
                                         : ((long long)q * p.rows_stride + t0 + my_cc);
#pragma unroll 1
      for (int i0 = 0; i0 < 16; i0 += 8) {
        float4 v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const long long ridx = __shfl_sync(0xffffffffu, my_row_idx, i0 + j);
          v[j] = ld_row16(p.table + ridx * MLP_D + lane * 4);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int c = warp * 16 + i0 + j;
          uint32_t h01, l01, h23, l23;
            split2_f16(v[j].x, v[j].y, h01, l01); split2_f16(v[j].z, v[j].w, h23, l23);
          const int k = lane * 4, slab = k >> 6, chunk = (k & 63) >> 3, sub = (k & 7) * 2;
          const uint32_t off = slab * TC_SLAB_BYTES + sw128_chunk_off(c, chunk) + sub;
          *reinterpret_cast<uint2*>(sA + off) = make_uint2(h01, h23);
          *reinterpret_cast<uint2*>(sA + 2 * TC_SLAB_BYTES + off) = make_uint2(l01, l23);
        }
      }
    }

    // ================================ phase 1: layer 1 ========================================
    const float* huq = p.hu + (int64_t)q * MLP_H;
    auto epilogue1 = [&](int c) {
      // D1[c&1] (128 columns): this thread's row, 64 of the columns
      const uint32_t bar = (c & 1) ? bar1 : bar0;
      uint32_t& ph = (c & 1) ? ph1 : ph0;
      mbar_wait(bar, ph); ph ^= 1;
      tc_fence_after();
#pragma unroll
      for (int part32 = 0; part32 < 2; ++part32) {
        const int col0 = col_half * 64 + part32 * 32;           // column inside the 128-chunk
        uint32_t v[32];
        tc_ld32(tmem + ((uint32_t)(lane_q * 32) << 16) + (uint32_t)((c & 1) * 128 + col0), v);
        const int neuron0 = c * 128 + col0;                      // k index of layer 2
        const int slab = neuron0 >> 6;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {                         // 4 chunks of 8 columns
          uint32_t hw[4], lw[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float a0 = __uint_as_float(v[ch * 8 + 2 * e]) + huq[neuron0 + ch * 8 + 2 * e];
            float a1 = __uint_as_float(v[ch * 8 + 2 * e + 1]) + huq[neuron0 + ch * 8 + 2 * e + 1];
            a0 = a0 > 0.f ? a0 : 0.f; a1 = a1 > 0.f ? a1 : 0.f;
            split2_f16(a0, a1, hw[e], lw[e]);
          }
          const int chunk = ((neuron0 & 63) >> 3) + ch;
          const uint32_t off = (uint32_t)slab * 2 * TC_SLAB_BYTES + sw128_chunk_off(row, chunk);
          *reinterpret_cast<uint4*>(scratch + off) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
          *reinterpret_cast<uint4*>(scratch + off + TC_SLAB_BYTES) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
        }
      }
      tc_fence_before();
    };

    for (int c = 0; c < 4; ++c) {
      uint8_t* bbuf = sB + (c & 1) * TC_B_BYTES;
      tc_copy_async(bbuf, (const uint8_t*)p.W1img + (size_t)c * TC_B_BYTES, TC_B_BYTES, tid);
      cp_async_commit();
      cp_async_wait<0>();
      fence_proxy_async();     // generic-proxy smem writes (gather, cp.async) -> visible to the tensor core
      __syncthreads();
      if (tid == 0) {
        tc_fence_after();
        const uint32_t d = tmem + (uint32_t)((c & 1) * 128);
        const uint32_t idesc = umma_idesc_f16(128, 128);
        const uint32_t b_u = sB_u + (c & 1) * TC_B_BYTES;        // [plane][slab][128 rows x 128 B]
        uint32_t acc = 0;
#pragma unroll
        for (int slab = 0; slab < 2; ++slab) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t ah = umma_desc_sw128(sA_u + slab * TC_SLAB_BYTES + ks * 32);
            const uint64_t al = umma_desc_sw128(sA_u + (2 + slab) * TC_SLAB_BYTES + ks * 32);
            const uint64_t bh = umma_desc_sw128(b_u + slab * TC_SLAB_BYTES + ks * 32);
            const uint64_t bl = umma_desc_sw128(b_u + (2 + slab) * TC_SLAB_BYTES + ks * 32);
            tc_mma_f16(d, ah, bh, idesc, acc); acc = 1;
            tc_mma_f16(d, ah, bl, idesc, 1);
            tc_mma_f16(d, al, bh, idesc, 1);
          }
        }
        tc_commit((c & 1) ? bar1 : bar0);
      }
      if (c > 0) epilogue1(c - 1);
    }
    epilogue1(3);
    __syncthreads();   // h1 scratch complete and visible to the whole CTA; x tile and D1 are free

    // ================================ phase 2: layer 2 ========================================
    for (int i = 0; i < 16; ++i) {
      const int s = i >> 1, h = i & 1;
      const uint32_t bar = (i & 1) ? bar1 : bar0;
      uint32_t& ph = (i & 1) ? ph1 : ph0;
      if (i >= 2) { mbar_wait(bar, ph); ph ^= 1; }               // MMA(i-2) done: its buffers are free
      uint8_t* bbuf = sB + (i & 1) * TC_B_BYTES;
      tc_copy_async(bbuf, (const uint8_t*)p.W2img + (size_t)i * TC_B_BYTES, TC_B_BYTES, tid);
      if (h == 0)   // A slab s (hi then lo, 32 KB contiguous in the scratch) -> sA[(s&1)]
        tc_copy_async(sA + (s & 1) * 2 * TC_SLAB_BYTES, scratch + (size_t)s * 2 * TC_SLAB_BYTES, 2 * TC_SLAB_BYTES, tid);
      cp_async_commit();
      cp_async_wait<0>();
      fence_proxy_async();
      __syncthreads();
      if (tid == 0) {
        tc_fence_after();
        const uint32_t d = tmem + (uint32_t)(h * 256);
        const uint32_t idesc = umma_idesc_f16(128, 256);
        const uint32_t a_u = sA_u + (s & 1) * 2 * TC_SLAB_BYTES;  // [hi 16 KB][lo 16 KB]
        const uint32_t b_u = sB_u + (i & 1) * TC_B_BYTES;         // [hi 32 KB][lo 32 KB], 256 rows each
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t ah = umma_desc_sw128(a_u + ks * 32);
          const uint64_t al = umma_desc_sw128(a_u + TC_SLAB_BYTES + ks * 32);
          const uint64_t bh = umma_desc_sw128(b_u + ks * 32);
          const uint64_t bl = umma_desc_sw128(b_u + 2 * TC_SLAB_BYTES + ks * 32);
          tc_mma_f16(d, ah, bh, idesc, (s > 0 || ks > 0) ? 1u : 0u);
          tc_mma_f16(d, ah, bl, idesc, 1);
          tc_mma_f16(d, al, bh, idesc, 1);
        }
        tc_commit(bar);
      }
    }
    mbar_wait(bar0, ph0); ph0 ^= 1;     // MMA(14)
    mbar_wait(bar1, ph1); ph1 ^= 1;     // MMA(15): D2 complete
    tc_fence_after();

    // ---- epilogue 2: s = sum_j w3[j] * relu(D2[row][j] + b2[j]); two column halves per row
    float acc = 0.f;
#pragma unroll 1
    for (int part32 = 0; part32 < 8; ++part32) {
      const int col0 = col_half * 256 + part32 * 32;
      uint32_t v[32];
      tc_ld32(tmem + ((uint32_t)(lane_q * 32) << 16) + (uint32_t)col0, v);
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float a = __uint_as_float(v[j]) + __ldg(p.b2 + col0 + j);
        a = a > 0.f ? a : 0.f;
        acc = fmaf(__ldg(p.w3 + col0 + j), a, acc);
      }
    }
    tc_fence_before();
    if (col_half == 1) part[row] = acc;
    __syncthreads();
    if (col_half == 0 && row < nt) p.out[(int64_t)q * p.out_stride + t0 + row] = acc + part[row];
    __syncthreads();   // part[] and TMEM are reused by the next tile
    tc_fence_after();
  }

  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

}  // namespace nann
#include "scorer_mlp_tc2.cuh"
#include "scorer_mlp_tc3.cuh"
#include "scorer_mlp_tc8.cuh"
namespace nann {

// ---- weight images ---------------------------------------------------------------------------------
// W1 [512][256] row-major (x half = columns 128..255); W2 [512][512] row-major.
__global__ void tc_build_w1_kernel(const float* __restrict__ W1, __half* __restrict__ img) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;   // one per (n, k): 512 x 128
  if (t >= 512 * 128) return;
  const int n = t >> 7, k = t & 127;
  __half hi, lo;
  split_f16(W1[n * 256 + 128 + k], hi, lo);
  const int c = n >> 7, r = n & 127, slab = k >> 6, kk = k & 63;
  const size_t base = ((size_t)c * 2 * 2 + slab) * TC_SLAB_BYTES;           // plane 0
  const size_t off = sw128_chunk_off(r, kk >> 3) + (kk & 7) * 2;
  *reinterpret_cast<__half*>((uint8_t*)img + base + off) = hi;
  *reinterpret_cast<__half*>((uint8_t*)img + base + 2 * TC_SLAB_BYTES + off) = lo;
}
__global__ void tc_build_w2_kernel(const float* __restrict__ W2, __half* __restrict__ img) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;   // 512 x 512
  if (t >= 512 * 512) return;
  const int n = t >> 9, k = t & 511;
  __half hi, lo;
  split_f16(W2[n * 512 + k], hi, lo);
  const int s = k >> 6, kk = k & 63, h = n >> 8, r = n & 255;
  const size_t base = ((size_t)(s * 2 + h)) * TC_B_BYTES;                    // stage (s,h): [hi 32 KB][lo 32 KB]
  const size_t off = sw128_chunk_off(r, kk >> 3) + (kk & 7) * 2;
  *reinterpret_cast<__half*>((uint8_t*)img + base + off) = hi;
  *reinterpret_cast<__half*>((uint8_t*)img + base + 2 * TC_SLAB_BYTES + off) = lo;
}

static nann_status mlp_tc_prepare(nann_scorer* s) {
  if (s->tc) return NANN_OK;
  NANN_CUDA(cudaSetDevice(s->device));
  auto* st = new MlpTcState();
  cudaDeviceProp pr;
  NANN_CUDA(cudaGetDeviceProperties(&pr, s->device));
  if (pr.major != 10) { delete st; return fail(NANN_FAILED_PRECONDITION, "tcgen05 path needs sm_100 (device is sm_%d%d)", pr.major, pr.minor); }
  st->n_ctas = pr.multiProcessorCount;
  // W1/W2 row-major are recovered from the k-major copies the exact path keeps
  DevBuf<float> W1, W2;
  NANN_TRY(W1.alloc(512 * 256));
  NANN_TRY(W2.alloc(512 * 512));
  NANN_CUDA(cudaMemsetAsync(W1.d, 0, 512 * 256 * 4, 0));
  // transpose_kernel(in [rows][ld], out [cols][rows]): W1xT [128][512] -> W1[:,128:256] needs out ld 256:
  // write through a temp [512][128] then strided copy
  DevBuf<float> tmp;
  NANN_TRY(tmp.alloc(512 * 128));
  NANN_LAUNCH(transpose_kernel, (unsigned)ceil_div(128 * 512, 256), 256, 0, 0, s->W1xT, 128, 512, 512, 0, tmp.d);
  NANN_CUDA(cudaMemcpy2DAsync(W1.d + 128, 256 * 4, tmp.d, 128 * 4, 128 * 4, 512, cudaMemcpyDeviceToDevice, 0));
  NANN_LAUNCH(transpose_kernel, (unsigned)ceil_div(512 * 512, 256), 256, 0, 0, s->W2T, 512, 512, 512, 0, W2.d);
  if (cudaMalloc(&st->W1img, 4 * TC_B_BYTES) != cudaSuccess || cudaMalloc(&st->W2img, 16 * TC_B_BYTES) != cudaSuccess ||
      cudaMalloc(&st->W8img, 2 * T8_IMG_BYTES_PER_RANK) != cudaSuccess) {
    cudaGetLastError();
    cudaFree(st->W1img); cudaFree(st->W2img); cudaFree(st->W8img); delete st;
    return fail(NANN_RESOURCE_EXHAUSTED, "OOM for tensor-core scorer state");
  }
  NANN_LAUNCH(tc_build_w1_kernel, (512 * 128) / 256, 256, 0, 0, W1.d, st->W1img);
  NANN_LAUNCH(tc_build_w2_kernel, (512 * 512) / 256, 256, 0, 0, W2.d, st->W2img);
  NANN_LAUNCH(tc_build_w8_kernel, (2 * T8_UNITS * 256 * 64) / 256, 256, 0, 0, W1.d, W2.d, st->W8img);
  NANN_CUDA(cudaFuncSetAttribute(mlp_tc8_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, T8_SMEM_BYTES));
  NANN_CUDA(cudaFuncSetAttribute(mlp_tc8_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, T8_SMEM_BYTES));
  NANN_CUDA(cudaFuncSetAttribute(mlp_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
  NANN_CUDA(cudaFuncSetAttribute(mlp_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, T2_SMEM_BYTES));
  NANN_CUDA(cudaFuncSetAttribute(mlp_tc3_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, T3_SMEM_BYTES));
  NANN_CUDA(cudaFuncSetAttribute(mlp_tc3_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, T3_SMEM_BYTES));
  NANN_CUDA(cudaMemcpy(st->h_b2, s->b2, sizeof(st->h_b2), cudaMemcpyDeviceToHost));
  NANN_CUDA(cudaMemcpy(st->h_w3, s->w3, sizeof(st->h_w3), cudaMemcpyDeviceToHost));
  NANN_CUDA(cudaDeviceSynchronize());
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
  a.n_ptr = c.n_ptr; a.n_fixed = c.n_fixed; a.tiles_per_q = (int)ceil_div(c.max_n, TC_M); a.B = c.B;
  a.hu = c.hu; a.W1img = st->W1img; a.W2img = st->W2img; a.W8img = st->W8img; a.b2 = s->b2; a.w3 = s->w3;
  a.out = c.out; a.out_stride = c.out_stride; a.status = c.status;
  { const char* g = getenv("NANN_TC_GAP"); a.mma_gap = g ? atoi(g) : 0; }
  memcpy(a.b2c, st->h_b2, sizeof(a.b2c));
  memcpy(a.w3c, st->h_w3, sizeof(a.w3c));
  const int64_t n_tiles = (int64_t)a.B * a.tiles_per_q;
  NANN_TRY(tc_ws_ensure(c.ws, st->n_ctas, c.B, n_tiles));
  a.scratch = c.ws->scratch; a.tiles = c.ws->tiles; a.tile_total = c.ws->tile_start + c.B;
  a.trace = g_tc_trace;
  // NANN_TC_KERNEL: 1 = bulk-synchronous cp.async kernel, 2 = warp-specialised TMA ring (h1 via L2 scratch),
  // 3 = 2 + dense tile list, 4 = 3 as 2-CTA clusters with multicast weight stages,
  // 8 = cluster-pair neuron split with on-chip h1 exchange (scorer_mlp_tc8.cuh, the default).
  // (v5/v6: layer-1 recompute with on-chip hand-off, v7: two tiles in flight through the L2 scratch -- measured
  // slower and removed; see DESIGN.md 4.1 and the git history)
  static const int version = [] { const char* e = std::getenv("NANN_TC_KERNEL"); return e ? atoi(e) : 8; }();
  static const int cta_cap = [] { const char* e = std::getenv("NANN_TC_CTAS"); return e ? atoi(e) : 1 << 30; }();   // debug
  const int grid = (int)std::min<int64_t>(std::min(st->n_ctas, cta_cap), n_tiles);
  if (version == 1) { NANN_LAUNCH(mlp_tc_kernel, grid, TC_THREADS, TC_SMEM_BYTES, stm, a); return NANN_OK; }
  if (version == 2) { NANN_LAUNCH(mlp_tc2_kernel, grid, T2_THREADS, T2_SMEM_BYTES, stm, a); return NANN_OK; }
  NANN_LAUNCH(tile_scan_kernel, 1, 1024, 0, stm, c.n_ptr, c.n_fixed, c.status, c.B, c.ws->tile_start);
  NANN_LAUNCH(tile_fill_kernel, c.B, 128, 0, stm, c.ws->tile_start, c.B, c.ws->tiles);
  if (version == 3) { NANN_LAUNCH(mlp_tc3_kernel<1>, grid, T3_THREADS, T3_SMEM_BYTES, stm, a); return NANN_OK; }
  if (version == 8) {   // cluster pairs, both CTAs on the same tile; the two CTAs red.add their partial scores into out
    NANN_CUDA(cudaMemset2DAsync(c.out, (size_t)c.out_stride * sizeof(float), 0, (size_t)c.max_n * sizeof(float), (size_t)c.B, stm));
    cudaLaunchConfig_t cfg8{};
    cfg8.gridDim = dim3((unsigned)(std::min(st->n_ctas, cta_cap) / 2 * 2));
    cfg8.blockDim = dim3(T8_THREADS);
    cfg8.dynamicSmemBytes = T8_SMEM_BYTES;
    cfg8.stream = stm;
    cudaLaunchAttribute attr8[1];
    attr8[0].id = cudaLaunchAttributeClusterDimension;
    attr8[0].val.clusterDim.x = 2; attr8[0].val.clusterDim.y = 1; attr8[0].val.clusterDim.z = 1;
    cfg8.attrs = attr8; cfg8.numAttrs = 1;
    if (a.trace) NANN_CUDA(cudaLaunchKernelEx(&cfg8, mlp_tc8_kernel<true>, a));    // debug timeline (scripts/tc_timeline.py)
    else         NANN_CUDA(cudaLaunchKernelEx(&cfg8, mlp_tc8_kernel<false>, a));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return NANN_OK;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(st->n_ctas / 2 * 2));   // whole clusters; CTAs without tiles fall through
  cfg.blockDim = dim3(T3_THREADS);
  cfg.dynamicSmemBytes = T3_SMEM_BYTES;
  cfg.stream = stm;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  NANN_CUDA(cudaLaunchKernelEx(&cfg, mlp_tc3_kernel<2>, a));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return NANN_OK;
}

}  // namespace nann
extern "C" nann_status nann_debug_tc_trace(long long* device_buffer_64x48) {
  nann::g_tc_trace = device_buffer_64x48;
  return NANN_OK;
}
namespace nann {

static void mlp_tc_release(nann_scorer* s) {
  auto* st = (MlpTcState*)s->tc;
  if (!st) return;
  cudaFree(st->W1img); cudaFree(st->W2img); cudaFree(st->W8img);
  delete st;
  s->tc = nullptr;
}

}  // namespace nann
