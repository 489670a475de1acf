// scorer_tc_common.cuh -- shared pieces of the tcgen05 scorer: constants, PTX wrappers (mbarrier, tcgen05, bulk copies),
// the fp16 hi/lo split helpers, the kernel argument block and the dense tile list kernels.
//
// Numerics of NANN_SCORER_TENSOR: every fp32 operand v is split v = hi + lo (hi = fp16(v), lo = fp16(v - hi)) and every
// product is issued as three kind::f16 MMAs  Ah*Bh + Al*Bh + Ah*Bl  into one fp32 TMEM accumulator (the dropped Al*Bl
// term is 2^-22 relative).  |score - exact| <= 1e-5 is the contract.
#pragma once
#include <cuda_fp16.h>

namespace nann {

constexpr int TC_M = 128;             // candidates per tile (UMMA M)
constexpr int TC_SLAB_BYTES = TC_M * 128;      // [128 rows][64 fp16] = 16 KB
constexpr int T2_EPI_WARPS = 8;                       // gather + epilogue warps: two per TMEM lane quarter
constexpr int T2_EPI_THREADS = T2_EPI_WARPS * 32;
constexpr int T2_STAGE = 32768;                       // one ring stage: [hi 16 KB][lo 16 KB] A slab or half a weight unit

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
// Never hang the GPU: a wait that lasts longer than NANN_MBAR_WATCHDOG_NS of wall clock (a lost arrival, i.e. a
// pipeline bug) turns into a trap -> a CUDA error on the host instead of a dead device.  10 s by default: long enough
// for a debugger single-step, an MPS time slice or compute-sanitizer's instrumentation (racecheck slows these kernels
// ~100x; build with -DNANN_MBAR_WATCHDOG_NS=0 to compile the watchdog out).
#ifndef NANN_MBAR_WATCHDOG_NS
#define NANN_MBAR_WATCHDOG_NS 10000000000ull
#endif
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
      else if (NANN_MBAR_WATCHDOG_NS && now - t0 > NANN_MBAR_WATCHDOG_NS) __trap();
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
        else if (NANN_MBAR_WATCHDOG_NS && now - t0 > NANN_MBAR_WATCHDOG_NS) __trap();
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

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// 1-D bulk copy global -> shared, completion counted on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void bulk_g2s_mc(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar, uint16_t mask) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
               ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void tc_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask) : "memory");
}

struct MlpTcArgs {
  const float* table; const int32_t* ids; int64_t ids_stride; int64_t rows_stride;
  const int32_t* n_ptr; int n_fixed; int tiles_per_q; int B;
  const float* hu;        // [B][512]
  const __half* W1img; const __half* W2img; const __half* W8img;
  const float* b2; const float* w3;
  uint8_t* scratch;       // [gridDim.x][TC_SCRATCH_BYTES]
  float* out; int64_t out_stride; const int32_t* status;
  const int2* tiles; const int32_t* tile_total;   // dense tile list (tile_scan_kernel + tile_fill_kernel)
  long long* trace;                               // optional CTA-0 timeline [64 tiles][48 events] (debug)
  // b2 / w3 by value: kernel parameters live in the constant bank, so the layer-2 epilogue reads them with
  // uniform constant loads.  (With ~225 KB of shared memory per CTA the L1 data cache is a few KB: __ldg of
  // these vectors missed to L2 on almost every access and throttled the epilogues.)
  int mma_gap;                                    // debug (NANN_TC_GAP): cycles between phase-2 MMA issues
  alignas(16) float b2c[MLP_H];
  alignas(16) float w3c[MLP_H];
};

// ---- tile list: tiles[g] = (q, t) for every 128-row tile of every live query ------------------------
__global__ void __launch_bounds__(1024)
tile_scan_kernel(const int32_t* __restrict__ n_ptr, int n_fixed, const int32_t* __restrict__ status, int B,
                 int32_t* __restrict__ tile_start /* [B+1] */) {
  __shared__ int warp_sum[32];
  __shared__ int carry;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  if (tid == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < B; base += 1024) {
    const int q = base + tid;
    int c = 0;
    if (q < B && !(status && status[q] != 0)) {
      const int n = n_ptr ? n_ptr[q] : n_fixed;
      c = n > 0 ? (n + TC_M - 1) / TC_M : 0;
    }
    int incl = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
    if (lane == 31) warp_sum[w] = incl;
    __syncthreads();
    if (w == 0) {
      int s = warp_sum[lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { int t = __shfl_up_sync(0xffffffffu, s, d); if (lane >= d) s += t; }
      warp_sum[lane] = s;
    }
    __syncthreads();
    const int excl = carry + (w > 0 ? warp_sum[w - 1] : 0) + incl - c;
    if (q < B) tile_start[q] = excl;
    __syncthreads();
    if (tid == 1023) carry = excl + c;
    __syncthreads();
  }
  if (tid == 0) tile_start[B] = carry;
}
// also zeroes the query's score row: the two CTAs of a cluster red.add their partial scores into it
__global__ void tile_fill_kernel(const int32_t* __restrict__ tile_start, int B, int2* __restrict__ tiles,
                                 const int32_t* __restrict__ n_ptr, int n_fixed, float* __restrict__ out, int64_t out_stride) {
  const int q = blockIdx.x;
  const int s = tile_start[q], e = tile_start[q + 1];
  for (int t = threadIdx.x; t < e - s; t += blockDim.x) tiles[s + t] = make_int2(q, t);
  if (e == s) return;                              // failed or empty query: no tile, nothing is added
  const int n = n_ptr ? n_ptr[q] : n_fixed;
  float* row = out + (int64_t)q * out_stride;
  for (int i = threadIdx.x; i < n; i += blockDim.x) row[i] = 0.f;
}


}  // namespace nann
