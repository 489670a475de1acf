// scorer_mlp_tc3.cuh -- scorer_mlp_tc2.cuh's warp-specialised tcgen05 kernel driven by a dense tile list
// and, for CL = 2, launched as thread-block clusters whose CTAs share the weight stream: each CTA
// fetches half of every 32-KB weight stage and TMA-multicasts it into both CTAs' rings
// (cp.async.bulk ... .multicast::cluster), halving the L2->SM weight traffic that bounds the kernel;
// ring slots are released by tcgen05.commit ... .multicast::cluster from both MMA threads.
// GENERATED from scorer_mlp_tc2.cuh's kernel body by a mechanical edit; keep the two in sync.
#pragma once

namespace nann {

__device__ __forceinline__ void bulk_g2s_mc(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar, uint16_t mask) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
               ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void tc_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask) : "memory");
}

struct T3Bars {           // mbarrier indices (8 B each) of mlp_tc3_kernel
  static constexpr int full = 0;        // [5] ring stage filled (TMA complete_tx)
  static constexpr int empty = 5;       // [5] ring stage consumed (tcgen05.commit)
  static constexpr int a_full = 10;     // [2] h1 slab in shared memory
  static constexpr int a_empty = 12;    // [2]
  static constexpr int x_ready = 14;
  static constexpr int d1_full = 15;    // [4] layer-1 accumulator chunk c (TMEM columns c*128..) complete
  static constexpr int d1_empty = 19;   // [4] ... drained by the epilogue
  static constexpr int h1_done = 23;    // [4]
  static constexpr int d2_full = 27;
  static constexpr int d2_empty = 28;
  static constexpr int count = 29;
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
__global__ void tile_fill_kernel(const int32_t* __restrict__ tile_start, int B, int2* __restrict__ tiles) {
  const int q = blockIdx.x;
  const int s = tile_start[q], e = tile_start[q + 1];
  for (int t = threadIdx.x; t < e - s; t += blockDim.x) tiles[s + t] = make_int2(q, t);
}

constexpr int T3_THREADS = T2_THREADS + 32;   // + one loader warp (h1 slabs: L2 scratch -> shared, cp.async)

__device__ __forceinline__ void st_global_v8(void* p, uint4 a, uint4 b) {   // one 32-byte store (STG.256)
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
}

template <int CL>
__global__ void __launch_bounds__(T3_THREADS, 1)
mlp_tc3_kernel(MlpTcArgs p) {
  // cluster rank / peers (CL == 1: a plain launch, rank 0)
  uint32_t cta_rank = 0;
  if (CL > 1) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));
  // optional timeline of CTA 0 (debug): trace[tile][event] = clock64() ; see scripts/tc_timeline.py
  long long* const trace = (blockIdx.x == 0) ? p.trace : nullptr;
  auto TR = [&](int tile_local, int ev) {
    if (trace && tile_local < 64) trace[tile_local * 48 + ev] = clock64();
  };
  extern __shared__ uint8_t tc_smem_raw[];
  // No alignment slack: the kernel has no static shared memory, so the dynamic window starts 1024-aligned.
  uint8_t* smem = tc_smem_raw;
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sX = smem;                                   // x tile, later A double buffer (2 x 32 KB)
  uint8_t* sR = smem + T2_X_BYTES;                      // ring
  uint64_t* bars = (uint64_t*)(sR + T2_NS * T2_STAGE);
  uint32_t* tmem_slot = (uint32_t*)(bars + T3Bars::count);
  static_assert(T3Bars::count < 32, "barrier block is 256 B");
  float* hu_s = (float*)(bars + 32);    // [512] this tile's hoisted layer-1 prefix (per query)
  float* part = (float*)sX;                            // [128] upper-column-half partial sums; sX is idle in epilogue 2

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar0 = smem_u32(bars);
  auto B = [&](int idx) { return bar0 + 8u * (uint32_t)idx; };

  if (tid == 0) {
    for (int i = 0; i < T2_NS; ++i) { mbar_init(B(T3Bars::full + i), 1); mbar_init(B(T3Bars::empty + i), CL); }
    for (int i = 0; i < 2; ++i) { mbar_init(B(T3Bars::a_full + i), 1); mbar_init(B(T3Bars::a_empty + i), 1); }
    for (int i = 0; i < 4; ++i) {
      mbar_init(B(T3Bars::d1_full + i), 1); mbar_init(B(T3Bars::d1_empty + i), T2_EPI_THREADS);
      mbar_init(B(T3Bars::h1_done + i), T2_EPI_THREADS);
    }
    mbar_init(B(T3Bars::x_ready), T2_EPI_THREADS);
    mbar_init(B(T3Bars::d2_full), 1);
    mbar_init(B(T3Bars::d2_empty), T2_EPI_THREADS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) {   // peers' barriers must be initialised before any remote arrive / multicast lands
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t sX_u = smem_u32(sX), sR_u = smem_u32(sR);
  uint8_t* scratch = p.scratch + (size_t)blockIdx.x * TC_SCRATCH_BYTES;
  // dense tile list built by tile_scan/tile_fill: tiles[g] = (query, tile index).  A cluster walks
  // tile GROUPS of CL tiles so that its CTAs consume the multicast weight stream in lockstep; a CTA
  // whose tile index is past the end runs a dummy tile (nt = 0, nothing written).
  const int64_t total = *p.tile_total;
  const int64_t n_tiles = (total + CL - 1) / CL * CL;
  const int64_t g_first = (int64_t)(blockIdx.x / CL) * CL + cta_rank, g_step = (int64_t)(gridDim.x / CL) * CL;
  auto tile_info = [&](int64_t g, int& q, int& t0, int& nt) -> bool {
    if (g >= total) { q = p.tiles[0].x; t0 = 0; nt = 0; return true; }
    const int2 e = p.tiles[g];
    q = e.x;
    const int n = p.n_ptr ? p.n_ptr[q] : p.n_fixed;
    t0 = e.y * TC_M;
    nt = min(TC_M, n - t0);
    return true;
  };

  if (warp == 0) {
    // =============================== producer ===============================
    {   // warp-uniform control flow; one elected lane issues the bulk copies (see elect_one())
      uint32_t it = 0;                         // ring stage counter across tiles
      auto ring_load = [&](const void* src) {
        const uint32_t slot = it % T2_NS, ph = (it / T2_NS) & 1;
        mbar_wait(B(T3Bars::empty + slot), ph ^ 1);     // released by the MMA threads of ALL CTAs of the cluster
        if (trace && lane == 0 && it >= 11 * 40 && it < 12 * 40) trace[62 * 48 + (it - 11 * 40)] = clock64();
        if (elect_one()) {
          mbar_expect_tx(B(T3Bars::full + slot), T2_STAGE);
          if (CL == 1) {
            bulk_g2s(sR_u + slot * T2_STAGE, src, T2_STAGE, B(T3Bars::full + slot));
          } else {   // this CTA fetches 1/CL of the stage and multicasts it into every CTA of the cluster
            constexpr uint32_t part_bytes = T2_STAGE / CL;
            bulk_g2s_mc(sR_u + slot * T2_STAGE + cta_rank * part_bytes, (const uint8_t*)src + cta_rank * part_bytes,
                        part_bytes, B(T3Bars::full + slot), (uint16_t)((1u << CL) - 1));
          }
        }
        ++it;
      };
      const int64_t my_tiles = (n_tiles > g_first) ? (n_tiles - g_first + g_step - 1) / g_step : 0;
      for (int64_t tl64 = 0; tl64 < my_tiles; ++tl64) {
        const int tl = (int)tl64;
        for (int c = 0; c < 4; ++c) {
          if (c == 0 && lane == 0) TR(tl, 32);
          ring_load((const uint8_t*)p.W1img + (size_t)c * TC_B_BYTES);
          ring_load((const uint8_t*)p.W1img + (size_t)c * TC_B_BYTES + T2_STAGE);
        }
        for (int s = 0; s < 8; ++s) {
          for (int h = 0; h < 2; ++h) {
            const uint8_t* w = (const uint8_t*)p.W2img + (size_t)(s * 2 + h) * TC_B_BYTES;
            ring_load(w);
            ring_load(w + T2_STAGE);
          }
        }
        if (lane == 0) TR(tl, 33);
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    // The whole warp runs the loops (uniform control flow, uniform-register descriptors); one elected lane
    // issues the tcgen05 instructions.
    {
      uint32_t it = 0, a_cnt[2] = {0, 0}, d1e_ph = 0, xr_ph = 0, d2e_ph = 1;
      long long ring_cyc = 0, afull_cyc = 0;    // debug: cycles this warp waited on the weight ring / on h1 slabs
      const uint32_t idesc1 = umma_idesc_f16(128, 128), idesc2 = umma_idesc_f16(128, 256);
      auto ring_wait = [&]() -> uint32_t {      // returns the smem address of the next stage
        const uint32_t slot = it % T2_NS, ph = (it / T2_NS) & 1;
        const long long w0 = trace ? clock64() : 0;
        mbar_wait(B(T3Bars::full + slot), ph);
        if (trace) ring_cyc += clock64() - w0;
        if (trace && lane == 0 && it >= 11 * 40 && it < 12 * 40) trace[63 * 48 + (it - 11 * 40)] = clock64();
        tc_fence_after();
        return sR_u + slot * T2_STAGE;
      };
      const uint32_t mma_gap = (uint32_t)p.mma_gap;
      auto gap = [&]() {                         // debug knob: minimum spacing between phase-2 MMA issues
        if (mma_gap) { const uint32_t t0 = (uint32_t)clock(); while ((uint32_t)clock() - t0 < mma_gap) {} }
      };
      auto ring_release = [&]() {               // call from the elected lane
        if (CL == 1) tc_commit(B(T3Bars::empty + (it % T2_NS)));
        else tc_commit_mc(B(T3Bars::empty + (it % T2_NS)), (uint16_t)((1u << CL) - 1));
      };
      const int64_t my_tiles = (n_tiles > g_first) ? (n_tiles - g_first + g_step - 1) / g_step : 0;
      for (int64_t tl64 = 0; tl64 < my_tiles; ++tl64) {
        const int tl = (int)tl64;
        mbar_wait(B(T3Bars::d2_empty), d2e_ph); d2e_ph ^= 1;      // previous tile's epilogue drained TMEM
        if (lane == 0) TR(tl, 2);
        mbar_wait(B(T3Bars::x_ready), xr_ph); xr_ph ^= 1;         // x tile (hi/lo, swizzled) is in smem
        if (lane == 0) TR(tl, 3);
        tc_fence_after();
        // ---- phase 1
        for (int c = 0; c < 4; ++c) {
          // four D1 chunks, one TMEM column range each (free since d2_empty): the layer-1 MMAs never wait for
          // the epilogue, which trails one chunk behind
          const int b = c;
          const uint32_t d = tmem + (uint32_t)(b * 128);
          const uint32_t bh = ring_wait();                        // W1 hi(c): [slab0 16 KB][slab1 16 KB]
          // Rolled k loops (descriptor += 32 B per k step): the per-MMA operand setup then overlaps the previous
          // MMA's execution.  Fully unrolled, ptxas front-loads ~40 R2UR moves before the first UTCHMMA of a
          // stage and the tensor pipe drains meanwhile (the issue queue is only a couple of MMAs deep).
          if (elect_one()) {
#pragma unroll 1
            for (int slab = 0; slab < 2; ++slab) {
              uint64_t xh = umma_desc_sw128(sX_u + slab * TC_SLAB_BYTES);
              uint64_t xl = umma_desc_sw128(sX_u + (2 + slab) * TC_SLAB_BYTES);
              uint64_t wh = umma_desc_sw128(bh + slab * TC_SLAB_BYTES);
#pragma unroll 1
              for (int ks = 0; ks < 4; ++ks) {
                tc_mma_f16(d, xh, wh, idesc1, (slab | ks) ? 1u : 0u);
                tc_mma_f16(d, xl, wh, idesc1, 1u);
                xh += 2; xl += 2; wh += 2;
              }
            }
            ring_release();
          }
          ++it;
          const uint32_t bl = ring_wait();                        // W1 lo(c)
          if (elect_one()) {
#pragma unroll 1
            for (int slab = 0; slab < 2; ++slab) {
              uint64_t xh = umma_desc_sw128(sX_u + slab * TC_SLAB_BYTES);
              uint64_t wl = umma_desc_sw128(bl + slab * TC_SLAB_BYTES);
#pragma unroll 1
              for (int ks = 0; ks < 4; ++ks) {
                tc_mma_f16(d, xh, wl, idesc1, 1u);
                xh += 2; wl += 2;
              }
            }
            ring_release();
            tc_commit(B(T3Bars::d1_full + b));
          }
          ++it;
          if (lane == 0) TR(tl, 4 + c);
        }
        if (elect_one()) {
          tc_commit(B(T3Bars::a_empty + 0));      // x tile fully consumed: both A buffers may be overwritten
          tc_commit(B(T3Bars::a_empty + 1));
        }
        // D2 overlaps the D1 buffers: wait until the epilogue drained the last two chunks
        for (int i = 0; i < 4; ++i) mbar_wait(B(T3Bars::d1_empty + i), d1e_ph);
        d1e_ph ^= 1;
        tc_fence_after();
        if (lane == 0) TR(tl, 16);
        // ---- phase 2
        for (int s = 0; s < 8; ++s) {
          const int b = s & 1;
          const long long w1 = trace ? clock64() : 0;
          mbar_wait(B(T3Bars::a_full + b), a_cnt[b] & 1); ++a_cnt[b];
          if (trace) afull_cyc += clock64() - w1;
          if (lane == 0) TR(tl, 17 + s);
          tc_fence_after();
          const uint32_t a_u = sX_u + b * T2_STAGE;               // [hi 16 KB][lo 16 KB]
          for (int h = 0; h < 2; ++h) {
            const uint32_t d = tmem + (uint32_t)(h * 256);
            const uint32_t bh = ring_wait();                      // W2 hi(s,h): [256 rows][64 k]
            if (elect_one()) {
              uint64_t ah = umma_desc_sw128(a_u), al = umma_desc_sw128(a_u + TC_SLAB_BYTES), wh = umma_desc_sw128(bh);
#pragma unroll 1
              for (int ks = 0; ks < 4; ++ks) {
                tc_mma_f16(d, ah, wh, idesc2, (s | ks) ? 1u : 0u);
                tc_mma_f16(d, al, wh, idesc2, 1u);
                ah += 2; al += 2; wh += 2;
              }
              ring_release();
            }
            ++it;
            const uint32_t bl = ring_wait();                      // W2 lo(s,h)
            if (elect_one()) {
              uint64_t ah = umma_desc_sw128(a_u), wl = umma_desc_sw128(bl);
#pragma unroll 1
              for (int ks = 0; ks < 4; ++ks) {
                tc_mma_f16(d, ah, wl, idesc2, 1u);
                ah += 2; wl += 2;
              }
              ring_release();
            }
            ++it;
          }
          if (s < 6 && elect_one()) tc_commit(B(T3Bars::a_empty + b));   // slab s consumed -> slab s+2 may load
        }
        if (elect_one()) tc_commit(B(T3Bars::d2_full));
        if (lane == 0) TR(tl, 25);
        if (trace && lane == 0 && tl < 64) { trace[tl * 48 + 46] = ring_cyc; trace[tl * 48 + 47] = afull_cyc; }
        ring_cyc = 0; afull_cyc = 0;
      }
    }
  } else if (warp == 2 + T2_EPI_WARPS) {
    // =============================== loader: h1 slabs, L2 scratch -> shared ===============================
    // cp.async (generic proxy, like the epilogue's st.global that produced the data): no cross-proxy
    // fence on global memory is needed, only the cheap shared-memory one before the MMAs read the slab.
    uint32_t a_cnt[2] = {0, 0}, h1_ph = 0;
    int tl = -1;
    for (int64_t g = g_first; g < n_tiles; g += g_step) {
      int q, t0, nt;
      if (!tile_info(g, q, t0, nt)) continue;
      ++tl;
      for (int s = 0; s < 8; ++s) {
        if ((s & 1) == 0) { mbar_wait(B(T3Bars::h1_done + (s >> 1)), h1_ph); if (lane == 0) TR(tl, 28 + (s >> 1)); }
        const int b = s & 1;
        mbar_wait(B(T3Bars::a_empty + b), a_cnt[b] & 1); ++a_cnt[b];        // x tile / slab s-2 no longer read
        if (lane == 0) TR(tl, 34 + s);
        uint8_t* dst = sX + b * T2_STAGE;
        const uint8_t* src = scratch + (size_t)s * T2_STAGE;
        // chunk-major scratch -> K-major SWIZZLE_128B tile: consecutive lanes read consecutive 16-B chunks of
        // the scratch (coalesced) and scatter them to their swizzled place in shared memory
#pragma unroll 4
        for (int idx = lane; idx < 8 * TC_M; idx += 32) {
          const int chunk = idx >> 7, r = idx & (TC_M - 1);
          const uint32_t o = sw128_chunk_off(r, chunk);
          cp_async16(dst + o, src + (size_t)idx * 16);                                   // hi plane
          cp_async16(dst + TC_SLAB_BYTES + o, src + TC_SLAB_BYTES + (size_t)idx * 16);  // lo plane
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(B(T3Bars::a_full + b));
      }
      h1_ph ^= 1;
    }
  } else {
    // =============================== gather + epilogues (warps 2..9) ===============================
    const int ew = warp - 2;                  // 0..7
    const int lane_q = warp & 3;              // TMEM lane quarter this warp may access
    const int col_half = ew >> 2;             // which half of the columns this warp's epilogues cover
    const int row = lane_q * 32 + lane;
    constexpr int ROWS_PER_WARP = TC_M / T2_EPI_WARPS;   // 16
    uint32_t d1f_ph = 0, d2f_ph = 0;
    auto row_index = [&](int q, int t0, int nt, int r) -> long long {
      const int cc = r < nt ? r : 0;          // pad with the tile's first row (scores not written)
      return p.ids ? (long long)p.ids[(int64_t)q * p.ids_stride + t0 + cc] : ((long long)q * p.rows_stride + t0 + cc);
    };
    int tl = -1;
    const bool tr_thread = (ew == 0 && lane == 0);
    for (int64_t g = g_first; g < n_tiles; g += g_step) {
      int q, t0, nt;
      if (!tile_info(g, q, t0, nt)) continue;
      ++tl;
      if (tr_thread) TR(tl, 0);
      // ---- gather + split (the previous tile's MMAs are complete: this thread waited on d2_full).
      // warp ew owns rows ew*16 .. +15: one coalesced id load, then row loads 8 at a time in flight
      // (a row = one 512-B warp request) before any of them is consumed.
      {
        const long long my_row_idx = row_index(q, t0, nt, ew * ROWS_PER_WARP + (lane & (ROWS_PER_WARP - 1)));
#pragma unroll 1
        for (int i0 = 0; i0 < ROWS_PER_WARP; i0 += 8) {
          float4 v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const long long ridx = __shfl_sync(0xffffffffu, my_row_idx, i0 + j);
            v[j] = ld_row16(p.table + ridx * MLP_D + lane * 4);
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int c = ew * ROWS_PER_WARP + i0 + j;
            uint32_t h01, l01, h23, l23;
            split2_f16(v[j].x, v[j].y, h01, l01); split2_f16(v[j].z, v[j].w, h23, l23);
            const int k = lane * 4, slab = k >> 6, chunk = (k & 63) >> 3, sub = (k & 7) * 2;
            const uint32_t off = slab * TC_SLAB_BYTES + sw128_chunk_off(c, chunk) + sub;
            *reinterpret_cast<uint2*>(sX + off) = make_uint2(h01, h23);
            *reinterpret_cast<uint2*>(sX + 2 * TC_SLAB_BYTES + off) = make_uint2(l01, l23);
          }
        }
      }
      fence_proxy_async();
      mbar_arrive(B(T3Bars::x_ready));
      if (tr_thread) TR(tl, 1);
      {  // hu[q] -> shared (2 KB): the layer-1 epilogue reads it as broadcast LDS instead of L1-missing LDGs
        const int et = ew * 32 + lane;                   // 0..255
        const float2 hv = *reinterpret_cast<const float2*>(p.hu + (int64_t)q * MLP_H + et * 2);
        *reinterpret_cast<float2*>(hu_s + et * 2) = hv;
        asm volatile("bar.sync 1, %0;" ::"n"(T2_EPI_THREADS) : "memory");
      }

      // ---- while this tile computes: pull the NEXT tile's rows towards L2 (4 x 128-B lines per row)
      {
        const int64_t g2 = g + g_step;
        int q2 = 0, t02 = 0, nt2 = 0;
        if (g2 < total) tile_info(g2, q2, t02, nt2);
        if (g2 < total && lane < ROWS_PER_WARP) {
          const float* r2 = p.table + row_index(q2, t02, nt2, ew * ROWS_PER_WARP + lane) * MLP_D;
#pragma unroll
          for (int ln = 0; ln < 4; ++ln) asm volatile("prefetch.global.L2 [%0];" ::"l"(r2 + ln * 32));
        }
      }

      // ---- epilogue 1: h1 = relu(D1 + hu) -> (hi, lo) fp16 -> L2 scratch, chunk by chunk
      const float* huq = hu_s;
      for (int c = 0; c < 4; ++c) {
        const int b = c;
        mbar_wait(B(T3Bars::d1_full + b), d1f_ph);
        tc_fence_after();
        if (tr_thread) TR(tl, 8 + c);
        uint32_t va[32], vb[32];
        tc_ld32_nowait(tmem + ((uint32_t)(lane_q * 32) << 16) + (uint32_t)(b * 128 + col_half * 64), va);
        tc_ld32_nowait(tmem + ((uint32_t)(lane_q * 32) << 16) + (uint32_t)(b * 128 + col_half * 64 + 32), vb);
        tc_ld_wait_dep(va);
        tc_ld_wait_dep(vb);
        if (tr_thread && c == 1) TR(tl, 42);
#pragma unroll
        for (int part32 = 0; part32 < 2; ++part32) {
          const int col0 = col_half * 64 + part32 * 32;
          const uint32_t (&v)[32] = part32 ? vb : va;
          const int neuron0 = c * 128 + col0;
          const int slab = neuron0 >> 6;
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {                      // 8 columns -> one 16-byte chunk per plane
            const float4 ha = *reinterpret_cast<const float4*>(huq + neuron0 + ch * 8);
            const float4 hb = *reinterpret_cast<const float4*>(huq + neuron0 + ch * 8 + 4);
            uint32_t hw[4], lw[4];
            bias_relu_split2(v[ch * 8 + 0], v[ch * 8 + 1], make_float2(ha.x, ha.y), hw[0], lw[0]);
            bias_relu_split2(v[ch * 8 + 2], v[ch * 8 + 3], make_float2(ha.z, ha.w), hw[1], lw[1]);
            bias_relu_split2(v[ch * 8 + 4], v[ch * 8 + 5], make_float2(hb.x, hb.y), hw[2], lw[2]);
            bias_relu_split2(v[ch * 8 + 6], v[ch * 8 + 7], make_float2(hb.z, hb.w), hw[3], lw[3]);
            // scratch slab layout is CHUNK-major: [plane][chunk 0..7][row 0..127][16 B]: a warp (32 rows, one
            // chunk) stores 512 contiguous bytes; the loader applies the UMMA swizzle when it copies to smem
            const int chunk = ((neuron0 & 63) >> 3) + ch;
            uint8_t* dstp = scratch + (size_t)slab * 2 * TC_SLAB_BYTES + ((size_t)chunk * TC_M + row) * 16;
            *reinterpret_cast<uint4*>(dstp) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
            *reinterpret_cast<uint4*>(dstp + TC_SLAB_BYTES) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
          }
        }
        if (tr_thread && c == 1) TR(tl, 43);
        tc_fence_before();
        mbar_arrive(B(T3Bars::d1_empty + b));     // D1[b] may be overwritten
        __threadfence_block();
        if (tr_thread && c == 1) TR(tl, 44);                     // scratch stores before the arrive; the loader reads them with cp.async
        mbar_arrive(B(T3Bars::h1_done + c));
        if (tr_thread) TR(tl, 12 + c);
      }

      d1f_ph ^= 1;

      // ---- epilogue 2: s = sum_j w3[j] * relu(D2[row][j] + b2[j]); each warp of a lane quarter takes
      // 256 of the 512 columns (j ascending inside a half), lower half + upper half
      mbar_wait(B(T3Bars::d2_full), d2f_ph); d2f_ph ^= 1;
      tc_fence_after();
      if (tr_thread) TR(tl, 26);
      float acc = 0.f;
      // The column half is a compile-time constant inside each copy, so b2 / w3 are read as constant-bank
      // operands of the FADD / FFMA themselves (a runtime half needed an LDC per pair and stalled on the MIO queue).
      auto epi2 = [&](auto colh_c) {
        constexpr int CH = decltype(colh_c)::value;
        const uint32_t tbase = tmem + ((uint32_t)(lane_q * 32) << 16) + (uint32_t)(CH * 256);
        uint32_t v0[32], v1[32];
        float2 acc2 = make_float2(0.f, 0.f);
        auto consume = [&](const uint32_t (&v)[32], auto col0_c) {
          constexpr int col0 = decltype(col0_c)::value;
#pragma unroll
          for (int j2 = 0; j2 < 16; ++j2) {
            float2 z = add2(make_float2(__uint_as_float(v[j2 * 2]), __uint_as_float(v[j2 * 2 + 1])),
                            make_float2(p.b2c[col0 + j2 * 2], p.b2c[col0 + j2 * 2 + 1]));
            z.x = fmaxf(z.x, 0.f); z.y = fmaxf(z.y, 0.f);
            acc2 = fma2(make_float2(p.w3c[col0 + j2 * 2], p.w3c[col0 + j2 * 2 + 1]), z, acc2);
          }
        };
        tc_ld32_nowait(tbase, v0);
        tc_ld_wait_dep(v0);
        // 8 blocks of 32 columns, the next one in flight while one is consumed
#define NANN_EPI2_STEP(PP)                                                                      \
        tc_ld32_nowait(tbase + (uint32_t)((2 * PP + 1) * 32), v1);                              \
        consume(v0, std::integral_constant<int, CH * 256 + (2 * PP) * 32>{});                   \
        tc_ld_wait_dep(v1);                                                                     \
        if (PP < 3) tc_ld32_nowait(tbase + (uint32_t)((2 * PP + 2) * 32), v0);                  \
        consume(v1, std::integral_constant<int, CH * 256 + (2 * PP + 1) * 32>{});               \
        if (PP < 3) tc_ld_wait_dep(v0);
        NANN_EPI2_STEP(0) NANN_EPI2_STEP(1) NANN_EPI2_STEP(2) NANN_EPI2_STEP(3)
#undef NANN_EPI2_STEP
        acc = acc2.x + acc2.y;
      };
      if (col_half == 0) epi2(std::integral_constant<int, 0>{}); else epi2(std::integral_constant<int, 1>{});
      if (tr_thread) TR(tl, 45);
      tc_fence_before();
      mbar_arrive(B(T3Bars::d2_empty));           // TMEM is free for the next tile's phase 1
      if (col_half == 1) part[row] = acc;
      asm volatile("bar.sync 1, %0;" ::"n"(T2_EPI_THREADS) : "memory");
      if (col_half == 0 && row < nt) p.out[(int64_t)q * p.out_stride + t0 + row] = acc + part[row];
      asm volatile("bar.sync 1, %0;" ::"n"(T2_EPI_THREADS) : "memory");   // part[] is reused by the next tile
      if (tr_thread) TR(tl, 27);
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (CL > 1) {   // no CTA may exit while a peer can still multicast into it or arrive on its barriers
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

}  // namespace nann
