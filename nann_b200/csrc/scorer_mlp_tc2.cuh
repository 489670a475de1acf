// scorer_mlp_tc2.cuh -- warp-specialised version of the tcgen05 scorer (same math, same weight
// images, same L2 scratch as scorer_mlp_tc.cuh; see that file for the numerics).
//
// Roles inside one persistent CTA (320 threads, 1 CTA / SM):
//   warp 0   producer: one thread streams 32-KB operand stages with cp.async.bulk (TMA 1-D bulk copy)
//            into a 5-deep shared-memory ring, completion on mbarriers (complete_tx)
//   warp 1   MMA issuer: one thread issues tcgen05.mma (kind::f16, M=128, N=128/256), releases ring
//            slots and publishes accumulators with tcgen05.commit -> mbarrier
//   warps 2-9 gather + epilogues: row gather/split into the x tile, TMEM -> registers epilogues
//            (two warps per TMEM lane quarter, each taking half of the columns)
// so loads, tensor-core math and epilogues of different stages overlap without any CTA-wide barrier.
//
// Shared memory: [x tile / phase-2 A double buffer 64 KB][ring 5 x 32 KB][barriers].
// Stage stream per tile (each 32 KB, contiguous in the pre-swizzled images):
//   phase 1, chunk c=0..3 :  W1hi(c) -> MMAs Xh*Bh, Xl*Bh ;  W1lo(c) -> MMAs Xh*Bl      (N=128)
//   phase 2, slab  s=0..7 :  A(s) = h1 slab from the L2 scratch (hi 16 KB + lo 16 KB), then for
//                            h=0,1: W2hi(s,h) -> Ah*Bh, Al*Bh ; W2lo(s,h) -> Ah*Bl        (N=256)
#pragma once

namespace nann {

constexpr int T2_EPI_WARPS = 8;                       // gather + epilogue warps: two per TMEM lane quarter
constexpr int T2_EPI_THREADS = T2_EPI_WARPS * 32;
constexpr int T2_THREADS = 64 + T2_EPI_THREADS;
constexpr int T2_STAGE = 32768;
constexpr int T2_NS = 5;
constexpr int T2_X_BYTES = 65536;
constexpr int T3_SMEM_BYTES = T2_X_BYTES + T2_NS * T2_STAGE + 3072 /*barriers, tmem slot, hu tile; == 227 KB*/;
constexpr int T2_SMEM_BYTES = T2_X_BYTES + T2_NS * T2_STAGE + 1024 /*align*/ + 1024 /*barriers, tmem slot, partials*/;

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

struct T2Bars {           // byte offsets inside the barrier block (8 B each)
  static constexpr int full = 0;        // [5]
  static constexpr int empty = 5;       // [5]
  static constexpr int a_full = 10;     // [2]
  static constexpr int a_empty = 12;    // [2]
  static constexpr int x_ready = 14;
  static constexpr int d1_full = 15;    // [2]
  static constexpr int d1_empty = 17;   // [2]
  static constexpr int h1_done = 19;    // [4]
  static constexpr int d2_full = 23;
  static constexpr int d2_empty = 24;
  static constexpr int count = 25;
};

__global__ void __launch_bounds__(T2_THREADS, 1)
mlp_tc2_kernel(MlpTcArgs p) {
  extern __shared__ uint8_t tc_smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)tc_smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sX = smem;                                   // x tile, later A double buffer (2 x 32 KB)
  uint8_t* sR = smem + T2_X_BYTES;                      // ring
  uint64_t* bars = (uint64_t*)(sR + T2_NS * T2_STAGE);
  uint32_t* tmem_slot = (uint32_t*)(bars + T2Bars::count);
  float* part = (float*)(bars + T2Bars::count + 1);    // [128] partial sums of the upper column half

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar0 = smem_u32(bars);
  auto B = [&](int idx) { return bar0 + 8u * (uint32_t)idx; };

  if (tid == 0) {
    for (int i = 0; i < T2_NS; ++i) { mbar_init(B(T2Bars::full + i), 1); mbar_init(B(T2Bars::empty + i), 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(B(T2Bars::a_full + i), 1); mbar_init(B(T2Bars::a_empty + i), 1);
      mbar_init(B(T2Bars::d1_full + i), 1); mbar_init(B(T2Bars::d1_empty + i), T2_EPI_THREADS);
    }
    for (int i = 0; i < 4; ++i) mbar_init(B(T2Bars::h1_done + i), T2_EPI_THREADS);
    mbar_init(B(T2Bars::x_ready), T2_EPI_THREADS);
    mbar_init(B(T2Bars::d2_full), 1);
    mbar_init(B(T2Bars::d2_empty), T2_EPI_THREADS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t sX_u = smem_u32(sX), sR_u = smem_u32(sR);
  uint8_t* scratch = p.scratch + (size_t)blockIdx.x * TC_SCRATCH_BYTES;
  const int64_t n_tiles = (int64_t)p.B * p.tiles_per_q;

  // every role walks the same tile sequence and skips the same tiles
  auto tile_info = [&](int64_t g, int& q, int& t0, int& nt) -> bool {
    q = (int)(g / p.tiles_per_q);
    const int t = (int)(g % p.tiles_per_q);
    if (p.status && p.status[q] != 0) return false;
    const int n = p.n_ptr ? p.n_ptr[q] : p.n_fixed;
    t0 = t * TC_M;
    if (t0 >= n) return false;
    nt = min(TC_M, n - t0);
    return true;
  };

  if (warp == 0) {
    // =============================== producer ===============================
    if (lane == 0) {
      uint32_t it = 0;                         // ring stage counter across tiles
      uint32_t a_cnt[2] = {0, 0};              // uses of the A buffers
      uint32_t h1_ph = 0;                      // parity of h1_done[*] (one completion per tile each)
      auto ring_load = [&](const void* src) {
        const uint32_t slot = it % T2_NS, ph = (it / T2_NS) & 1;
        mbar_wait(B(T2Bars::empty + slot), ph ^ 1);
        mbar_expect_tx(B(T2Bars::full + slot), T2_STAGE);
        bulk_g2s(sR_u + slot * T2_STAGE, src, T2_STAGE, B(T2Bars::full + slot));
        ++it;
      };
      for (int64_t g = blockIdx.x; g < n_tiles; g += gridDim.x) {
        int q, t0, nt;
        if (!tile_info(g, q, t0, nt)) continue;
        for (int c = 0; c < 4; ++c) {
          ring_load((const uint8_t*)p.W1img + (size_t)c * TC_B_BYTES);
          ring_load((const uint8_t*)p.W1img + (size_t)c * TC_B_BYTES + T2_STAGE);
        }
        for (int s = 0; s < 8; ++s) {
          if ((s & 1) == 0) mbar_wait(B(T2Bars::h1_done + (s >> 1)), h1_ph);   // epilogue of chunk s/2 wrote slabs s, s+1
          const int b = s & 1;
          mbar_wait(B(T2Bars::a_empty + b), a_cnt[b] & 1);                        // x tile / slab s-2 no longer read
          mbar_expect_tx(B(T2Bars::a_full + b), T2_STAGE);
          bulk_g2s(sX_u + b * T2_STAGE, scratch + (size_t)s * T2_STAGE, T2_STAGE, B(T2Bars::a_full + b));
          ++a_cnt[b];
          for (int h = 0; h < 2; ++h) {
            const uint8_t* w = (const uint8_t*)p.W2img + (size_t)(s * 2 + h) * TC_B_BYTES;
            ring_load(w);
            ring_load(w + T2_STAGE);
          }
        }
        h1_ph ^= 1;
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    if (lane == 0) {
      uint32_t it = 0, a_cnt[2] = {0, 0}, d1e_cnt[2] = {0, 0}, xr_ph = 0, d2e_ph = 1;
      const uint32_t idesc1 = umma_idesc_f16(128, 128), idesc2 = umma_idesc_f16(128, 256);
      auto ring_wait = [&]() -> uint32_t {      // returns the smem address of the next stage
        const uint32_t slot = it % T2_NS, ph = (it / T2_NS) & 1;
        mbar_wait(B(T2Bars::full + slot), ph);
        tc_fence_after();
        return sR_u + slot * T2_STAGE;
      };
      auto ring_release = [&]() { tc_commit(B(T2Bars::empty + (it % T2_NS))); ++it; };
      for (int64_t g = blockIdx.x; g < n_tiles; g += gridDim.x) {
        int q, t0, nt;
        if (!tile_info(g, q, t0, nt)) continue;
        mbar_wait(B(T2Bars::d2_empty), d2e_ph); d2e_ph ^= 1;      // previous tile's epilogue drained TMEM
        mbar_wait(B(T2Bars::x_ready), xr_ph); xr_ph ^= 1;         // x tile (hi/lo, swizzled) is in smem
        tc_fence_after();
        // ---- phase 1
        for (int c = 0; c < 4; ++c) {
          const int b = c & 1;
          mbar_wait(B(T2Bars::d1_empty + b), (d1e_cnt[b] & 1) ^ 1); ++d1e_cnt[b];
          tc_fence_after();
          const uint32_t d = tmem + (uint32_t)(b * 128);
          const uint32_t bh = ring_wait();                        // W1 hi(c): [slab0 16 KB][slab1 16 KB]
#pragma unroll
          for (int slab = 0; slab < 2; ++slab)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint64_t xh = umma_desc_sw128(sX_u + slab * TC_SLAB_BYTES + ks * 32);
              const uint64_t xl = umma_desc_sw128(sX_u + (2 + slab) * TC_SLAB_BYTES + ks * 32);
              const uint64_t wh = umma_desc_sw128(bh + slab * TC_SLAB_BYTES + ks * 32);
              tc_mma_f16(d, xh, wh, idesc1, (slab | ks) ? 1u : 0u);
              tc_mma_f16(d, xl, wh, idesc1, 1u);
            }
          ring_release();
          const uint32_t bl = ring_wait();                        // W1 lo(c)
#pragma unroll
          for (int slab = 0; slab < 2; ++slab)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint64_t xh = umma_desc_sw128(sX_u + slab * TC_SLAB_BYTES + ks * 32);
              const uint64_t wl = umma_desc_sw128(bl + slab * TC_SLAB_BYTES + ks * 32);
              tc_mma_f16(d, xh, wl, idesc1, 1u);
            }
          ring_release();
          tc_commit(B(T2Bars::d1_full + b));
        }
        tc_commit(B(T2Bars::a_empty + 0));      // x tile fully consumed: both A buffers may be overwritten
        tc_commit(B(T2Bars::a_empty + 1));
        // D2 overlaps the D1 buffers: wait until the epilogue drained the last two chunks
        mbar_wait(B(T2Bars::d1_empty + 0), (d1e_cnt[0] & 1) ^ 1);
        mbar_wait(B(T2Bars::d1_empty + 1), (d1e_cnt[1] & 1) ^ 1);
        tc_fence_after();
        // ---- phase 2
        for (int s = 0; s < 8; ++s) {
          const int b = s & 1;
          mbar_wait(B(T2Bars::a_full + b), a_cnt[b] & 1); ++a_cnt[b];
          tc_fence_after();
          const uint32_t a_u = sX_u + b * T2_STAGE;               // [hi 16 KB][lo 16 KB]
          for (int h = 0; h < 2; ++h) {
            const uint32_t d = tmem + (uint32_t)(h * 256);
            const uint32_t bh = ring_wait();                      // W2 hi(s,h): [256 rows][64 k]
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint64_t ah = umma_desc_sw128(a_u + ks * 32);
              const uint64_t al = umma_desc_sw128(a_u + TC_SLAB_BYTES + ks * 32);
              const uint64_t wh = umma_desc_sw128(bh + ks * 32);
              tc_mma_f16(d, ah, wh, idesc2, (s | ks) ? 1u : 0u);
              tc_mma_f16(d, al, wh, idesc2, 1u);
            }
            ring_release();
            const uint32_t bl = ring_wait();                      // W2 lo(s,h)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint64_t ah = umma_desc_sw128(a_u + ks * 32);
              const uint64_t wl = umma_desc_sw128(bl + ks * 32);
              tc_mma_f16(d, ah, wl, idesc2, 1u);
            }
            ring_release();
          }
          if (s < 6) tc_commit(B(T2Bars::a_empty + b));           // slab s consumed -> slab s+2 may load
        }
        tc_commit(B(T2Bars::d2_full));
      }
    }
  } else {
    // =============================== gather + epilogues (warps 2..9) ===============================
    const int ew = warp - 2;                  // 0..7
    const int lane_q = warp & 3;              // TMEM lane quarter this warp may access
    const int col_half = ew >> 2;             // which half of the columns this warp's epilogues cover
    const int row = lane_q * 32 + lane;
    constexpr int ROWS_PER_WARP = TC_M / T2_EPI_WARPS;   // 16
    uint32_t d1f_cnt[2] = {0, 0}, d2f_ph = 0;
    auto row_index = [&](int q, int t0, int nt, int r) -> long long {
      const int cc = r < nt ? r : 0;          // pad with the tile's first row (scores not written)
      return p.ids ? (long long)p.ids[(int64_t)q * p.ids_stride + t0 + cc] : ((long long)q * p.rows_stride + t0 + cc);
    };
    for (int64_t g = blockIdx.x; g < n_tiles; g += gridDim.x) {
      int q, t0, nt;
      if (!tile_info(g, q, t0, nt)) continue;
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
      mbar_arrive(B(T2Bars::x_ready));

      // ---- while this tile computes: pull the NEXT tile's rows towards L2 (4 x 128-B lines per row)
      {
        int64_t g2 = g + gridDim.x;
        int q2 = 0, t02 = 0, nt2 = 0;
        while (g2 < n_tiles && !tile_info(g2, q2, t02, nt2)) g2 += gridDim.x;
        if (g2 < n_tiles && lane < ROWS_PER_WARP) {
          const float* r2 = p.table + row_index(q2, t02, nt2, ew * ROWS_PER_WARP + lane) * MLP_D;
#pragma unroll
          for (int ln = 0; ln < 4; ++ln) asm volatile("prefetch.global.L2 [%0];" ::"l"(r2 + ln * 32));
        }
      }

      // ---- epilogue 1: h1 = relu(D1 + hu) -> (hi, lo) fp16 -> L2 scratch, chunk by chunk
      const float* huq = p.hu + (int64_t)q * MLP_H;
      for (int c = 0; c < 4; ++c) {
        const int b = c & 1;
        mbar_wait(B(T2Bars::d1_full + b), d1f_cnt[b] & 1); ++d1f_cnt[b];
        tc_fence_after();
#pragma unroll 1
        for (int part32 = 0; part32 < 2; ++part32) {
          const int col0 = col_half * 64 + part32 * 32;
          uint32_t v[32];
          tc_ld32(tmem + ((uint32_t)(lane_q * 32) << 16) + (uint32_t)(b * 128 + col0), v);
          const int neuron0 = c * 128 + col0;
          const int slab = neuron0 >> 6;
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {
            uint32_t hw[4], lw[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float a0 = __uint_as_float(v[ch * 8 + 2 * e]) + __ldg(huq + neuron0 + ch * 8 + 2 * e);
              float a1 = __uint_as_float(v[ch * 8 + 2 * e + 1]) + __ldg(huq + neuron0 + ch * 8 + 2 * e + 1);
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
        mbar_arrive(B(T2Bars::d1_empty + b));     // D1[b] may be overwritten
        fence_proxy_async_all();                   // scratch writes (generic proxy) -> bulk-copy reads (async proxy)
        mbar_arrive(B(T2Bars::h1_done + c));
      }

      // ---- epilogue 2: s = sum_j w3[j] * relu(D2[row][j] + b2[j]); each warp of a lane quarter takes
      // 256 of the 512 columns (j ascending inside a half), lower half + upper half
      mbar_wait(B(T2Bars::d2_full), d2f_ph); d2f_ph ^= 1;
      tc_fence_after();
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
      mbar_arrive(B(T2Bars::d2_empty));           // TMEM is free for the next tile's phase 1
      if (col_half == 1) part[row] = acc;
      asm volatile("bar.sync 1, %0;" ::"n"(T2_EPI_THREADS) : "memory");
      if (col_half == 0 && row < nt) p.out[(int64_t)q * p.out_stride + t0 + row] = acc + part[row];
      asm volatile("bar.sync 1, %0;" ::"n"(T2_EPI_THREADS) : "memory");   // part[] is reused by the next tile
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

}  // namespace nann
