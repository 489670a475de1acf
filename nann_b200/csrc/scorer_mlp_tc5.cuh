// scorer_mlp_tc5.cuh -- tensor-core scorer, fully on-chip hand-off between the two layers.
//
// What the timeline of the previous kernels showed (profiles/r01_tc_timeline_v3.log): the tile period
// was 72k cycles for 31k cycles of MMA because h1 went through an L2 scratch (uncoalesced 16-B
// stores + a GPU-scope proxy fence per chunk) and layer 2 could not start before layer 1 had ended
// (its 512 accumulator columns fill TMEM).  This kernel trades 20 % more MMA work for a pipeline
// with no such stops:
//   * layer 2 is computed in two neuron halves (passes); a pass keeps a 256-column accumulator, so
//     TMEM also holds FOUR 64-column layer-1 accumulators;
//   * inside a pass, layer 1 runs four 64-neuron chunks ahead of layer 2; the epilogue turns each
//     chunk into the (hi, lo) fp16 K-slab of layer 2 DIRECTLY in shared memory (UMMA layout, 16-B
//     st.shared, bank-conflict free) -- no global scratch, no global proxy fence;
//   * layer 1 is recomputed for the second pass (x tile stays in shared memory).
// Roles (320 threads): warp 0 producer (TMA bulk ring, optional 2-CTA multicast), warp 1 MMA issuer,
// warps 2-5 "group A" (layer-1 epilogue -> smem slabs), warps 6-9 "group B" (row gather for the next
// tile, layer-2 epilogue + score).  Every hand-off is an mbarrier; there is no CTA-wide barrier in the
// tile loop.
#pragma once

namespace nann {

constexpr int T5_THREADS = 320;
constexpr int T5_STAGE = 32768;
constexpr int T5_NS = 3;                                   // ring depth
constexpr int T5_X_BYTES = 65536;                          // x tile: hi 2 slabs + lo 2 slabs (16 KB each)
constexpr int T5_A_BYTES = 65536;                          // 2 h1 slabs, each [hi 16 KB][lo 16 KB]
constexpr int T5_SMEM_BYTES = T5_X_BYTES + T5_A_BYTES + T5_NS * T5_STAGE + 1024 /*align*/ + 512 /*barriers*/;

struct T5Bars {
  static constexpr int full = 0;       // [3]
  static constexpr int empty = 3;      // [3]
  static constexpr int a_full = 6;     // [2]   group A -> MMA      (128 arrivals)
  static constexpr int a_empty = 8;    // [2]   MMA commit -> group A
  static constexpr int d1_full = 10;   // [4]   MMA commit -> group A
  static constexpr int d1_empty = 14;  // [4]   group A -> MMA      (128)
  static constexpr int d2_full = 18;   //       MMA commit -> group B
  static constexpr int d2_empty = 19;  //       group B -> MMA      (128)
  static constexpr int x_ready = 20;   //       group B -> MMA      (128)
  static constexpr int x_free = 21;    //       MMA commit -> group B
  static constexpr int count = 22;
};


template <int CL>
__global__ void __launch_bounds__(T5_THREADS, 1)
mlp_tc5_kernel(MlpTcArgs p) {
  uint32_t cta_rank = 0;
  if (CL > 1) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));
  extern __shared__ uint8_t tc_smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)tc_smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sX = smem;
  uint8_t* sA = smem + T5_X_BYTES;
  uint8_t* sR = sA + T5_A_BYTES;
  uint64_t* bars = (uint64_t*)(sR + T5_NS * T5_STAGE);
  uint32_t* tmem_slot = (uint32_t*)(bars + T5Bars::count);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar0 = smem_u32(bars);
  auto B = [&](int idx) { return bar0 + 8u * (uint32_t)idx; };

  if (tid == 0) {
    for (int i = 0; i < T5_NS; ++i) { mbar_init(B(T5Bars::full + i), 1); mbar_init(B(T5Bars::empty + i), CL); }
    for (int i = 0; i < 2; ++i) { mbar_init(B(T5Bars::a_full + i), 128); mbar_init(B(T5Bars::a_empty + i), 1); }
    for (int i = 0; i < 4; ++i) { mbar_init(B(T5Bars::d1_full + i), 1); mbar_init(B(T5Bars::d1_empty + i), 128); }
    mbar_init(B(T5Bars::d2_full), 1); mbar_init(B(T5Bars::d2_empty), 128);
    mbar_init(B(T5Bars::x_ready), 128); mbar_init(B(T5Bars::x_free), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t sX_u = smem_u32(sX), sA_u = smem_u32(sA), sR_u = smem_u32(sR);
  constexpr uint32_t D2_COL = 0, D1_COL = 256;            // TMEM columns: D2 half [0,256), D1[b] at 256 + 64 b
  long long* const trace = (blockIdx.x == 0) ? p.trace : nullptr;   // optional CTA-0 timeline (debug)
  auto TR = [&](int tile_local, int ev) {
    if (trace && tile_local < 64) trace[tile_local * 48 + ev] = clock64();
  };

  // dense tile list; a cluster walks groups of CL tiles in lockstep (dummy tile past the end: nt = 0)
  const int64_t total = *p.tile_total;
  const int64_t n_tiles = (total + CL - 1) / CL * CL;
  const int64_t g_first = (int64_t)(blockIdx.x / CL) * CL + cta_rank, g_step = (int64_t)(gridDim.x / CL) * CL;
  auto tile_info = [&](int64_t g, int& q, int& t0, int& nt) {
    if (g >= total) { q = p.tiles[0].x; t0 = 0; nt = 0; return; }
    const int2 e = p.tiles[g];
    q = e.x;
    const int n = p.n_ptr ? p.n_ptr[q] : p.n_fixed;
    t0 = e.y * TC_M;
    nt = min(TC_M, n - t0);
  };
  auto row_index = [&](int q, int t0, int nt, int r) -> long long {
    const int cc = r < nt ? r : 0;          // pad with the tile's first row (scores not written)
    return p.ids ? (long long)p.ids[(int64_t)q * p.ids_stride + t0 + cc] : ((long long)q * p.rows_stride + t0 + cc);
  };

  if (warp == 0) {
    // ================================== producer ==================================
    if (lane == 0) {
      uint32_t it = 0;
      auto ring_load = [&](const void* src) {
        const uint32_t slot = it % T5_NS, ph = (it / T5_NS) & 1;
        mbar_wait(B(T5Bars::empty + slot), ph ^ 1);
        mbar_expect_tx(B(T5Bars::full + slot), T5_STAGE);
        if (CL == 1) {
          bulk_g2s(sR_u + slot * T5_STAGE, src, T5_STAGE, B(T5Bars::full + slot));
        } else {
          constexpr uint32_t part = T5_STAGE / CL;
          bulk_g2s_mc(sR_u + slot * T5_STAGE + cta_rank * part, (const uint8_t*)src + cta_rank * part, part,
                      B(T5Bars::full + slot), (uint16_t)((1u << CL) - 1));
        }
        ++it;
      };
      const uint8_t* W1 = (const uint8_t*)p.W1img5;
      const uint8_t* W2 = (const uint8_t*)p.W2img;
      for (int64_t g = g_first; g < n_tiles; g += g_step) {
        for (int h = 0; h < 2; ++h) {
          for (int c = 0; c < 4; ++c) ring_load(W1 + (size_t)c * T5_STAGE);
          for (int c = 0; c < 8; ++c) {
            const uint8_t* w = W2 + (size_t)(c * 2 + h) * TC_B_BYTES;
            ring_load(w);                      // W2 hi (slab c, half h)
            ring_load(w + T5_STAGE);           // W2 lo
            if (c + 4 < 8) ring_load(W1 + (size_t)(c + 4) * T5_STAGE);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================== MMA issuer ==================================
    if (lane == 0) {
      uint32_t it = 0, a_cnt[2] = {0, 0}, d1e_cnt[4] = {0, 0, 0, 0}, xr_ph = 0, d2e_cnt = 0;
      const uint32_t idesc1 = umma_idesc_f16(128, 64), idesc2 = umma_idesc_f16(128, 256);
      auto ring_wait = [&]() -> uint32_t {
        const uint32_t slot = it % T5_NS, ph = (it / T5_NS) & 1;
        mbar_wait(B(T5Bars::full + slot), ph);
        tc_fence_after();
        return sR_u + slot * T5_STAGE;
      };
      auto ring_release = [&]() {
        if (CL == 1) tc_commit(B(T5Bars::empty + (it % T5_NS)));
        else tc_commit_mc(B(T5Bars::empty + (it % T5_NS)), (uint16_t)((1u << CL) - 1));
        ++it;
      };
      // layer-1 chunk c: D1[c&3] = x . W1x[64c .. 64c+63]^T   (K = 128: 2 slabs x 4 k-steps, 3 products)
      auto L1 = [&](int c) {
        const int b = c & 3;
        const uint32_t w = ring_wait();        // [hi: slab0 8 KB, slab1 8 KB][lo: slab0, slab1]
        mbar_wait(B(T5Bars::d1_empty + b), (d1e_cnt[b] & 1) ^ 1); ++d1e_cnt[b];
        tc_fence_after();
        const uint32_t d = tmem + D1_COL + (uint32_t)(b * 64);
#pragma unroll
        for (int s = 0; s < 2; ++s)
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t xh = umma_desc_sw128(sX_u + s * TC_SLAB_BYTES + ks * 32);
            const uint64_t xl = umma_desc_sw128(sX_u + (2 + s) * TC_SLAB_BYTES + ks * 32);
            const uint64_t wh = umma_desc_sw128(w + s * 8192 + ks * 32);
            const uint64_t wl = umma_desc_sw128(w + 16384 + s * 8192 + ks * 32);
            tc_mma_f16(d, xh, wh, idesc1, (s | ks) ? 1u : 0u);
            tc_mma_f16(d, xl, wh, idesc1, 1u);
            tc_mma_f16(d, xh, wl, idesc1, 1u);
          }
        ring_release();
        tc_commit(B(T5Bars::d1_full + b));
      };
      // layer-2 slab c of pass h: D2 += h1[:, 64c..] . W2[256h.., 64c..]^T
      auto P2 = [&](int c) {
        const int b = c & 1;
        mbar_wait(B(T5Bars::a_full + b), a_cnt[b] & 1); ++a_cnt[b];
        if (c == 0) { mbar_wait(B(T5Bars::d2_empty), (d2e_cnt & 1) ^ 1); ++d2e_cnt; }   // group B drained D2
        tc_fence_after();
        const uint32_t d = tmem + D2_COL;
        const uint32_t a_u = sA_u + b * 32768;                  // [hi 16 KB][lo 16 KB]
        const uint32_t wh_u = ring_wait();
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t ah = umma_desc_sw128(a_u + ks * 32);
          const uint64_t al = umma_desc_sw128(a_u + TC_SLAB_BYTES + ks * 32);
          const uint64_t wh = umma_desc_sw128(wh_u + ks * 32);
          tc_mma_f16(d, ah, wh, idesc2, (c | ks) ? 1u : 0u);
          tc_mma_f16(d, al, wh, idesc2, 1u);
        }
        ring_release();
        const uint32_t wl_u = ring_wait();
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t ah = umma_desc_sw128(a_u + ks * 32);
          const uint64_t wl = umma_desc_sw128(wl_u + ks * 32);
          tc_mma_f16(d, ah, wl, idesc2, 1u);
        }
        ring_release();
        tc_commit(B(T5Bars::a_empty + b));
      };
      int tl = -1;
      for (int64_t g = g_first; g < n_tiles; g += g_step) {
        ++tl;
        mbar_wait(B(T5Bars::x_ready), xr_ph); xr_ph ^= 1;
        tc_fence_after();
        TR(tl, 0);
        for (int h = 0; h < 2; ++h) {
          for (int c = 0; c < 4; ++c) L1(c);
          for (int c = 0; c < 8; ++c) {
            P2(c);
            TR(tl, 26 + h * 8 + c);
            if (c + 4 < 8) {
              L1(c + 4);
              if (h == 1 && c + 4 == 7) tc_commit(B(T5Bars::x_free));   // the x tile has been read for the last time
            }
          }
          tc_commit(B(T5Bars::d2_full));
          TR(tl, 1 + h);
        }
      }
    }
  } else if (warp < 6) {
    // ======================= group A: layer-1 epilogue -> h1 slab in shared memory =======================
    const int lane_q = warp & 3;
    const int row = lane_q * 32 + lane;
    uint32_t d1f_cnt[4] = {0, 0, 0, 0}, ae_cnt[2] = {0, 0};
    int tl = -1;
    const bool trt = (warp == 2 && lane == 0);
    for (int64_t g = g_first; g < n_tiles; g += g_step) {
      int q, t0, nt;
      tile_info(g, q, t0, nt);
      ++tl;
      const float* huq = p.hu + (int64_t)q * MLP_H;
      for (int h = 0; h < 2; ++h) {
        for (int c = 0; c < 8; ++c) {
          const int b4 = c & 3, b2 = c & 1;
          mbar_wait(B(T5Bars::d1_full + b4), d1f_cnt[b4] & 1); ++d1f_cnt[b4];
          tc_fence_after();
          uint32_t v0[32], v1[32];
          tc_ld32(tmem + ((uint32_t)(lane_q * 32) << 16) + D1_COL + (uint32_t)(b4 * 64), v0);
          tc_ld32(tmem + ((uint32_t)(lane_q * 32) << 16) + D1_COL + (uint32_t)(b4 * 64 + 32), v1);
          tc_fence_before();
          mbar_arrive(B(T5Bars::d1_empty + b4));                 // D1[b4] is in registers
          mbar_wait(B(T5Bars::a_empty + b2), (ae_cnt[b2] & 1) ^ 1); ++ae_cnt[b2];   // slab buffer no longer read by MMAs
          uint8_t* slab = sA + b2 * 32768;
#pragma unroll
          for (int part32 = 0; part32 < 2; ++part32) {
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {                      // 8 columns -> one 16-B chunk per plane
              const int col = part32 * 32 + ch * 8;              // column inside the 64-neuron chunk
              const float4 ha = ldg4(huq + c * 64 + col), hb = ldg4(huq + c * 64 + col + 4);
              const float hv[8] = {ha.x, ha.y, ha.z, ha.w, hb.x, hb.y, hb.z, hb.w};
              uint32_t hw[4], lw[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const uint32_t r0 = part32 ? v1[ch * 8 + 2 * e] : v0[ch * 8 + 2 * e];
                const uint32_t r1 = part32 ? v1[ch * 8 + 2 * e + 1] : v0[ch * 8 + 2 * e + 1];
                float a0 = __uint_as_float(r0) + hv[2 * e], a1 = __uint_as_float(r1) + hv[2 * e + 1];
                a0 = a0 > 0.f ? a0 : 0.f; a1 = a1 > 0.f ? a1 : 0.f;
                split2_f16(a0, a1, hw[e], lw[e]);
              }
              const uint32_t off = sw128_chunk_off(row, col >> 3);
              *reinterpret_cast<uint4*>(slab + off) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
              *reinterpret_cast<uint4*>(slab + TC_SLAB_BYTES + off) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
            }
          }
          fence_proxy_async();                                    // st.shared (generic) -> tcgen05.mma operand reads
          mbar_arrive(B(T5Bars::a_full + b2));
          if (trt) TR(tl, 4 + h * 8 + c);
        }
      }
    }
  } else {
    // ============ group B: row gather of the next tile, layer-2 epilogue, scores ============
    const int gw = warp - 6;                    // 0..3 -> rows gw*32 .. +31
    const int lane_q = warp & 3;
    const int row = lane_q * 32 + lane;
    uint32_t d2f_cnt = 0, xf_ph = 0;
    auto gather = [&](int q, int t0, int nt) {
      const long long my_row_idx = row_index(q, t0, nt, gw * 32 + lane);
#pragma unroll 1
      for (int i0 = 0; i0 < 32; i0 += 8) {
        float4 v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const long long ridx = __shfl_sync(0xffffffffu, my_row_idx, i0 + j);
          v[j] = ld_row16(p.table + ridx * MLP_D + lane * 4);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int c = gw * 32 + i0 + j;
          uint32_t h01, l01, h23, l23;
            split2_f16(v[j].x, v[j].y, h01, l01); split2_f16(v[j].z, v[j].w, h23, l23);
          const int k = lane * 4, slab = k >> 6, chunk = (k & 63) >> 3, sub = (k & 7) * 2;
          const uint32_t off = slab * TC_SLAB_BYTES + sw128_chunk_off(c, chunk) + sub;
          *reinterpret_cast<uint2*>(sX + off) = make_uint2(h01, h23);
          *reinterpret_cast<uint2*>(sX + 2 * TC_SLAB_BYTES + off) = make_uint2(l01, l23);
        }
      }
      fence_proxy_async();
      mbar_arrive(B(T5Bars::x_ready));
    };
    auto epilogue2 = [&](int h) -> float {    // sum_j w3[j] relu(D2[row][j] + b2[j]) over this pass's 256 neurons
      mbar_wait(B(T5Bars::d2_full), d2f_cnt & 1); ++d2f_cnt;
      tc_fence_after();
      float acc = 0.f;
#pragma unroll 1
      for (int part32 = 0; part32 < 8; ++part32) {
        uint32_t v[32];
        tc_ld32(tmem + ((uint32_t)(lane_q * 32) << 16) + D2_COL + (uint32_t)(part32 * 32), v);
        const float* b2p = p.b2 + h * 256 + part32 * 32;
        const float* w3p = p.w3 + h * 256 + part32 * 32;
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const float4 bb = ldg4(b2p + j4 * 4), ww = ldg4(w3p + j4 * 4);
          float a;
          a = __uint_as_float(v[j4 * 4 + 0]) + bb.x; a = a > 0.f ? a : 0.f; acc = fmaf(ww.x, a, acc);
          a = __uint_as_float(v[j4 * 4 + 1]) + bb.y; a = a > 0.f ? a : 0.f; acc = fmaf(ww.y, a, acc);
          a = __uint_as_float(v[j4 * 4 + 2]) + bb.z; a = a > 0.f ? a : 0.f; acc = fmaf(ww.z, a, acc);
          a = __uint_as_float(v[j4 * 4 + 3]) + bb.w; a = a > 0.f ? a : 0.f; acc = fmaf(ww.w, a, acc);
        }
      }
      tc_fence_before();
      mbar_arrive(B(T5Bars::d2_empty));
      return acc;
    };
    int q, t0, nt;
    int tl = -1;
    const bool trt = (warp == 6 && lane == 0);
    if (g_first < n_tiles) { tile_info(g_first, q, t0, nt); gather(q, t0, nt); }
    for (int64_t g = g_first; g < n_tiles; g += g_step) {
      tile_info(g, q, t0, nt);
      ++tl;
      if (trt) TR(tl, 20);
      const float s0 = epilogue2(0);
      if (trt) TR(tl, 21);
      const int64_t g2 = g + g_step;
      if (g2 < n_tiles) {                       // during pass 1: stage the next tile's rows
        int q2, t02, nt2;
        tile_info(g2, q2, t02, nt2);
        mbar_wait(B(T5Bars::x_free), xf_ph); xf_ph ^= 1;
        if (trt) TR(tl, 22);
        gather(q2, t02, nt2);
        if (trt) TR(tl, 23);
      }
      const float s1 = epilogue2(1);
      if (trt) TR(tl, 25);
      if (row < nt) p.out[(int64_t)q * p.out_stride + t0 + row] = s0 + s1;
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (CL > 1) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

// W1 image for this kernel: chunk c (64 neurons) = one 32-KB stage [hi: slab0, slab1][lo: slab0, slab1],
// each slab [64 rows][64 k] fp16, K-major SWIZZLE_128B.
__global__ void tc_build_w1_v5_kernel(const float* __restrict__ W1, __half* __restrict__ img) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;   // (n, k): 512 x 128
  if (t >= 512 * 128) return;
  const int n = t >> 7, k = t & 127;
  __half hi, lo;
  split_f16(W1[n * 256 + 128 + k], hi, lo);
  const int c = n >> 6, r = n & 63, slab = k >> 6, kk = k & 63;
  const size_t base = (size_t)c * T5_STAGE + (size_t)slab * 8192;
  const size_t off = sw128_chunk_off(r, kk >> 3) + (kk & 7) * 2;
  *reinterpret_cast<__half*>((uint8_t*)img + base + off) = hi;
  *reinterpret_cast<__half*>((uint8_t*)img + base + 16384 + off) = lo;
}

}  // namespace nann
