// Tensor-core scorer v7: the two layers of CONSECUTIVE tiles overlap (software pipeline across tiles).
// Included by scorer_mlp_tc.cuh; shares its helpers, the tile list kernels and the weight/scratch formats.
//
// v3 runs gather -> layer 1 -> layer 2 -> epilogue of one tile back to back; the tensor pipe idles during the
// gather, the layer-1 epilogue chain and the layer-2 epilogue (ncu: 61 % active).  Here iteration i of a CTA
// issues, in this order on the MMA warp,
//     L1c0(i)   pass0(i-1)   L1c1(i)   pass1(i-1)
// where L1c{0,1}(i) = layer 1 of tile i for neurons [0,256) / [256,512) (N=256 MMAs into TMEM columns 256..511)
// and pass{0,1}(i-1) = layer 2 of tile i-1 for output columns [0,256) / [256,512) (TMEM columns 0..255, all
// 8 k-slabs of h1 each, so h1 is streamed from the L2 scratch twice).  The 8 epilogue warps follow with
//     epi1c0(i)  epi2h0(i-1)  epi1c1(i)  gather(i+1)  epi2h1(i-1)
// so every wait of the MMA warp on an epilogue is covered by an independent block of MMAs.
//
// Shared memory (227 KB): x tile 64 KB | A slab double buffer 2 x 32 KB | weight ring 3 x 32 KB | 3 KB misc.
// h1 scratch: two 256-KB buffers per CTA (tile parity), L2-resident.
#pragma once

namespace nann {

constexpr int T7_NS = 3;                         // ring stages (32 KB each)
constexpr int T7_LOADER_WARPS = 2;
constexpr int T7_THREADS = 32 * (2 + T2_EPI_WARPS + T7_LOADER_WARPS);   // 384
constexpr int T7_SMEM_BYTES = 65536 + 65536 + T7_NS * T2_STAGE + 3072;
static_assert(T7_SMEM_BYTES <= 232448, "shared memory budget");

struct T7Bars {
  static constexpr int full = 0;        // [3]
  static constexpr int empty = 3;       // [3]
  static constexpr int a_full = 6;      // [2]
  static constexpr int a_empty = 8;     // [2]
  static constexpr int x_ready = 10;
  static constexpr int x_free = 11;
  static constexpr int d1_full = 12;    // [2] chunk c of layer 1 complete in TMEM
  static constexpr int d1_empty = 14;   // [2] ... drained
  static constexpr int h1_done = 16;    // [2] chunk c of h1 is in the scratch
  static constexpr int d2_full = 18;    // [2] pass h of layer 2 complete
  static constexpr int d2_empty = 20;   // [2] ... drained
  static constexpr int count = 22;
};

__global__ void __launch_bounds__(T7_THREADS, 1)
mlp_tc7_kernel(MlpTcArgs p) {
  long long* const trace = (blockIdx.x == 0) ? p.trace : nullptr;
  auto TR = [&](int it_local, int ev) {
    if (trace && it_local < 64) trace[it_local * 48 + ev] = clock64();
  };
  extern __shared__ uint8_t tc_smem_raw[];
  uint8_t* smem = tc_smem_raw;
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sX = smem;                                   // x tile: hi slab0, hi slab1, lo slab0, lo slab1 (16 KB each)
  uint8_t* sA = smem + 65536;                           // 2 x [hi 16 KB][lo 16 KB]
  uint8_t* sR = smem + 131072;                          // ring
  uint64_t* bars = (uint64_t*)(sR + T7_NS * T2_STAGE);
  static_assert(T7Bars::count < 31, "barrier block is 256 B");
  uint32_t* tmem_slot = (uint32_t*)(bars + 31);
  float* hu_s = (float*)(bars + 32);                    // [512]
  float* part = hu_s + MLP_H;                           // [128]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar0 = smem_u32(bars);
  auto B = [&](int idx) { return bar0 + 8u * (uint32_t)idx; };

  if (tid == 0) {
    for (int i = 0; i < T7_NS; ++i) { mbar_init(B(T7Bars::full + i), 1); mbar_init(B(T7Bars::empty + i), 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(B(T7Bars::a_full + i), T7_LOADER_WARPS); mbar_init(B(T7Bars::a_empty + i), 1);
      mbar_init(B(T7Bars::d1_full + i), 1); mbar_init(B(T7Bars::d1_empty + i), T2_EPI_THREADS);
      mbar_init(B(T7Bars::h1_done + i), T2_EPI_THREADS);
      mbar_init(B(T7Bars::d2_full + i), 1); mbar_init(B(T7Bars::d2_empty + i), T2_EPI_THREADS);
    }
    mbar_init(B(T7Bars::x_ready), T2_EPI_THREADS);
    mbar_init(B(T7Bars::x_free), 1);
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
  const uint32_t tmem_d2 = tmem, tmem_d1 = tmem + 256;
  const uint32_t sX_u = smem_u32(sX), sA_u = smem_u32(sA), sR_u = smem_u32(sR);
  uint8_t* scratch0 = p.scratch + (size_t)blockIdx.x * 2 * TC_SCRATCH_BYTES;   // [tile parity][256 KB]

  const int64_t total = *p.tile_total;
  const int64_t g_first = blockIdx.x, g_step = gridDim.x;
  const int my = (int)((total > g_first) ? (total - g_first + g_step - 1) / g_step : 0);   // tiles of this CTA
  auto tile_info = [&](int j, int& q, int& t0, int& nt) {
    const int2 e = p.tiles[g_first + (int64_t)j * g_step];
    q = e.x;
    const int n = p.n_ptr ? p.n_ptr[q] : p.n_fixed;
    t0 = e.y * TC_M;
    nt = min(TC_M, n - t0);
  };

  if (warp == 0) {
    // =============================== producer: weight stages in the MMA warp's order ===============================
    uint32_t it = 0;
    auto ring_load = [&](const void* src) {
      const uint32_t slot = it % T7_NS, ph = (it / T7_NS) & 1;
      mbar_wait(B(T7Bars::empty + slot), ph ^ 1);
      if (elect_one()) {
        mbar_expect_tx(B(T7Bars::full + slot), T2_STAGE);
        bulk_g2s(sR_u + slot * T2_STAGE, src, T2_STAGE, B(T7Bars::full + slot));
      }
      ++it;
    };
    auto unit = [&](const __half* img, int u) {          // one (k-slab, N-half) unit = hi stage + lo stage
      const uint8_t* w = (const uint8_t*)img + (size_t)u * TC_B_BYTES;
      ring_load(w);
      ring_load(w + T2_STAGE);
    };
    for (int i = 0; i <= my; ++i) {
      if (i < my) { unit(p.W1img7, 0 * 2 + 0); unit(p.W1img7, 1 * 2 + 0); }       // L1c0: slabs 0,1 of neurons [0,256)
      if (i >= 1) for (int s = 0; s < 8; ++s) unit(p.W2img, s * 2 + 0);           // pass0
      if (i < my) { unit(p.W1img7, 0 * 2 + 1); unit(p.W1img7, 1 * 2 + 1); }       // L1c1
      if (i >= 1) for (int s = 0; s < 8; ++s) unit(p.W2img, s * 2 + 1);           // pass1
    }
  } else if (warp == 1) {
    // =============================== MMA issuer (whole warp, elected lane issues) ===============================
    uint32_t it = 0, a_n = 0;
    const uint32_t idesc = umma_idesc_f16(128, 256);
    auto ring_wait = [&]() -> uint32_t {
      const uint32_t slot = it % T7_NS, ph = (it / T7_NS) & 1;
      mbar_wait(B(T7Bars::full + slot), ph);
      tc_fence_after();
      return sR_u + slot * T2_STAGE;
    };
    // D[128 x 256] (+)= A[128 x 64] (hi,lo) * W[256 x 64]^T (hi,lo): stage "hi" 8 MMAs, stage "lo" 4 MMAs
    auto unit = [&](uint32_t d, uint32_t a_hi, uint32_t a_lo, uint32_t first) {
      const uint32_t bh = ring_wait();
      if (elect_one()) {
        uint64_t ah = umma_desc_sw128(a_hi), al = umma_desc_sw128(a_lo), wh = umma_desc_sw128(bh);
#pragma unroll 1
        for (int ks = 0; ks < 4; ++ks) {
          tc_mma_f16(d, ah, wh, idesc, (first && ks == 0) ? 0u : 1u);
          tc_mma_f16(d, al, wh, idesc, 1u);
          ah += 2; al += 2; wh += 2;
        }
        tc_commit(B(T7Bars::empty + (it % T7_NS)));
      }
      ++it;
      const uint32_t bl = ring_wait();
      if (elect_one()) {
        uint64_t ah = umma_desc_sw128(a_hi), wl = umma_desc_sw128(bl);
#pragma unroll 1
        for (int ks = 0; ks < 4; ++ks) {
          tc_mma_f16(d, ah, wl, idesc, 1u);
          ah += 2; wl += 2;
        }
        tc_commit(B(T7Bars::empty + (it % T7_NS)));
      }
      ++it;
    };
    auto layer1 = [&](int i, int c) {
      if (c == 0) { mbar_wait(B(T7Bars::x_ready), i & 1); }
      // TMEM columns 256..511 are free once the previous chunk was drained (c=1: chunk 0 of this tile, c=0:
      // chunk 1 of the previous tile)
      const int prev = c ^ 1;
      const int n_prev = (c == 1) ? i + 1 : i;            // completed drains of chunk `prev` so far
      if (n_prev > 0) mbar_wait(B(T7Bars::d1_empty + prev), (n_prev - 1) & 1);
      tc_fence_after();
      for (int slab = 0; slab < 2; ++slab)
        unit(tmem_d1, sX_u + slab * TC_SLAB_BYTES, sX_u + (2 + slab) * TC_SLAB_BYTES, slab == 0);
      if (elect_one()) {
        tc_commit(B(T7Bars::d1_full + c));
        if (c == 1) tc_commit(B(T7Bars::x_free));         // the x tile may be overwritten by the next gather
      }
      if (lane == 0) TR(i, 4 + c);
    };
    auto pass = [&](int i, int h) {                        // layer 2 of tile i-1, output columns h*256..
      const int j = i - 1;
      // TMEM columns 0..255 hold the previous pass until its epilogue drained them
      const int prev = h ^ 1;
      const int n_prev = (h == 1) ? j + 1 : j;
      if (n_prev > 0) mbar_wait(B(T7Bars::d2_empty + prev), (n_prev - 1) & 1);
      tc_fence_after();
      for (int s = 0; s < 8; ++s) {
        const uint32_t b = a_n & 1;
        mbar_wait(B(T7Bars::a_full + b), (a_n >> 1) & 1);
        tc_fence_after();
        const uint32_t a_u = sA_u + b * T2_STAGE;
        unit(tmem_d2, a_u, a_u + TC_SLAB_BYTES, s == 0);
        if (elect_one()) tc_commit(B(T7Bars::a_empty + b));
        ++a_n;
      }
      if (elect_one()) tc_commit(B(T7Bars::d2_full + h));
      if (lane == 0) TR(i, 6 + h);
    };
    for (int i = 0; i <= my; ++i) {
      if (lane == 0) TR(i, 3);
      if (i < my) layer1(i, 0);
      if (i >= 1) pass(i, 0);
      if (i < my) layer1(i, 1);
      if (i >= 1) pass(i, 1);
    }
  } else if (warp >= 2 + T2_EPI_WARPS) {
    // =============================== loaders: h1 slabs, L2 scratch -> shared (twice per tile) ===============================
    const int lw = warp - (2 + T2_EPI_WARPS);            // 0..1: each copies half of a slab
    uint32_t a_n = 0;
    for (int j = 0; j < my; ++j) {
      const uint8_t* scratch = scratch0 + (size_t)(j & 1) * TC_SCRATCH_BYTES;
      for (int n = 0; n < 16; ++n) {
        const int s = n & 7;
        if (n == 0) mbar_wait(B(T7Bars::h1_done + 0), j & 1);
        if (n == 4) mbar_wait(B(T7Bars::h1_done + 1), j & 1);
        const uint32_t b = a_n & 1;
        mbar_wait(B(T7Bars::a_empty + b), ((a_n >> 1) & 1) ^ 1);
        uint8_t* dst = sA + b * T2_STAGE;
        const uint8_t* src = scratch + (size_t)s * T2_STAGE;
#pragma unroll 4
        for (int idx = lw * 512 + lane; idx < (lw + 1) * 512; idx += 32) {
          const int chunk = idx >> 7, r = idx & (TC_M - 1);
          const uint32_t o = sw128_chunk_off(r, chunk);
          cp_async16(dst + o, src + (size_t)idx * 16);                                   // hi plane
          cp_async16(dst + TC_SLAB_BYTES + o, src + TC_SLAB_BYTES + (size_t)idx * 16);  // lo plane
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(B(T7Bars::a_full + b));
        ++a_n;
      }
    }
  } else {
    // =============================== gather + epilogues (warps 2..9) ===============================
    const int ew = warp - 2;                  // 0..7
    const int lane_q = warp & 3;              // TMEM lane quarter this warp may access
    const int col_half = ew >> 2;             // which 128 of a 256-column block this warp covers
    const int row = lane_q * 32 + lane;
    const uint32_t t_lane = (uint32_t)(lane_q * 32) << 16;
    constexpr int ROWS_PER_WARP = TC_M / T2_EPI_WARPS;   // 16
    const bool tr_thread = (ew == 0 && lane == 0);
    auto row_index = [&](int q, int t0, int nt, int r) -> long long {
      const int cc = r < nt ? r : 0;          // pad with the tile's first row (scores not written)
      return p.ids ? (long long)p.ids[(int64_t)q * p.ids_stride + t0 + cc] : ((long long)q * p.rows_stride + t0 + cc);
    };
    auto gather = [&](int j) {                // rows of tile j -> (hi, lo) fp16 swizzled x tile; hu[q] -> shared
      int q, t0, nt;
      tile_info(j, q, t0, nt);
      if (j > 0) mbar_wait(B(T7Bars::x_free), (j - 1) & 1);
      const long long my_row_idx = row_index(q, t0, nt, ew * ROWS_PER_WARP + (lane & (ROWS_PER_WARP - 1)));
#pragma unroll 1
      for (int i0 = 0; i0 < ROWS_PER_WARP; i0 += 8) {
        float4 v[8];
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          const long long ridx = __shfl_sync(0xffffffffu, my_row_idx, i0 + jj);
          v[jj] = ld_row16(p.table + ridx * MLP_D + lane * 4);
        }
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          const int c = ew * ROWS_PER_WARP + i0 + jj;
          uint32_t h01, l01, h23, l23;
          split2_f16(v[jj].x, v[jj].y, h01, l01); split2_f16(v[jj].z, v[jj].w, h23, l23);
          const int k = lane * 4, slab = k >> 6, chunk = (k & 63) >> 3, sub = (k & 7) * 2;
          const uint32_t off = slab * TC_SLAB_BYTES + sw128_chunk_off(c, chunk) + sub;
          *reinterpret_cast<uint2*>(sX + off) = make_uint2(h01, h23);
          *reinterpret_cast<uint2*>(sX + 2 * TC_SLAB_BYTES + off) = make_uint2(l01, l23);
        }
      }
      fence_proxy_async();
      mbar_arrive(B(T7Bars::x_ready));
      {  // hu of tile j's query (read by epi1 of tile j, which runs after this in program order)
        asm volatile("bar.sync 1, %0;" ::"n"(T2_EPI_THREADS) : "memory");   // every thread is past epi1 of tile j-1
        const int et = ew * 32 + lane;
        const float2 hv = *reinterpret_cast<const float2*>(p.hu + (int64_t)q * MLP_H + et * 2);
        *reinterpret_cast<float2*>(hu_s + et * 2) = hv;
        asm volatile("bar.sync 1, %0;" ::"n"(T2_EPI_THREADS) : "memory");
      }
      if (j + 1 < my && lane < ROWS_PER_WARP) {   // pull the next tile's rows towards L2
        int q2, t02, nt2;
        tile_info(j + 1, q2, t02, nt2);
        const float* r2 = p.table + row_index(q2, t02, nt2, ew * ROWS_PER_WARP + lane) * MLP_D;
#pragma unroll
        for (int ln = 0; ln < 4; ++ln) asm volatile("prefetch.global.L2 [%0];" ::"l"(r2 + ln * 32));
      }
    };
    // epilogue 1, chunk c of tile j: h1 = relu(D1 + hu) -> (hi, lo) fp16 -> scratch[j & 1]
    auto epi1 = [&](int j, int c) {
      uint8_t* scratch = scratch0 + (size_t)(j & 1) * TC_SCRATCH_BYTES;
      mbar_wait(B(T7Bars::d1_full + c), j & 1);
      tc_fence_after();
      const uint32_t tb = tmem_d1 + t_lane + (uint32_t)(col_half * 128);
      uint32_t va[32], vb[32];
      tc_ld32_nowait(tb, va);
      tc_ld_wait_dep(va);
      auto block = [&](const uint32_t (&v)[32], int blk) {      // 32 columns
        const int neuron0 = c * 256 + col_half * 128 + blk * 32;
        const int slab = neuron0 >> 6;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          const float4 ha = *reinterpret_cast<const float4*>(hu_s + neuron0 + ch * 8);
          const float4 hb = *reinterpret_cast<const float4*>(hu_s + neuron0 + ch * 8 + 4);
          uint32_t hw[4], lw[4];
          bias_relu_split2(v[ch * 8 + 0], v[ch * 8 + 1], make_float2(ha.x, ha.y), hw[0], lw[0]);
          bias_relu_split2(v[ch * 8 + 2], v[ch * 8 + 3], make_float2(ha.z, ha.w), hw[1], lw[1]);
          bias_relu_split2(v[ch * 8 + 4], v[ch * 8 + 5], make_float2(hb.x, hb.y), hw[2], lw[2]);
          bias_relu_split2(v[ch * 8 + 6], v[ch * 8 + 7], make_float2(hb.z, hb.w), hw[3], lw[3]);
          const int chunk = ((neuron0 & 63) >> 3) + ch;        // chunk-major scratch (see scorer_mlp_tc3.cuh)
          uint8_t* dstp = scratch + (size_t)slab * 2 * TC_SLAB_BYTES + ((size_t)chunk * TC_M + row) * 16;
          *reinterpret_cast<uint4*>(dstp) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
          *reinterpret_cast<uint4*>(dstp + TC_SLAB_BYTES) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
        }
      };
      tc_ld32_nowait(tb + 32, vb);
      block(va, 0);
      tc_ld_wait_dep(vb);
      tc_ld32_nowait(tb + 64, va);
      block(vb, 1);
      tc_ld_wait_dep(va);
      tc_ld32_nowait(tb + 96, vb);
      block(va, 2);
      tc_ld_wait_dep(vb);
      block(vb, 3);
      tc_fence_before();
      mbar_arrive(B(T7Bars::d1_empty + c));
      __threadfence_block();                             // scratch stores before the arrive; the loaders read them with cp.async
      mbar_arrive(B(T7Bars::h1_done + c));
    };
    // epilogue 2, pass h of tile j: acc2 += sum_j w3[j] * relu(D2[row][j] + b2[j]) over this warp's 128 columns
    float2 acc2 = make_float2(0.f, 0.f);
    auto epi2 = [&](int j, auto h_c, auto ch_c) {
      constexpr int H = decltype(h_c)::value, CH = decltype(ch_c)::value;
      constexpr int COL = H * 256 + CH * 128;            // first b2 / w3 index of this warp's columns
      mbar_wait(B(T7Bars::d2_full + H), j & 1);
      tc_fence_after();
      const uint32_t tb = tmem_d2 + t_lane + (uint32_t)(CH * 128);
      uint32_t v0[32], v1[32];
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
      tc_ld32_nowait(tb, v0);
      tc_ld_wait_dep(v0);
      tc_ld32_nowait(tb + 32, v1);
      consume(v0, std::integral_constant<int, COL>{});
      tc_ld_wait_dep(v1);
      tc_ld32_nowait(tb + 64, v0);
      consume(v1, std::integral_constant<int, COL + 32>{});
      tc_ld_wait_dep(v0);
      tc_ld32_nowait(tb + 96, v1);
      consume(v0, std::integral_constant<int, COL + 64>{});
      tc_ld_wait_dep(v1);
      consume(v1, std::integral_constant<int, COL + 96>{});
      tc_fence_before();
      mbar_arrive(B(T7Bars::d2_empty + H));
    };
    auto epi2_dispatch = [&](int j, auto h_c) {
      if (col_half == 0) epi2(j, h_c, std::integral_constant<int, 0>{});
      else               epi2(j, h_c, std::integral_constant<int, 1>{});
    };
    auto finish = [&](int j) {                 // combine the two column halves of a row and write the score
      int q, t0, nt;
      tile_info(j, q, t0, nt);
      const float acc = acc2.x + acc2.y;
      acc2 = make_float2(0.f, 0.f);
      if (col_half == 1) part[row] = acc;
      asm volatile("bar.sync 1, %0;" ::"n"(T2_EPI_THREADS) : "memory");
      if (col_half == 0 && row < nt) p.out[(int64_t)q * p.out_stride + t0 + row] = acc + part[row];
      asm volatile("bar.sync 1, %0;" ::"n"(T2_EPI_THREADS) : "memory");   // part[] is reused by the next tile
    };

    if (my > 0) gather(0);
    for (int i = 0; i <= my; ++i) {
      if (tr_thread) TR(i, 0);
      if (i < my) epi1(i, 0);
      if (tr_thread) TR(i, 8);
      if (i >= 1) epi2_dispatch(i - 1, std::integral_constant<int, 0>{});
      if (tr_thread) TR(i, 9);
      if (i < my) epi1(i, 1);
      if (tr_thread) TR(i, 10);
      if (i + 1 < my) gather(i + 1);
      if (tr_thread) TR(i, 11);
      if (i >= 1) { epi2_dispatch(i - 1, std::integral_constant<int, 1>{}); finish(i - 1); }
      if (tr_thread) TR(i, 12);
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

}  // namespace nann
