// Tensor-core scorer v8: a 128-candidate tile is computed by a CLUSTER PAIR, each CTA owning half of the
// neurons of both layers; h1 never leaves the chip.
// Included by scorer_mlp_tc.cuh; shares its helpers, the tile list kernels and the numerics (fp16 hi/lo split).
//
// Why: TMEM holds 128 x 512 fp32 = exactly one tile's layer-2 accumulator, so a single CTA cannot overlap
// layer 1 of the next tile with layer 2 of this one, and h1 (256 KB as hi/lo fp16) does not fit in shared
// memory next to the weight ring -- v3 therefore parks h1 in an L2 scratch (global stores at ~24 B/clk/SM:
// 10.8k of its 50.9k cycles per tile) and serialises layer 1 -> epilogue -> layer 2.  Splitting the NEURONS over
// two CTAs halves every accumulator: CTA r keeps D1 = x.W1x[256r..256r+255]^T (TMEM columns 256..511) and
// D2 = h1.W2[256r..256r+255]^T (columns 0..255).  Layer 2 needs all 512 h1 values as its K dimension, so each
// CTA converts its 256 h1 columns into four 64-k slabs ((hi,lo) fp16, UMMA K-major SWIZZLE_128B), written into its own
// shared memory with st.shared and shipped to the peer's with DSMEM bulk copies (cp.async.bulk.shared::cluster
// .shared::cta, completion on the peer's mbarrier); a slab lives only until its 12 MMAs ran.
//
// The MMA warp sees one uniform stream of 10 "units" per tile, each = one 32-KB A slab x one 64-KB weight unit
// (hi stage: Ah.Wh + Al.Wh, lo stage: Ah.Wl; 12 MMAs M128 N256 K16 = 1536 tensor cycles):
//     u = 0,1      A = x slab (k 0..63 / 64..127 of the gathered rows, written by this CTA's gather)   -> D1
//     u = 2..9     A = h1 slab: own js0, own js1, peer js0, peer js1, own js2, own js3, peer js2, peer js3          -> D2
//                  (own slabs first, so a slab has two units of MMA time to cross DSMEM before the peer needs it)
// A slabs flow through a 4-slot ring (a_full: one arrival, plus 32 KB of complete_tx for a peer slab; a_empty: one
// multicast tcgen05.commit from EACH CTA's MMA warp, so a producer knows both consumers are done with a ring
// position), weights through a 3-stage cp.async.bulk ring fed from a per-rank image that is already in consumption
// order (640 KB, cyclic).  7 x 32 KB + 3 KB = all 227 KB of shared memory.
// Per tile and CTA: 120 MMAs = 15.4k tensor cycles for 128 rows per PAIR = the 30.7k-cycle/SM floor of the 3-MMA split.
//
// Epilogue warps (8) per tile i:  epi1(i): D1 -> +hu, relu, split -> slabs (local + remote)
//                                  gather(i+1) into registers, x slabs(i+1) as soon as their ring slots free up
//                                  epi2(i): partial score over this CTA's 256 layer-2 neurons -> red.add into out
// (out is zeroed by the host wrapper; two commutative adds onto 0 are deterministic).
#pragma once

namespace nann {

#ifndef NANN_T8_NA
#define NANN_T8_NA 4
#endif
constexpr int T8_NA = NANN_T8_NA;                       // A-slab ring slots (32 KB: [hi 16 KB][lo 16 KB])
constexpr int T8_NW = 7 - T8_NA;                        // weight ring stages (32 KB)
constexpr int T8_UNITS = 10;                            // units per tile
constexpr int T8_THREADS = 64 + T2_EPI_THREADS;         // 320
constexpr int T8_SMEM_BYTES = (T8_NA + T8_NW) * T2_STAGE + 3072;
constexpr int T8_IMG_BYTES_PER_RANK = T8_UNITS * 2 * T2_STAGE;   // 640 KB
static_assert(T8_SMEM_BYTES <= 232448, "shared memory budget");

struct T8Bars {
  static constexpr int w_full = 0;              // [NW]
  static constexpr int w_empty = T8_NW;         // [NW]
  static constexpr int a_full = 2 * T8_NW;      // [NA]  one arrival (+ 32 KB of complete_tx when the slab comes from the peer)
  // [3][2]  2 arrivals: multicast commit of both CTAs' MMA warps.  TWO barriers per slot, alternating by use: the
  // producers of a slot's use k are not the same threads every time, so a thread can come to its wait one phase
  // "early" (the peer has not yet consumed the use before last) and a single parity bit would let it through.
  static constexpr int a_empty = 2 * T8_NW + T8_NA;
  static constexpr int d1_full = a_empty + 2 * T8_NA;
  static constexpr int d1_empty = d1_full + 1;   // 256
  static constexpr int d2_full = d1_full + 2;
  static constexpr int d2_empty = d1_full + 3;   // 256
  static constexpr int count = d1_full + 4;
};
// barrier index / parity of "ring position q has been consumed by BOTH CTAs"
__device__ __forceinline__ int t8_consumed_idx(uint32_t q) { return T8Bars::a_empty + (int)(q % T8_NA) * 2 + (int)((q / T8_NA) & 1); }
__device__ __forceinline__ uint32_t t8_consumed_par(uint32_t q) { return ((q / T8_NA) >> 1) & 1; }
// ring positions 2,3,6,7 of a tile hold this CTA's own h1 slabs (sources of DSMEM copies to the peer's position + 2)
__device__ __forceinline__ bool t8_is_own_slab(uint32_t q) { const uint32_t u = q % T8_UNITS; return u == 2 || u == 3 || u == 6 || u == 7; }

__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, uint4 v) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// arrive on the PEER's barrier and announce `bytes` of bulk-copy traffic to it
__device__ __forceinline__ void mbar_expect_tx_remote(uint32_t cluster_addr, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.release.cluster.shared::cluster.b64 _, [%0], %1;" ::"r"(cluster_addr), "r"(bytes) : "memory");
}
// bulk copy local shared memory -> peer shared memory (async proxy / TMA engine), completion on the peer's mbarrier
__device__ __forceinline__ void bulk_s2peer(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes, uint32_t bar_cluster) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst_cluster), "r"(src_cta), "r"(bytes), "r"(bar_cluster) : "memory");
}
__device__ __forceinline__ void mbar_arrive_rel_cluster(uint32_t bar) {   // local arrive, cluster-scope release
  asm volatile("mbarrier.arrive.release.cluster.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// wait with cluster-scope acquire: the data guarded by the barrier may have been written by the peer CTA
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  unsigned long long t0 = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (!done && (spin & 63) == 63) {
      const unsigned long long now = global_ns();
      if (t0 == 0) t0 = now;
      else if (NANN_MBAR_WATCHDOG_NS && now - t0 > NANN_MBAR_WATCHDOG_NS) __trap();
    }
  }
}

// The MMA warp's wait: as few instructions as possible between two units (the tensor core's issue queue is ~2 MMAs
// = 256 cycles deep; the generic mbar_wait + bookkeeping path was ~130 SASS instructions per unit and left the pipe
// idle for ~300 cycles per unit).  try_wait with a 1 ms suspend hint; a lost arrival still traps after ~2 s.
__device__ __forceinline__ void mbar_wait_lean(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile("{\n\t.reg .pred p;\n\t"
               "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 1000000;\n\t"
               "selp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  if (done) return;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 1000000;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (NANN_MBAR_WATCHDOG_NS && spin > (uint32_t)(NANN_MBAR_WATCHDOG_NS / 1000000ull)) __trap();   // each try_wait suspends up to 1 ms
  }
}

template <bool TRACE>
__global__ void __launch_bounds__(T8_THREADS, 1)
mlp_tc8_kernel(MlpTcArgs p) {
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const uint32_t peer = rank ^ 1u;
  long long* const trace = (TRACE && blockIdx.x == 0) ? p.trace : nullptr;     // TRACE=false: every stamp compiles out
  auto TR = [&](int it_local, int ev) {
    if (TRACE && trace && it_local < 60) trace[it_local * 48 + ev] = clock64();
  };
  extern __shared__ uint8_t tc_smem_raw[];
  uint8_t* smem = tc_smem_raw;
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sA = smem;                                   // A-slab ring
  uint8_t* sW = smem + T8_NA * T2_STAGE;                // weight ring
  uint64_t* bars = (uint64_t*)(sW + T8_NW * T2_STAGE);
  static_assert(T8Bars::count < 31, "barrier block is 256 B");
  uint32_t* tmem_slot = (uint32_t*)(bars + 31);
  float* hu_s = (float*)(bars + 32);                    // [256] hoisted layer-1 prefix of this CTA's neurons
  float* part = hu_s + 256;                             // [128]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar0 = smem_u32(bars);
  auto B = [&](int idx) { return bar0 + 8u * (uint32_t)idx; };

  if (tid == 0) {
    for (int i = 0; i < T8_NW; ++i) { mbar_init(B(T8Bars::w_full + i), 1); mbar_init(B(T8Bars::w_empty + i), 1); }
    for (int i = 0; i < T8_NA; ++i) { mbar_init(B(T8Bars::a_full + i), 1); }
    for (int i = 0; i < 2 * T8_NA; ++i) { mbar_init(B(T8Bars::a_empty + i), 2); }
    mbar_init(B(T8Bars::d1_full), 1); mbar_init(B(T8Bars::d1_empty), T2_EPI_THREADS);
    mbar_init(B(T8Bars::d2_full), 1); mbar_init(B(T8Bars::d2_empty), T2_EPI_THREADS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  // the peer's barriers must be initialised before any remote arrive / multicast commit lands
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tmem_d2 = tmem, tmem_d1 = tmem + 256;
  const uint32_t sA_u = smem_u32(sA), sW_u = smem_u32(sW);

  // both CTAs of a cluster walk the SAME tiles
  const int64_t total = *p.tile_total;
  const int64_t g_first = blockIdx.x >> 1, g_step = gridDim.x >> 1;
  const int my = (int)((total > g_first) ? (total - g_first + g_step - 1) / g_step : 0);
  auto tile_info = [&](int j, int& q, int& t0, int& nt) {
    const int2 e = p.tiles[g_first + (int64_t)j * g_step];
    q = e.x;
    const int n = p.n_ptr ? p.n_ptr[q] : p.n_fixed;
    t0 = e.y * TC_M;
    nt = min(TC_M, n - t0);
  };

  if (warp == 0) {
    // =============================== producer: this rank's weight image, cyclically ===============================
    const uint8_t* img = (const uint8_t*)p.W8img + (size_t)rank * T8_IMG_BYTES_PER_RANK;
    uint32_t it = 0;
    for (int i = 0; i < my; ++i) {
#pragma unroll 1
      for (int st = 0; st < 2 * T8_UNITS; ++st) {
        const uint32_t slot = it % T8_NW, ph = (it / T8_NW) & 1;
        mbar_wait(B(T8Bars::w_empty + slot), ph ^ 1);
        if (elect_one()) {
          mbar_expect_tx(B(T8Bars::w_full + slot), T2_STAGE);
#ifndef NANN_T8_W_CHUNK
#define NANN_T8_W_CHUNK 32768
#endif
#pragma unroll
          for (uint32_t o = 0; o < (uint32_t)T2_STAGE; o += NANN_T8_W_CHUNK)
            bulk_g2s(sW_u + slot * T2_STAGE + o, img + (size_t)st * T2_STAGE + o, NANN_T8_W_CHUNK, B(T8Bars::w_full + slot));
        }
        ++it;
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer (whole warp, elected lane issues) ===============================
    // D[128 x 256] (+)= A[128 x 64] (hi,lo) * W[256 x 64]^T (hi,lo) per unit: stage "hi" 8 MMAs, stage "lo" 4 MMAs.
    // Ring positions are kept as incrementing slot/parity pairs (no divisions in the loop).
    uint32_t w_slot = 0, w_par = 0, a_slot = 0, a_par = 0, ae_sel = 0;   // ae_sel: which of the slot's two a_empty barriers
    long long ring_cyc = 0, afull_cyc = 0;
    const uint32_t idesc = umma_idesc_f16(128, 256);
    const uint64_t dA0 = umma_desc_sw128(sA_u), dW0 = umma_desc_sw128(sW_u);
    auto w_advance = [&]() { if (++w_slot == T8_NW) { w_slot = 0; w_par ^= 1; } };
    auto unit = [&](int i, int u, uint32_t d, uint32_t first) {
      long long w0 = 0;
      if (TRACE) w0 = clock64();
      // CTA-scope acquire: the slab is read by the tensor core (async proxy), never by this warp
      mbar_wait_lean(B(T8Bars::a_full + a_slot), a_par);
      if (TRACE) { const long long w1 = clock64(); afull_cyc += w1 - w0; w0 = w1; }
      const bool st11 = TRACE && trace && lane == 0 && i == 11;                 // debug: unit stamps of tile 11
      long long* const st = trace + 60 * 48 + (st11 ? u * 4 : 0);              // rows 60.. (tiles >= 60 do not stamp)
      if (st11) st[0] = clock64();
      mbar_wait_lean(B(T8Bars::w_full + w_slot), w_par);
      if (TRACE) ring_cyc += clock64() - w0;
      tc_fence_after();
      if (st11) st[1] = clock64();
      // descriptors differ only in their 14-bit address field (bytes >> 4): base + slot * (32 KB >> 4)
      const uint64_t ah0 = dA0 + (uint64_t)(a_slot * (T2_STAGE >> 4));
      if (elect_one()) {
        uint64_t ah = ah0, al = ah0 + (TC_SLAB_BYTES >> 4), wh = dW0 + (uint64_t)(w_slot * (T2_STAGE >> 4));
#pragma unroll 1
        for (int ks = 0; ks < 4; ++ks) {
          tc_mma_f16(d, ah, wh, idesc, (first && ks == 0) ? 0u : 1u);
          tc_mma_f16(d, al, wh, idesc, 1u);
          ah += 2; al += 2; wh += 2;
        }
        tc_commit(B(T8Bars::w_empty + w_slot));
      }
      w_advance();
      if (TRACE) w0 = clock64();
      mbar_wait_lean(B(T8Bars::w_full + w_slot), w_par);
      if (TRACE) ring_cyc += clock64() - w0;
      tc_fence_after();
      if (st11) st[2] = clock64();
      if (elect_one()) {
        uint64_t ah = ah0, wl = dW0 + (uint64_t)(w_slot * (T2_STAGE >> 4));
#pragma unroll 1
        for (int ks = 0; ks < 4; ++ks) {
          tc_mma_f16(d, ah, wl, idesc, 1u);
          ah += 2; wl += 2;
        }
        tc_commit(B(T8Bars::w_empty + w_slot));
        tc_commit_mc(B(T8Bars::a_empty + a_slot * 2 + ae_sel), (uint16_t)3);   // slab consumed: tell the producers of BOTH CTAs
        if (u == 1) tc_commit(B(T8Bars::d1_full));
        if (u == T8_UNITS - 1) tc_commit(B(T8Bars::d2_full));
      }
      if (st11) st[3] = clock64();
      w_advance();
      if (++a_slot == T8_NA) { a_slot = 0; a_par ^= 1; ae_sel ^= 1; }
    };
    for (int i = 0; i < my; ++i) {
      if (lane == 0) TR(i, 0);
      if (i > 0) mbar_wait_lean(B(T8Bars::d1_empty), (i - 1) & 1);
      unit(i, 0, tmem_d1, 1u);
      unit(i, 1, tmem_d1, 0u);
      if (lane == 0) TR(i, 1);
      if (i > 0) mbar_wait_lean(B(T8Bars::d2_empty), (i - 1) & 1);
      if (lane == 0) TR(i, 2);
#pragma unroll 1
      for (int u = 2; u < T8_UNITS; ++u) {
        unit(i, u, tmem_d2, u == 2 ? 1u : 0u);
        if (TRACE && lane == 0 && u < 6) TR(i, 1 + u);    // events 3..6: units 2..5 issued
      }
      if (lane == 0) TR(i, 7);
      if (TRACE && trace && lane == 0 && i < 60) { trace[i * 48 + 46] = ring_cyc; trace[i * 48 + 47] = afull_cyc; }
      ring_cyc = 0; afull_cyc = 0;
    }
  } else {
    // =============================== gather + epilogues (warps 2..9) ===============================
    const int ew = warp - 2;                  // 0..7
    const int lane_q = warp & 3;              // TMEM lane quarter this warp may access
    const int col_half = ew >> 2;             // which 128 of a 256-column accumulator this warp covers
    const int row = lane_q * 32 + lane;
    const uint32_t t_lane = (uint32_t)(lane_q * 32) << 16;
    constexpr int ROWS_PER_WARP = TC_M / T2_EPI_WARPS;   // 16
    const bool tr_thread = (ew == 0 && lane == 0);
    const uint32_t rA_u = mapa_u32(sA_u, peer);           // the peer's A ring / barrier block in the cluster window
    const uint32_t rbar0 = mapa_u32(bar0, peer);
    auto row_index = [&](int q, int t0, int nt, int r) -> long long {
      const int cc = r < nt ? r : 0;          // pad with the tile's first row (scores not written)
      return p.ids ? (long long)p.ids[(int64_t)q * p.ids_stride + t0 + cc] : ((long long)q * p.rows_stride + t0 + cc);
    };
    // ---- gather of tile j: 16 rows x 16 B per lane into registers (loads only)
    auto gather_load = [&](int j, float4 (&v)[ROWS_PER_WARP]) {
      int q, t0, nt;
      tile_info(j, q, t0, nt);
      const long long my_row_idx = row_index(q, t0, nt, ew * ROWS_PER_WARP + (lane & (ROWS_PER_WARP - 1)));
#pragma unroll
      for (int jj = 0; jj < ROWS_PER_WARP; ++jj) {
        const long long ridx = __shfl_sync(0xffffffffu, my_row_idx, jj);
        v[jj] = ld_row16(p.table + ridx * MLP_D + lane * 4);
      }
    };
    // ---- x slabs of tile j: units 10j (k 0..63, lanes 0..15) and 10j+1 (k 64..127, lanes 16..31).
    // A slot that held one of this CTA's own h1 slabs is also the SOURCE of the DSMEM copy to the peer: it may be
    // overwritten only once the peer has consumed the copy (then it has certainly been read out).
    auto x_write = [&](int j, const float4 (&v)[ROWS_PER_WARP]) {
      const uint32_t n0 = (uint32_t)j * T8_UNITS;
      const int k = (lane & 15) * 4, chunk = k >> 3, sub = (k & 7) * 2;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const uint32_t nn = n0 + half;
        if (nn >= T8_NA) {
          const uint32_t m = nn - T8_NA;                      // previous occupant of the slot
          const uint32_t q = t8_is_own_slab(m) ? m + 2 : m;   // an own slab is free once the peer consumed its copy
          mbar_wait(B(t8_consumed_idx(q)), t8_consumed_par(q));
        }
        if ((lane >> 4) == half) {
          uint8_t* dst = sA + (nn % T8_NA) * T2_STAGE;
#pragma unroll
          for (int jj = 0; jj < ROWS_PER_WARP; ++jj) {
            const int c = ew * ROWS_PER_WARP + jj;
            uint32_t h01, l01, h23, l23;
            split2_f16(v[jj].x, v[jj].y, h01, l01); split2_f16(v[jj].z, v[jj].w, h23, l23);
            const uint32_t off = sw128_chunk_off(c, chunk) + sub;
            *reinterpret_cast<uint2*>(dst + off) = make_uint2(h01, h23);
            *reinterpret_cast<uint2*>(dst + TC_SLAB_BYTES + off) = make_uint2(l01, l23);
          }
        }
        fence_proxy_async();
        asm volatile("bar.sync 1, %0;" ::"n"(T2_EPI_THREADS) : "memory");
        if (ew == 0 && lane == 0) mbar_arrive(B(T8Bars::a_full + nn % T8_NA));
      }
    };
    // ---- epilogue 1 of tile j: ALL eight warps build every slab (own neurons js*64 ..: this warp's 32 rows x the
    // 32 columns of its col_half), so a slab is ready one block-time after the previous one; local + peer copy
    auto epi1 = [&](int j) {
      mbar_wait(B(T8Bars::d1_full), j & 1);
      tc_fence_after();
      if (tr_thread) TR(j, 9);
      uint32_t va[32], vb[32];
      const uint32_t tb = tmem_d1 + t_lane + (uint32_t)(col_half * 32);
      tc_ld32_nowait(tb, va);
      auto slab = [&](uint32_t (&v)[32], uint32_t (&vnext)[32], int js) {
        // Each CTA consumes two OWN slabs before the peer's two (unit order x0 x1 | own0 own1 peer0 peer1 | own2 own3
        // peer2 peer3), so a slab has ~3k cycles of MMA work to cross DSMEM (32 KB take 2-3k cycles) before the peer
        // needs it.  The same slab therefore sits at ring position n in this CTA and n + 2 in the peer's.
        const uint32_t n = (uint32_t)j * T8_UNITS + 2 + (uint32_t)((js >> 1) * 4 + (js & 1));
        const uint32_t nr = n + 2;
        const uint32_t slot = n % T8_NA, rslot = nr % T8_NA;
        // a_empty barriers collect the multicast commits of BOTH CTAs for a ring position: position n - NA frees the
        // local slot, position nr - NA the peer's
        // (if the local slot's previous occupant was an own slab too, its copy sits at the peer's position
        // n - NA + 2 = nr - NA: the second wait covers that as well)
        if (n >= T8_NA) mbar_wait(B(t8_consumed_idx(n - T8_NA)), t8_consumed_par(n - T8_NA));
        if (nr >= T8_NA) mbar_wait(B(t8_consumed_idx(nr - T8_NA)), t8_consumed_par(nr - T8_NA));
        uint8_t* dl = sA + slot * T2_STAGE;
        tc_ld_wait_dep(v);
        if (js < 3) tc_ld32_nowait(tb + (uint32_t)((js + 1) * 64), vnext);
        else { tc_fence_before(); mbar_arrive(B(T8Bars::d1_empty)); }      // this thread's last read of D1 has landed
        const int neuron0 = js * 64 + col_half * 32;                       // own neuron index of v[0]
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          const float4 ha = *reinterpret_cast<const float4*>(hu_s + neuron0 + ch * 8);
          const float4 hb = *reinterpret_cast<const float4*>(hu_s + neuron0 + ch * 8 + 4);
          uint32_t hw[4], lw[4];
          bias_relu_split2(v[ch * 8 + 0], v[ch * 8 + 1], make_float2(ha.x, ha.y), hw[0], lw[0]);
          bias_relu_split2(v[ch * 8 + 2], v[ch * 8 + 3], make_float2(ha.z, ha.w), hw[1], lw[1]);
          bias_relu_split2(v[ch * 8 + 4], v[ch * 8 + 5], make_float2(hb.x, hb.y), hw[2], lw[2]);
          bias_relu_split2(v[ch * 8 + 6], v[ch * 8 + 7], make_float2(hb.z, hb.w), hw[3], lw[3]);
          const uint32_t off = sw128_chunk_off(row, col_half * 4 + ch);
          *reinterpret_cast<uint4*>(dl + off) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
          *reinterpret_cast<uint4*>(dl + TC_SLAB_BYTES + off) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
        }
        // the slab is complete in THIS CTA's slot once all epilogue threads are here; one thread publishes it locally
        // and ships it to the peer's slot with DSMEM bulk copies (thread-level st.shared::cluster stores, 16 B per
        // lane into 32 different rows, ran at ~9 B/clk; bulk copies reach ~17 B/clk with two in flight)
        fence_proxy_async();                                               // generic-proxy writes -> async proxy (UMMA, bulk copy)
        asm volatile("bar.sync 1, %0;" ::"n"(T2_EPI_THREADS) : "memory");
        if (ew == 0 && lane == 0) {
          mbar_arrive(B(T8Bars::a_full + slot));
          const uint32_t rb = rbar0 + 8u * (uint32_t)(T8Bars::a_full + rslot);
          const uint32_t src = sA_u + slot * T2_STAGE, dst = rA_u + rslot * T2_STAGE;
          mbar_expect_tx_remote(rb, T2_STAGE);
#ifndef NANN_T8_COPY_CHUNK
#define NANN_T8_COPY_CHUNK 16384
#endif
#pragma unroll
          for (uint32_t o = 0; o < (uint32_t)T2_STAGE; o += NANN_T8_COPY_CHUNK) bulk_s2peer(dst + o, src + o, NANN_T8_COPY_CHUNK, rb);
        }
      };
      slab(va, vb, 0);
      slab(vb, va, 1);
      slab(va, vb, 2);
      slab(vb, va, 3);
    };
    // ---- epilogue 2 of tile j: partial score over this CTA's 256 layer-2 neurons
    auto epi2 = [&](int j, auto r_c, auto ch_c) -> float {
      constexpr int R = decltype(r_c)::value, CH = decltype(ch_c)::value;
      constexpr int COL = R * 256 + CH * 128;            // first b2 / w3 index of this warp's columns
      mbar_wait(B(T8Bars::d2_full), j & 1);
      tc_fence_after();
      if (tr_thread) TR(j, 13);
      const uint32_t tb = tmem_d2 + t_lane + (uint32_t)(CH * 128);
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
      mbar_arrive(B(T8Bars::d2_empty));
      return acc2.x + acc2.y;
    };
    auto epi2_dispatch = [&](int j) -> float {
      using I0 = std::integral_constant<int, 0>; using I1 = std::integral_constant<int, 1>;
      if (rank == 0) return col_half == 0 ? epi2(j, I0{}, I0{}) : epi2(j, I0{}, I1{});
      return col_half == 0 ? epi2(j, I1{}, I0{}) : epi2(j, I1{}, I1{});
    };

    float4 xv[ROWS_PER_WARP];
    if (my > 0) { gather_load(0, xv); x_write(0, xv); }
    for (int i = 0; i < my; ++i) {
      int q, t0, nt;
      tile_info(i, q, t0, nt);
      if (tr_thread) TR(i, 8);
      {  // hu of this tile's query, own neurons (every epilogue thread is past epi1 of tile i-1: bar.sync below)
        const int et = ew * 32 + lane;
        hu_s[et] = p.hu[(int64_t)q * MLP_H + rank * 256 + et];
        asm volatile("bar.sync 1, %0;" ::"n"(T2_EPI_THREADS) : "memory");
      }
      epi1(i);
      if (tr_thread) TR(i, 10);
      if (i + 1 < my) {
        gather_load(i + 1, xv);
        if (tr_thread) TR(i, 11);
        x_write(i + 1, xv);
      }
      if (tr_thread) TR(i, 12);
      const float acc = epi2_dispatch(i);
      if (col_half == 1) part[row] = acc;
      asm volatile("bar.sync 1, %0;" ::"n"(T2_EPI_THREADS) : "memory");
      if (col_half == 0 && row < nt) atomicAdd(p.out + (int64_t)q * p.out_stride + t0 + row, acc + part[row]);
      // the bar.sync at the top of the next iteration orders these part[] reads before the next tile's writes
      if (tr_thread) TR(i, 14);
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  // no CTA may exit while the peer can still write into its shared memory or arrive on its barriers
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

// weight image of mlp_tc8_kernel: [rank 2][unit 10][plane hi|lo][256 own neurons x 64 k], K-major SWIZZLE_128B
__global__ void tc_build_w8_kernel(const float* __restrict__ W1, const float* __restrict__ W2, __half* __restrict__ img) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;   // (rank, unit, row, kk)
  if (t >= 2 * T8_UNITS * 256 * 64) return;
  const int kk = t & 63, rowi = (t >> 6) & 255, u = (t >> 14) % T8_UNITS, r = (t >> 14) / T8_UNITS;
  const int nrn = 256 * r + rowi;
  float w;
  if (u < 2) w = W1[nrn * 256 + 128 + 64 * u + kk];
  else {   // unit order of rank r: own js0, own js1, peer js0, peer js1, own js2, own js3, peer js2, peer js3
    const int idx = u - 2, js = (idx >> 2) * 2 + (idx & 1), o = (idx & 2) ? 1 - r : r;
    w = W2[nrn * 512 + 256 * o + 64 * js + kk];
  }
  __half hi, lo;
  split_f16(w, hi, lo);
  const size_t base = ((size_t)(r * T8_UNITS + u) * 2) * T2_STAGE;
  const size_t off = sw128_chunk_off(rowi, kk >> 3) + (kk & 7) * 2;
  *reinterpret_cast<__half*>((uint8_t*)img + base + off) = hi;
  *reinterpret_cast<__half*>((uint8_t*)img + base + T2_STAGE + off) = lo;
}

}  // namespace nann
