// builder_kernels.cuh -- HNSW index construction on the GPU: what the reference does offline with
// faiss IndexHNSWFlat(d, 32) + the CSR dump of NANN_impls/nann/delivery/build_hnsw_index.py:33-67.
//
// Batch construction, per level (members = nodes whose level reaches it):
//   1. k-NN candidates by BRUTE FORCE on the tensor cores: the members' rows as one fp16 image in UMMA layout,
//      D = X.X^T tile by tile with tcgen05.mma (M128 N256 K128 per tile, fp32 accumulate in TMEM, double-buffered),
//      and a filter epilogue that never materialises D: a pair (i, j) survives only if its distance beats row i's
//      current threshold tau_i and is appended to the row's 512-entry buffer (knn_filter_kernel).  Columns are
//      visited in rounds of doubling width; after each round a warp per row keeps the best 128 and tightens tau_i
//      (knn_compact_kernel), so a round appends ~128 pairs per row, whatever the corpus size.
//   2. exact refinement: fp32 distances (sequential fmaf chains) of the 128 survivors, the closest n_cand of them,
//      HNSW's diversity heuristic on exact pair distances -> at most M forward links (knn_refine_kernel).
//   3. reverse links, de-duplication, closest `cap` per node, rows closest-first (link_* kernels).
// The arithmetic of steps 2-3 has a CPU restatement among the test infrastructure and the files are compared bit for
// bit in tests/test_builder_gpu.py; step 1 only has to deliver a superset of the n_cand exact
// nearest neighbours (fp16 products: |d2 error| ~ 1e-3, far below the rank-96 / rank-128 distance gap).
#pragma once
#include "scorer_tc_common.cuh"

namespace nann {

constexpr int KB_D = 128;           // embedding width the tensor-core pass is built for
constexpr int KB_KC = 128;          // candidates kept per row between rounds
constexpr int KB_CAP = 512;         // append buffer per row
constexpr int KB_TN = 128;          // columns per tile (UMMA N)
constexpr int KB_STRIP = 128;       // column tiles per work item (A tile re-read once per item: 1% of the B traffic)
constexpr int KB_NB = 4;            // B-tile ring stages = accumulator buffers in TMEM (4 x 128 columns)
constexpr int KB_QUEUE = 8;         // survivors a filter thread parks in shared memory before it reserves buffer slots
constexpr int KB_THREADS = 320;     // warp 0 producer, warp 1 MMA, warps 2..9 filter epilogue
constexpr int KB_A_BYTES = 32768;   // 128 rows x 128 k fp16
constexpr int KB_B_BYTES = 32768;   // one B tile: 128 rows x 128 k fp16 (half an image block)
constexpr int KB_BLK_BYTES = 65536; // image block: 256 rows x 128 k fp16
constexpr int KB_HJ_SLOTS = 16;     // column-norm ring: the producer is at most 2 * KB_NB tiles ahead of the slowest epilogue thread
constexpr int KB_SMEM_BYTES = 2 * KB_A_BYTES + KB_NB * KB_B_BYTES + KB_HJ_SLOTS * KB_TN * 4 + KB_QUEUE * 256 * 8 + 1024;
constexpr int KB_MAX_CAND = 96;     // n_cand limit of the refine kernel (cap + M at level 0 with M = 32)

// ---- exact row norms: sq[r] = sum_k x[k]^2, sequential fmaf chain in k (the oracle's definition)
__global__ void row_sq_kernel(const float* __restrict__ X, int64_t s, float* __restrict__ sq) {
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= s) return;
  const float4* x = reinterpret_cast<const float4*>(X + r * KB_D);
  float a = 0.f;
#pragma unroll 4
  for (int k = 0; k < KB_D / 4; ++k) {
    const float4 v = __ldg(x + k);
    a = fmaf(v.x, v.x, a); a = fmaf(v.y, v.y, a); a = fmaf(v.z, v.z, a); a = fmaf(v.w, v.w, a);
  }
  sq[r] = a;
}
__global__ void gather_members_kernel(const float* __restrict__ emb, const int32_t* __restrict__ nodes, int64_t s,
                                      float* __restrict__ X) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;       // one float4 each
  if (t >= s * (KB_D / 4)) return;
  const int64_t r = t / (KB_D / 4);
  reinterpret_cast<float4*>(X)[t] = __ldg(reinterpret_cast<const float4*>(emb + (int64_t)nodes[r] * KB_D) + (t % (KB_D / 4)));
}

// ---- fp16 image of X in UMMA K-major SWIZZLE_128B layout, blocked by 256 rows:
//   block b (64 KB) = [slab k 0..63: 256 rows x 128 B][slab k 64..127: 256 rows x 128 B]
// A 256-row block is one B operand tile (one 64-KB bulk copy); rows 0..127 / 128..255 of it are A operand tiles
// (the first / second 16 KB of each slab).  Pad rows are zero with hj = +inf so that they never pass the filter.
__global__ void knn_image_kernel(const float* __restrict__ X, const float* __restrict__ sq, int64_t s, int64_t s_pad,
                                 uint8_t* __restrict__ img, float* __restrict__ hj) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;       // one 16-byte chunk (8 halves) each
  if (t >= s_pad * 16) return;
  const int64_t r = t >> 4;
  const int c16 = (int)(t & 15);                 // chunk of 8 k values: k = c16*8 ..
  uint4 out = make_uint4(0, 0, 0, 0);
  if (r < s) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(X + r * KB_D + c16 * 8));
    const float4 b = __ldg(reinterpret_cast<const float4*>(X + r * KB_D + c16 * 8 + 4));
    const __half2 h0 = __floats2half2_rn(a.x, a.y), h1 = __floats2half2_rn(a.z, a.w);
    const __half2 h2 = __floats2half2_rn(b.x, b.y), h3 = __floats2half2_rn(b.z, b.w);
    out = make_uint4(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1),
                     *reinterpret_cast<const uint32_t*>(&h2), *reinterpret_cast<const uint32_t*>(&h3));
  }
  const int64_t blk = r >> 8;
  const int rr = (int)(r & 255), slab = c16 >> 3, chunk = c16 & 7;
  *reinterpret_cast<uint4*>(img + blk * KB_BLK_BYTES + slab * 32768 + sw128_chunk_off(rr, chunk)) = out;
  if (c16 == 0) hj[r] = r < s ? 0.5f * sq[r] : __int_as_float(0x7f800000);
}

struct KnnArgs {
  const uint8_t* img; const float* hj; const float* sq;
  float* tau; int* cnt; uint2* buf;      // per row of THIS launch's row range (index = row - row0)
  int64_t row0; int n_rb;                // first row (multiple of 128) and number of 128-row blocks of the range
  int64_t s;                             // members (rows >= s are padding)
  int64_t col0; int n_ct;                // column chunk: first column (multiple of 256), number of 128-column tiles
  unsigned long long* overflow;          // appends dropped because a row's buffer was full
};

// A survivor of the threshold test is parked in the thread's shared-memory queue; a full queue reserves its slots in
// the row's buffer with ONE atomicAdd.  Deliberately not inlined: the filter loop has 128 call sites, and with the body
// inlined the kernel grew to 470 KB of SASS -- rounds in which most 32-column chunks hold a survivor then ran out of the
// instruction cache and the whole build got 2.8x slower.
__device__ __noinline__ void knn_flush(uint2* my_q, int n_q, int* cnt, uint2* buf_row, unsigned long long* overflow) {
  const int slot0 = atomicAdd(cnt, n_q);
  for (int e = 0; e < n_q; ++e) {
    if (slot0 + e < KB_CAP) buf_row[slot0 + e] = my_q[e * 256];
    else atomicAdd(overflow, 1ull);
  }
}
__device__ __noinline__ int knn_park(uint2* my_q, int n_q, uint32_t d2bits, uint32_t j, int* cnt, uint2* buf_row,
                                     unsigned long long* overflow) {
  if (n_q == KB_QUEUE) { knn_flush(my_q, n_q, cnt, buf_row, overflow); n_q = 0; }
  my_q[n_q * 256] = make_uint2(d2bits, j);
  return n_q + 1;
}

__global__ void __launch_bounds__(KB_THREADS, 1)
knn_filter_kernel(KnnArgs p) {
  extern __shared__ uint8_t kb_smem[];
  uint8_t* smem = kb_smem;
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sA = smem;                         // 2 x 32 KB
  uint8_t* sB = smem + 2 * KB_A_BYTES;        // KB_NB x 32 KB
  float* hj_s = (float*)(sB + KB_NB * KB_B_BYTES);              // [KB_HJ_SLOTS][128] sq_j / 2 of the tile's columns
  uint2* queue = (uint2*)(hj_s + KB_HJ_SLOTS * KB_TN);          // [KB_QUEUE][256 filter threads] parked survivors
  uint64_t* bars = (uint64_t*)(queue + KB_QUEUE * 256);
  uint32_t* tmem_slot = (uint32_t*)(bars + 24);
  enum { A_FULL = 0, A_EMPTY = 2, B_FULL = 4, B_EMPTY = 4 + KB_NB, D_FULL = 4 + 2 * KB_NB, D_EMPTY = 4 + 3 * KB_NB };
  static_assert(4 + 4 * KB_NB <= 24, "barrier block");
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(BAR(A_FULL + i), 1); mbar_init(BAR(A_EMPTY + i), 1); }
    for (int i = 0; i < KB_NB; ++i) {
      mbar_init(BAR(B_FULL + i), 1); mbar_init(BAR(B_EMPTY + i), 1);
      mbar_init(BAR(D_FULL + i), 1); mbar_init(BAR(D_EMPTY + i), T2_EPI_THREADS);
    }
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

  const int n_strips = (p.n_ct + KB_STRIP - 1) / KB_STRIP;
  const int64_t n_items = (int64_t)n_strips * p.n_rb;          // row block varies fastest: concurrent CTAs share a strip
  auto item_of = [&](int64_t it, int& rbi, int& ct0, int& nct) {
    rbi = (int)(it % p.n_rb);
    const int strip = (int)(it / p.n_rb);
    ct0 = strip * KB_STRIP;
    nct = min(KB_STRIP, p.n_ct - ct0);
  };

  if (warp == 0) {
    // ================= producer: A tile per item, B tiles per column tile (TMA bulk copies) =================
    uint32_t ia = 0, ib = 0;
    for (int64_t it = blockIdx.x; it < n_items; it += gridDim.x, ++ia) {
      int rbi, ct0, nct;
      item_of(it, rbi, ct0, nct);
      const uint32_t ab = ia & 1;
      mbar_wait(BAR(A_EMPTY + ab), ((ia >> 1) & 1) ^ 1);
      if (elect_one()) {
        const int64_t rb = (p.row0 >> 7) + rbi;                // global 128-row block
        const uint8_t* src = p.img + (rb >> 1) * KB_BLK_BYTES + (rb & 1) * 16384;
        mbar_expect_tx(BAR(A_FULL + ab), KB_A_BYTES);
        bulk_g2s(smem_u32(sA) + ab * KB_A_BYTES, src, 16384, BAR(A_FULL + ab));
        bulk_g2s(smem_u32(sA) + ab * KB_A_BYTES + 16384, src + 32768, 16384, BAR(A_FULL + ab));
      }
      for (int t = 0; t < nct; ++t, ++ib) {
        const uint32_t st = ib % KB_NB;
        mbar_wait(BAR(B_EMPTY + st), ((ib / KB_NB) & 1) ^ 1);
        if (elect_one()) {
          const int64_t ct = (p.col0 >> 7) + ct0 + t;          // 128-column tile = half of a 256-row image block
          const uint8_t* src = p.img + (ct >> 1) * KB_BLK_BYTES + (ct & 1) * 16384;
          // the column norms ride along (a global load per compare missed the few-KB L1 on every tile).  Slot ib % 16 is
          // free: B ring (4) + accumulator ring (4) keep this warp fewer than 16 tiles ahead of the slowest filter thread.
          mbar_expect_tx(BAR(B_FULL + st), KB_B_BYTES + KB_TN * 4);
          bulk_g2s(smem_u32(sB) + st * KB_B_BYTES, src, 16384, BAR(B_FULL + st));
          bulk_g2s(smem_u32(sB) + st * KB_B_BYTES + 16384, src + 32768, 16384, BAR(B_FULL + st));
          bulk_g2s(smem_u32(hj_s) + (ib % KB_HJ_SLOTS) * KB_TN * 4, p.hj + ct * KB_TN, KB_TN * 4, BAR(B_FULL + st));
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer: D[buf] = A (128 x 128) . B^T (128 x 128), 8 x (M128 N128 K16) =================
    const uint32_t idesc = umma_idesc_f16(128, 128);
    uint32_t ia = 0, ib = 0;
    for (int64_t it = blockIdx.x; it < n_items; it += gridDim.x, ++ia) {
      int rbi, ct0, nct;
      item_of(it, rbi, ct0, nct);
      const uint32_t ab = ia & 1;
      mbar_wait(BAR(A_FULL + ab), (ia >> 1) & 1);
      const uint64_t dA = umma_desc_sw128(smem_u32(sA) + ab * KB_A_BYTES);
      for (int t = 0; t < nct; ++t, ++ib) {
        const uint32_t st = ib % KB_NB;
        mbar_wait(BAR(B_FULL + st), (ib / KB_NB) & 1);
        mbar_wait(BAR(D_EMPTY + st), ((ib / KB_NB) & 1) ^ 1);  // accumulator buffer index == B stage index
        tc_fence_after();
        if (elect_one()) {
          const uint64_t dB = umma_desc_sw128(smem_u32(sB) + st * KB_B_BYTES);
          const uint32_t d = tmem + st * 128;
#pragma unroll
          for (int slab = 0; slab < 2; ++slab)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              tc_mma_f16(d, dA + (uint64_t)((slab * 16384 + ks * 32) >> 4), dB + (uint64_t)((slab * 16384 + ks * 32) >> 4), idesc,
                         (slab | ks) ? 1u : 0u);
          tc_commit(BAR(B_EMPTY + st));
          tc_commit(BAR(D_FULL + st));
          if (t == nct - 1) tc_commit(BAR(A_EMPTY + ab));
        }
        __syncwarp();
      }
    }
  } else {
    // ================= filter epilogue: 2 threads per row (column halves), never writes D =================
    const int ew = warp - 2, lane_q = warp & 3, col_half = ew >> 2;
    const uint32_t t_lane = (uint32_t)(lane_q * 32) << 16;
    uint32_t ib = 0;
    // A survivor costs an atomicAdd on the row's counter, i.e. an L2 round trip (~1.5k cycles) that the whole warp waits
    // for; with ~2 survivors per warp and tile that was 4x the MMA time of a tile (tensor pipe 19 % busy).  Survivors are
    // parked in shared memory instead and the slots of all of them are reserved with ONE atomicAdd per thread at the end
    // of the work item (or when the thread's queue is full: the first rounds, where everything survives).
    uint2* my_q = queue + (tid - 64);
    int n_q = 0;
    auto flush = [&](int64_t lrow) {
      if (n_q == 0) return;
      knn_flush(my_q, n_q, p.cnt + lrow, p.buf + lrow * KB_CAP, p.overflow);
      n_q = 0;
    };
    for (int64_t it = blockIdx.x; it < n_items; it += gridDim.x) {
      int rbi, ct0, nct;
      item_of(it, rbi, ct0, nct);
      const int64_t lrow = (int64_t)rbi * 128 + lane_q * 32 + lane;     // row inside this launch's range
      const int64_t row = p.row0 + lrow;
      float hi = __int_as_float(0x7f800000), sqi = 0.f;                  // pad rows: nothing passes
      if (row < p.s) { sqi = p.sq[row]; hi = 0.5f * (sqi - p.tau[lrow]); }
      for (int t = 0; t < nct; ++t, ++ib) {
        const uint32_t st = ib % KB_NB;
        mbar_wait(BAR(D_FULL + st), (ib / KB_NB) & 1);
        tc_fence_after();
        const int64_t c_tile = p.col0 + (int64_t)(ct0 + t) * KB_TN + col_half * 64;
        const float* hj_t = hj_s + (ib % KB_HJ_SLOTS) * KB_TN + col_half * 64;
        const uint32_t tb = tmem + t_lane + st * 128 + (uint32_t)(col_half * 64);
        // One test per 32 columns instead of one per 4: t = max_c (dot_c - sq_c / 2) > h_i ?  (survivors are rare: ~128 per
        // row and round).  The next 32 columns are already on their way from TMEM while these are reduced, and the
        // accumulator is handed back to the MMA warp as soon as the last TMEM read has landed, not after the math.
        uint32_t va[32], vb[32];
        auto chunk = [&](uint32_t (&v)[32], uint32_t (&vnext)[32], int ch) {
          tc_ld_wait_dep(v);
          if (ch < 1) tc_ld32_nowait(tb + (ch + 1) * 32, vnext);
          else { tc_fence_before(); mbar_arrive(BAR(D_EMPTY + st)); }
          float4 h[8];
#pragma unroll
          for (int g = 0; g < 8; ++g) h[g] = *reinterpret_cast<const float4*>(hj_t + ch * 32 + g * 4);
          float m[8];
#pragma unroll
          for (int g = 0; g < 8; ++g)
            m[g] = fmaxf(fmaxf(__uint_as_float(v[g * 4 + 0]) - h[g].x, __uint_as_float(v[g * 4 + 1]) - h[g].y),
                         fmaxf(__uint_as_float(v[g * 4 + 2]) - h[g].z, __uint_as_float(v[g * 4 + 3]) - h[g].w));
          const float best = fmaxf(fmaxf(fmaxf(m[0], m[1]), fmaxf(m[2], m[3])), fmaxf(fmaxf(m[4], m[5]), fmaxf(m[6], m[7])));
          if (best > hi) {               // pass <=> tau > sq_i + sq_j - 2 dot  <=>  dot - sq_j/2 > (sq_i - tau)/2
            const int64_t c0 = c_tile + ch * 32;
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              if (m[g] > hi) {
                const float hh[4] = {h[g].x, h[g].y, h[g].z, h[g].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float dot = __uint_as_float(v[g * 4 + e]);
                  const int64_t j = c0 + g * 4 + e;
                  if ((dot - hh[e]) > hi && j != row)
                    n_q = knn_park(my_q, n_q, __float_as_uint(sqi + 2.0f * (hh[e] - dot)), (uint32_t)j, p.cnt + lrow,
                                   p.buf + lrow * KB_CAP, p.overflow);
                }
              }
            }
          }
        };
        tc_ld32_nowait(tb, va);
        chunk(va, vb, 0);
        chunk(vb, va, 1);
      }
      flush(lrow);
    }
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

// ---- between rounds: a warp per row keeps the KC entries with the smallest (d2, j) and sets tau to the KC-th d2.
// The KC-th smallest 64-bit key is found by bisection on the key bits with warp ballots (entries live in registers).
__global__ void __launch_bounds__(256)
knn_compact_kernel(float* __restrict__ tau, int* __restrict__ cnt, uint2* __restrict__ buf, int64_t n_rows, int kc) {
  const int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= n_rows) return;
  int n = cnt[r];
  if (n <= kc) return;
  n = min(n, KB_CAP);
  uint2* rb = buf + r * KB_CAP;
  constexpr int PER = KB_CAP / 32;
  unsigned long long key[PER];
#pragma unroll
  for (int t = 0; t < PER; ++t) {
    const int e = lane + 32 * t;
    const uint2 v = e < n ? rb[e] : make_uint2(0, 0);
    key[t] = e < n ? (((unsigned long long)order_key(__uint_as_float(v.x)) << 32) | v.y) : ~0ull;
  }
  // smallest K with |{key <= K}| >= kc, bit by bit from the top
  unsigned long long K = 0;
  for (int bit = 63; bit >= 0; --bit) {
    const unsigned long long trial = K | ((1ull << bit) - 1ull);     // all keys with the prefix so far and this bit 0
    int c = 0;
#pragma unroll
    for (int t = 0; t < PER; ++t) c += key[t] <= trial;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
    if (c < kc) K |= 1ull << bit;
  }
  // compact the winners (exactly kc: keys are distinct) to the front, in lane-major order
  int base = 0;
  __syncwarp();
#pragma unroll
  for (int t = 0; t < PER; ++t) {
    const bool keep = key[t] <= K;
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (keep) {
      const uint32_t ok = (uint32_t)(key[t] >> 32);
      const uint32_t fb = (ok & 0x80000000u) ? (ok & 0x7fffffffu) : ~ok;     // inverse of order_key
      rb[base + __popc(m & ((1u << lane) - 1u))] = make_uint2(fb, (uint32_t)key[t]);
    }
    base += __popc(m);
    __syncwarp();     // entry e is read into registers before anything is written: all loads happened above
  }
  if (lane == 0) {
    const uint32_t ok = (uint32_t)(K >> 32);
    tau[r] = __uint_as_float((ok & 0x80000000u) ? (ok & 0x7fffffffu) : ~ok);
    cnt[r] = kc;
  }
}

// ---- exact refinement + diversity heuristic: one CTA per row.
//   cand rows -> smem, exact d2 (fmaf chain), sort by (d2, id), first n_cand; pair tests on exact pair distances;
//   keep candidate j unless an already kept m is closer to it than the node is (pair(j,m) < d_j), up to M links.
constexpr int KR_THREADS = 128;
constexpr int KR_LD = KB_D + 4;      // smem row stride in floats: conflict-free float4 reads across rows
struct RefineArgs {
  const float* X; const float* sq;       // level-local rows
  const int* cnt; const uint2* buf;      // survivors of the tensor-core pass (index = row - row0)
  int64_t row0, n_rows, s;
  int n_cand, M;
  int32_t* fwd; float* fwd_d; int32_t* fwd_cnt;   // [s][M] level-local ids, exact d2; [s]
};
constexpr int KR_SMEM_BYTES = (KB_KC + 1) * KR_LD * 4 + KB_KC * 8 + KB_KC * 4 + KB_MAX_CAND * 4 * 4 + 64;

__device__ __forceinline__ float dot128_seq(const float* a, const float* b) {   // sequential fmaf chain, k ascending
  float acc = 0.f;
#pragma unroll 8
  for (int k = 0; k < KB_D; k += 4) {
    const float4 x = *reinterpret_cast<const float4*>(a + k), y = *reinterpret_cast<const float4*>(b + k);
    acc = fmaf(x.x, y.x, acc); acc = fmaf(x.y, y.y, acc); acc = fmaf(x.z, y.z, acc); acc = fmaf(x.w, y.w, acc);
  }
  return acc;
}

__global__ void __launch_bounds__(KR_THREADS)
knn_refine_kernel(RefineArgs p) {
  extern __shared__ uint8_t kr_smem[];
  float* rows = (float*)kr_smem;                                   // [KC + 1][KR_LD]: candidates, then the node itself
  unsigned long long* keys = (unsigned long long*)(rows + (KB_KC + 1) * KR_LD);   // [KC] (order_key(d2) << 32 | slot)
  int32_t* cid = (int32_t*)(keys + KB_KC);                         // [KC] candidate ids by slot
  uint32_t* bits = (uint32_t*)(cid + KB_KC);                       // [MAX_CAND][4] pair-test bit rows
  const int64_t lrow = blockIdx.x;
  const int64_t row = p.row0 + lrow;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (row >= p.s) return;
  const int n = min(p.cnt[lrow], KB_KC);
  const uint2* rb = p.buf + lrow * KB_CAP;
  for (int e = tid; e < KB_KC; e += KR_THREADS) cid[e] = e < n ? (int32_t)rb[e].y : -1;
  __syncthreads();
  // stage rows: a warp copies one 512-byte row per request
  for (int e = warp; e <= n; e += KR_THREADS / 32) {
    const int64_t src = e < n ? (int64_t)cid[e] : row;
    const float4 v = __ldg(reinterpret_cast<const float4*>(p.X + src * KB_D) + lane);
    *reinterpret_cast<float4*>(rows + (e < n ? e : KB_KC) * KR_LD + lane * 4) = v;
  }
  __syncthreads();
  const float* xi = rows + KB_KC * KR_LD;
  const float sqi = p.sq[row];
  for (int e = tid; e < KB_KC; e += KR_THREADS) {
    unsigned long long key = ~0ull;
    if (e < n) {
      const float d2 = (sqi + p.sq[cid[e]]) - 2.0f * dot128_seq(xi, rows + e * KR_LD);
      // ties in d2 -> smaller id first: the id is looked up through the slot, so sort on (d2, id) explicitly
      key = ((unsigned long long)order_key(d2) << 32) | (uint32_t)cid[e];
    }
    keys[e] = key;
  }
  __syncthreads();
  // bitonic sort of 128 keys (ascending); the slot of an id is recovered afterwards
  for (int size = 2; size <= KB_KC; size <<= 1)
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      if (tid < KB_KC / 2) {
        const int lo = ((tid / stride) * (stride << 1)) + (tid % stride), hi2 = lo + stride;
        const bool up = ((lo & size) == 0);
        const unsigned long long x = keys[lo], y = keys[hi2];
        if ((x > y) == up) { keys[lo] = y; keys[hi2] = x; }
      }
      __syncthreads();
    }
  const int c = min(p.n_cand, n);
  // slot_of[rank]: which staged row holds the candidate of rank `rank` (ids are distinct)
  __shared__ int slot_of[KB_MAX_CAND];
  __shared__ float dist_of[KB_MAX_CAND];
  for (int r = tid; r < c; r += KR_THREADS) {
    const int32_t id = (int32_t)(uint32_t)keys[r];
    int sl = 0;
    for (int e = 0; e < n; ++e) if (cid[e] == id) { sl = e; break; }
    slot_of[r] = sl;
    const uint32_t ok = (uint32_t)(keys[r] >> 32);
    dist_of[r] = __uint_as_float((ok & 0x80000000u) ? (ok & 0x7fffffffu) : ~ok);
  }
  for (int w = tid; w < KB_MAX_CAND * 4; w += KR_THREADS) bits[w] = 0;
  __syncthreads();
  // pair tests: bit (j, m), m < j, set when pair(j, m) < d_j
  const int n_pairs = c * (c - 1) / 2;
  for (int t = tid; t < n_pairs; t += KR_THREADS) {
    int j = (int)((1.0f + sqrtf(1.0f + 8.0f * (float)t)) * 0.5f);
    while (j * (j - 1) / 2 > t) --j;
    while ((j + 1) * j / 2 <= t) ++j;
    const int m = t - j * (j - 1) / 2;
    const int idj = (int32_t)(uint32_t)keys[j], idm = (int32_t)(uint32_t)keys[m];
    const float pd = (p.sq[idj] + p.sq[idm]) - 2.0f * dot128_seq(rows + slot_of[j] * KR_LD, rows + slot_of[m] * KR_LD);
    if (pd < dist_of[j]) atomicOr(&bits[j * 4 + (m >> 5)], 1u << (m & 31));
  }
  __syncthreads();
  if (tid == 0) {
    uint32_t kept[4] = {0, 0, 0, 0};
    int nk = 0;
    for (int j = 0; j < c && nk < p.M; ++j) {
      const bool blocked = (bits[j * 4] & kept[0]) | (bits[j * 4 + 1] & kept[1]) | (bits[j * 4 + 2] & kept[2]) | (bits[j * 4 + 3] & kept[3]);
      if (!blocked) {
        kept[j >> 5] |= 1u << (j & 31);
        p.fwd[row * p.M + nk] = (int32_t)(uint32_t)keys[j];
        p.fwd_d[row * p.M + nk] = dist_of[j];
        ++nk;
      }
    }
    p.fwd_cnt[row] = nk;
  }
}

// ---- small levels (s - 1 <= KC): every other member is a candidate, no tensor-core pass needed
__global__ void knn_all_pairs_kernel(int* __restrict__ cnt, uint2* __restrict__ buf, int64_t s) {
  const int64_t r = blockIdx.x;
  for (int j = threadIdx.x; j < s; j += blockDim.x) {
    if (j == r) continue;
    buf[r * KB_CAP + (j < r ? j : j - 1)] = make_uint2(0u, (uint32_t)j);
  }
  if (threadIdx.x == 0) cnt[r] = (int)(s - 1);
}

// ---- reverse links -----------------------------------------------------------------------------------------------
__global__ void link_count_rev_kernel(const int32_t* __restrict__ fwd, const int32_t* __restrict__ fwd_cnt, int64_t s, int M,
                                      int* __restrict__ rev_cnt) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= s * M) return;
  const int64_t i = t / M;
  if ((int)(t % M) < fwd_cnt[i]) atomicAdd(rev_cnt + fwd[t], 1);
}
__global__ void link_fill_rev_kernel(const int32_t* __restrict__ fwd, const float* __restrict__ fwd_d, const int32_t* __restrict__ fwd_cnt,
                                     int64_t s, int M, const int64_t* __restrict__ rev_off, int* __restrict__ rev_fill,
                                     uint2* __restrict__ rev) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= s * M) return;
  const int64_t i = t / M;
  if ((int)(t % M) >= fwd_cnt[i]) return;
  const int32_t j = fwd[t];
  const int pos = atomicAdd(rev_fill + j, 1);
  rev[rev_off[j] + pos] = make_uint2(__float_as_uint(fwd_d[t]), (uint32_t)i);
}
// a warp per node: links = forward U reverse, one entry per neighbour, closest `cap` by (d2, id), closest first.
// A neighbour that is both a forward and a reverse link carries the SAME d2 in both roles (the formula is symmetric),
// so after sorting on (d2, id) its two entries are adjacent.  Lists longer than 64 are folded in 64-entry chunks:
// best-so-far (<= 64) + chunk (<= 64) -> sort 128 -> drop adjacent duplicates -> keep 64.
__global__ void __launch_bounds__(128)
link_finalize_kernel(const int32_t* __restrict__ fwd, const float* __restrict__ fwd_d, const int32_t* __restrict__ fwd_cnt,
                     const int64_t* __restrict__ rev_off, const uint2* __restrict__ rev, int64_t s, int M, int cap,
                     int32_t* __restrict__ out /* [s][cap] level-local ids */, int32_t* __restrict__ out_cnt) {
  __shared__ unsigned long long sk[4][128];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t i = blockIdx.x * 4 + w;
  if (i >= s) return;
  unsigned long long* k = sk[w];
  const int nf = fwd_cnt[i];
  const int64_t r0 = rev_off[i], r1 = rev_off[i + 1];
  auto key_of = [](uint32_t d2bits, uint32_t id) { return ((unsigned long long)order_key(__uint_as_float(d2bits)) << 32) | id; };
  // best-so-far = forward links (<= M <= 64)
  for (int e = lane; e < 64; e += 32) k[e] = e < nf ? key_of(__float_as_uint(fwd_d[i * M + e]), (uint32_t)fwd[i * M + e]) : ~0ull;
  int have = nf;
  int64_t pos = r0;
  do {
    const int take = (int)((r1 - pos) < 64 ? (r1 - pos) : 64);
    for (int e = lane; e < 64; e += 32) {
      unsigned long long key = ~0ull;
      if (e < take) { const uint2 v = rev[pos + e]; key = key_of(v.x, v.y); }
      k[64 + e] = key;
    }
    pos += take;
    __syncwarp();
    for (int size = 2; size <= 128; size <<= 1)
      for (int stride = size >> 1; stride > 0; stride >>= 1) {
        for (int t = lane; t < 64; t += 32) {
          const int lo = ((t / stride) * (stride << 1)) + (t % stride), hi = lo + stride;
          const bool up = ((lo & size) == 0);
          const unsigned long long x = k[lo], y = k[hi];
          if ((x > y) == up) { k[lo] = y; k[hi] = x; }
        }
        __syncwarp();
      }
    // drop duplicates (adjacent equal keys), compact to the front, keep at most 64
    const int tot = have + take;
    int base = 0;
    unsigned long long mine[4];
    bool keepf[4];
    for (int t = 0; t < 4; ++t) {
      const int e = lane + 32 * t;
      mine[t] = k[e];
      keepf[t] = e < tot && (e == 0 || k[e - 1] != k[e]);
    }
    __syncwarp();
    for (int t = 0; t < 4; ++t) {
      const unsigned m = __ballot_sync(0xffffffffu, keepf[t]);
      const int o = base + __popc(m & ((1u << lane) - 1u));
      if (keepf[t] && o < 64) k[o] = mine[t];
      base += __popc(m);
      __syncwarp();
    }
    have = min(base, 64);
    for (int e = have + lane; e < 64; e += 32) k[e] = ~0ull;
    __syncwarp();
  } while (pos < r1);
  const int n_out = min(have, cap);
  for (int e = lane; e < n_out; e += 32) out[i * cap + e] = (int32_t)(uint32_t)k[e];
  if (lane == 0) out_cnt[i] = n_out;
}

// ---- exclusive scan of int32 counts into int64 offsets: block sums, one-block scan of the sums, add back
constexpr int SCAN_BLOCK = 1024;
__global__ void __launch_bounds__(SCAN_BLOCK)
scan_block_kernel(const int32_t* __restrict__ in, int64_t n, int64_t* __restrict__ out, int64_t* __restrict__ block_sums) {
  __shared__ int64_t ws[32];
  const int64_t i = blockIdx.x * (int64_t)SCAN_BLOCK + threadIdx.x;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t v = i < n ? in[i] : 0;
  int64_t incl = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { const int64_t t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
  if (lane == 31) ws[w] = incl;
  __syncthreads();
  if (w == 0) {
    int64_t x = ws[lane];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int64_t t = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d) x += t; }
    ws[lane] = x;
  }
  __syncthreads();
  const int64_t excl = (w > 0 ? ws[w - 1] : 0) + incl - v;
  if (i < n) out[i] = excl;
  if (threadIdx.x == SCAN_BLOCK - 1) block_sums[blockIdx.x] = excl + v;
}
__global__ void __launch_bounds__(SCAN_BLOCK)
scan_sums_kernel(int64_t* __restrict__ block_sums, int64_t n_blocks) {      // in place, exclusive; [n_blocks] = total
  __shared__ int64_t ws[32];
  __shared__ int64_t carry;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int64_t base = 0; base < n_blocks; base += SCAN_BLOCK) {
    const int64_t i = base + threadIdx.x;
    const int64_t v = i < n_blocks ? block_sums[i] : 0;
    int64_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int64_t t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
    if (lane == 31) ws[w] = incl;
    __syncthreads();
    if (w == 0) {
      int64_t x = ws[lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const int64_t t = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d) x += t; }
      ws[lane] = x;
    }
    __syncthreads();
    const int64_t excl = carry + (w > 0 ? ws[w - 1] : 0) + incl - v;
    if (i < n_blocks) block_sums[i] = excl;
    __syncthreads();
    if (threadIdx.x == SCAN_BLOCK - 1) carry = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) block_sums[n_blocks] = carry;
}
__global__ void scan_add_kernel(int64_t* __restrict__ out, int64_t n, const int64_t* __restrict__ block_sums, int64_t n_blocks) {
  const int64_t i = blockIdx.x * (int64_t)SCAN_BLOCK + threadIdx.x;
  if (i < n) out[i] += block_sums[blockIdx.x];
  if (i == n) out[n] = block_sums[n_blocks];           // grid covers n + 1 elements
}

// ---- CSR emission: per-member link rows (level-local ids) -> global ids, rows of non-members stay empty
__global__ void csr_counts_kernel(const int32_t* __restrict__ nodes, const int32_t* __restrict__ out_cnt, int64_t s,
                                  int32_t* __restrict__ counts /* [n], zeroed */) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < s) counts[nodes[i]] = out_cnt[i];
}
__global__ void csr_values_kernel(const int32_t* __restrict__ nodes, const int32_t* __restrict__ links, const int32_t* __restrict__ out_cnt,
                                  int64_t s, int cap, const int64_t* __restrict__ row_splits, int32_t* __restrict__ values) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= s * cap) return;
  const int64_t i = t / cap;
  const int e = (int)(t % cap);
  if (e < out_cnt[i]) values[row_splits[nodes[i]] + e] = nodes[links[t]];
}

}  // namespace nann
