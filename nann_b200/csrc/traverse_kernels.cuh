// traverse_kernels.cuh -- the integer side of the hot path on sm_100a:
//   K3 expand (GroupGather, non-unique)      GroupGather_kernel.cc:136-170
//   K4 visit filter (BitmapRefDifference)    bitmap_ops.cc:221-234
//   K5 stable top-k (TopKV2 order)           topk_op.cc:142-150
//   K1 row gather (GatherV2)                 build_opt_graph.py:92,144
// All of it is HBM/L2-latency-bound integer work: one warp walks one query's (or one op
// call's) id list IN ORDER, 32 ids per step, so that "first occurrence wins" and the output
// order are those of the reference's serial loops by construction.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nann {

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ uint32_t ld_cg_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ unsigned lanemask_lt() {
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

// One step of the order-preserving test-and-set over 32 consecutive list elements held one per
// lane (valid==false for padding lanes).  Returns the number kept; kept elements are written to
// out[o + rank] in lane order.  bm is this list's bitmap (word = v>>5, bit = v&31).
// bad is set when v is outside [0, 32*n_words) (caller decides what to do; nothing is written
// for such a lane).
template <typename T>
__device__ __forceinline__ int filter_step(T v, bool valid, uint32_t* bm, int64_t n_words,
                                           T* out, int64_t o, int* bad) {
  const unsigned lane = threadIdx.x & 31;
  bool inrange = valid && v >= 0 && (int64_t)(v >> 5) < n_words;
  if (valid && !inrange) *bad = 1;
  // lanes holding the same id: the lowest lane is the first occurrence inside this step
  unsigned peers = __match_any_sync(FULL, (unsigned long long)(inrange ? (long long)v : -1ll - (long long)lane));
  bool leader = inrange && ((unsigned)(__ffs(peers) - 1) == lane);
  bool keep = false;
  uint32_t bit = 0;
  uint32_t* wp = nullptr;
  if (leader) {
    wp = bm + (int64_t)(v >> 5);
    bit = 1u << (unsigned)(v & 31);
    keep = !(ld_cg_u32(wp) & bit);
  }
  unsigned keepmask = __ballot_sync(FULL, keep);
  if (keep) {
    atomicOr(wp, bit);  // distinct ids may share a word
    out[o + __popc(keepmask & lanemask_lt())] = v;
  }
  __syncwarp();  // orders this step's bitmap writes before the next step's reads
  return __popc(keepmask);
}

// ---- batched K3+K4: one warp per query ---------------------------------------------------
// frontier: f_n node ids per query at frontier + q*f_stride.  Neighbours of the frontier nodes
// are concatenated in frontier order (GroupGather) and filtered through the query's bitmap
// (BitmapRefDifference).  out_n[q] = kept, out_exp[q] = ids before filtering.
__global__ void __launch_bounds__(128)
expand_filter_kernel(const int32_t* __restrict__ nbr_vals, const int64_t* __restrict__ nbr_rs,
                     const int32_t* __restrict__ frontier, int64_t f_stride, int f_n,
                     uint32_t* __restrict__ bitmap, int64_t n_words, int32_t* __restrict__ out_ids,
                     int64_t out_stride, int32_t* __restrict__ out_n, int32_t* __restrict__ out_exp,
                     const int32_t* __restrict__ status, int B) {
  const int q = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5);
  const unsigned lane = threadIdx.x & 31;
  if (q >= B) return;
  if (status[q] != 0) {
    if (lane == 0) { out_n[q] = 0; out_exp[q] = 0; }
    return;
  }
  uint32_t* bm = bitmap + (int64_t)q * n_words;
  int32_t* out = out_ids + (int64_t)q * out_stride;
  const int32_t* fr = frontier + (int64_t)q * f_stride;
  int64_t o = 0, expanded = 0;
  int bad = 0;
  // Software pipeline (all of it latency-bound pointer chasing): the CSR row bounds of the NEXT
  // 32 frontier nodes and the neighbour ids of the NEXT 32-id step are requested before the
  // current step's bitmap round trip, so each step exposes one L2 latency instead of three.
  auto load_rows = [&](int base, long long& s, int& len) {
    const int fi = base + (int)lane;
    s = 0; len = 0;
    if (fi < f_n) {
      const int32_t node = fr[fi];
      s = nbr_rs[node];
      len = (int)(nbr_rs[node + 1] - s);
    }
  };
  long long s_nx = 0;
  int len_nx = 0;
  load_rows(0, s_nx, len_nx);
  for (int base = 0; base < f_n; base += 32) {
    const long long s = s_nx;
    const int len = len_nx;
    if (base + 32 < f_n) load_rows(base + 32, s_nx, len_nx);
    int incl = len;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int t = __shfl_up_sync(FULL, incl, d);
      if ((int)lane >= d) incl += t;
    }
    const int total = __shfl_sync(FULL, incl, 31);
    const int excl = incl - len;
    auto load_step = [&](int p0) -> int32_t {
      const int p = p0 + (int)lane;
      int j = 0;  // owner = last lane whose exclusive offset is <= p
#pragma unroll
      for (int step = 16; step > 0; step >>= 1) {
        const int cand = j + step;
        const int e = __shfl_sync(FULL, excl, cand & 31);
        if (cand < 32 && e <= p) j = cand;
      }
      const long long sj = __shfl_sync(FULL, s, j);
      const int ej = __shfl_sync(FULL, excl, j);
      return p < total ? nbr_vals[sj + (p - ej)] : -1;
    };
    int32_t v_nx = total > 0 ? load_step(0) : -1;
    for (int p0 = 0; p0 < total; p0 += 32) {
      const int32_t v = v_nx;
      if (p0 + 32 < total) v_nx = load_step(p0 + 32);
      const bool valid = p0 + (int)lane < total;
      o += filter_step<int32_t>(v, valid, bm, n_words, out, o, &bad);
    }
    expanded += total;
  }
  if (lane == 0) { out_n[q] = (int32_t)o; out_exp[q] = (int32_t)expanded; }
}

// ---- batched K3+K4, one CTA per query (the default when a round's expansion fits in shared memory) ----------
// The warp-per-query kernel above is a chain of dependent L2 round trips (~1100 cycles per 32 ids, 0.2 ms per
// level-0 round at batch 256).  The same result can be computed in parallel: an id is kept iff it was not
// visited before the round AND it is the FIRST occurrence in the concatenated list, and the output is the kept ids
// in list order.  So: stage the round's ids in shared memory, find the minimum position of every unvisited id with
// a shared-memory hash table (open addressing; an entry holds a position, its key is ids[position]; CAS to claim,
// CAS-min to lower), then compact "position == minimum position of my id" in list order with a block scan and set
// the visited bits.  Bit-identical to the serial loop (GroupGather_kernel.cc:136-170 + bitmap_ops.cc:221-234).
constexpr int EFC_THREADS = 512;
constexpr int EFC_MAX_FRONTIER = 1024;
constexpr uint32_t EFC_SKIP = 0x80000000u;        // flag on a staged id: visited before the round / out of range

__device__ __forceinline__ int efc_block_scan(int v, int* s_warp, int* tot) {   // exclusive scan over EFC_THREADS threads
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(FULL, incl, d); if (lane >= d) incl += t; }
  if (lane == 31) s_warp[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    int w = lane < EFC_THREADS / 32 ? s_warp[lane] : 0;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(FULL, w, d); if (lane >= d) w += t; }
    s_warp[lane] = w;
  }
  __syncthreads();
  const int before = wid > 0 ? s_warp[wid - 1] : 0;
  *tot = s_warp[EFC_THREADS / 32 - 1];
  __syncthreads();
  return before + incl - v;
}

__global__ void __launch_bounds__(EFC_THREADS)
expand_filter_cta_kernel(const int32_t* __restrict__ nbr_vals, const int64_t* __restrict__ nbr_rs,
                         const int32_t* __restrict__ frontier, int64_t f_stride, int f_n,
                         uint32_t* __restrict__ bitmap, int64_t n_words, int32_t* __restrict__ out_ids,
                         int64_t out_stride, int32_t* __restrict__ out_n, int32_t* __restrict__ out_exp,
                         const int32_t* __restrict__ status, int cap, int tsize) {
  extern __shared__ uint32_t efc_smem[];
  uint32_t* ids = efc_smem;                                  // [cap] staged ids (| EFC_SKIP)
  unsigned short* tab = (unsigned short*)(ids + cap);        // [tsize] min position per distinct unvisited id, 0xFFFF = empty
  __shared__ int s_off[EFC_MAX_FRONTIER + 1];
  __shared__ long long s_row[EFC_MAX_FRONTIER];
  __shared__ int s_warp[32];
  const int q = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (status[q] != 0) {
    if (tid == 0) { out_n[q] = 0; out_exp[q] = 0; }
    return;
  }
  uint32_t* bm = bitmap + (int64_t)q * n_words;
  int32_t* out = out_ids + (int64_t)q * out_stride;
  const int32_t* fr = frontier + (int64_t)q * f_stride;
  // ---- row bounds of the frontier nodes -> exclusive offsets of their neighbour lists in the concatenation
  int total = 0;
  for (int base = 0; base < f_n; base += EFC_THREADS) {
    const int i = base + tid;
    int len = 0;
    if (i < f_n) {
      const int32_t node = fr[i];
      const long long s = nbr_rs[node];
      len = (int)(nbr_rs[node + 1] - s);
      s_row[i] = s;
    }
    int chunk_tot;
    const int ex = efc_block_scan(len, s_warp, &chunk_tot);
    if (i < f_n) s_off[i] = total + ex;
    total += chunk_tot;
  }
  if (tid == 0) s_off[f_n] = total;
  for (int i = tid; i < tsize; i += EFC_THREADS) tab[i] = 0xFFFFu;
  __syncthreads();
  if (total > cap) __trap();                                 // the host sized cap from f_n * max_deg
  // ---- stage the ids (one warp per frontier node) and test them against the bitmap as it is before the round
  for (int i = wid; i < f_n; i += EFC_THREADS / 32) {
    const int o = s_off[i], len = s_off[i + 1] - o;
    const long long s = s_row[i];
    for (int j = lane; j < len; j += 32) {
      const int32_t v = nbr_vals[s + j];
      uint32_t e = (uint32_t)v;
      if (v < 0 || (int64_t)(v >> 5) >= n_words) e = EFC_SKIP;
      else if (ld_cg_u32(bm + (v >> 5)) & (1u << (v & 31))) e |= EFC_SKIP;
      ids[o + j] = e;
    }
  }
  __syncthreads();
  const uint32_t tmask = (uint32_t)tsize - 1u;
  auto slot_of = [&](uint32_t v) { return (v * 2654435761u) >> 7 & tmask; };
  // ---- minimum position per id
  for (int p = tid; p < total; p += EFC_THREADS) {
    const uint32_t v = ids[p];
    if (v & EFC_SKIP) continue;
    uint32_t h = slot_of(v);
    for (;;) {
      unsigned short cur = *(volatile unsigned short*)(tab + h);
      if (cur == 0xFFFFu) {
        cur = atomicCAS(tab + h, (unsigned short)0xFFFFu, (unsigned short)p);
        if (cur == 0xFFFFu) break;                           // claimed the slot for this id
      }
      if (ids[cur] == v) {                                   // the slot's key is this id (its position may still drop)
        while ((int)cur > p) {
          const unsigned short old = atomicCAS(tab + h, cur, (unsigned short)p);
          if (old == cur) break;
          cur = old;
        }
        break;
      }
      h = (h + 1) & tmask;
    }
  }
  __syncthreads();
  // ---- first occurrences in list order -> output; visited bits
  int o = 0;
  for (int base = 0; base < total; base += EFC_THREADS) {
    const int p = base + tid;
    bool keep = false;
    uint32_t v = 0;
    if (p < total) {
      v = ids[p];
      if (!(v & EFC_SKIP)) {
        uint32_t h = slot_of(v);
        for (;;) {
          const unsigned short cur = tab[h];
          if (ids[cur] == v) { keep = (int)cur == p; break; }
          h = (h + 1) & tmask;
        }
      }
    }
    int chunk_tot;
    const int rank = efc_block_scan(keep ? 1 : 0, s_warp, &chunk_tot);
    if (keep) {
      out[o + rank] = (int32_t)v;
      atomicOr(bm + (v >> 5), 1u << (v & 31));
    }
    o += chunk_tot;
  }
  if (tid == 0) { out_n[q] = o; out_exp[q] = total; }
}

// ---- fills as kernels.  cudaMemsetAsync / device-to-device cudaMemcpyAsync are "implicit synchronisation" points
// between streams on some driver paths (CUDA programming guide, "Implicit Synchronization"): a memset of one
// searcher could not start while another stream's kernel -- e.g. the shard group's flag-wait kernel -- was still
// running.  Kernels have no such coupling, and they fold into a CUDA graph as ordinary nodes.
__global__ void fill_words_kernel(uint32_t* __restrict__ p, int64_t n_words, uint32_t v) {
  const int64_t n4 = n_words >> 2;            // p is 16-byte aligned (cudaMalloc + row sizes in words: callers check)
  uint4* p4 = reinterpret_cast<uint4*>(p);
  const uint4 v4 = make_uint4(v, v, v, v);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) p4[i] = v4;
  for (int64_t i = (n4 << 2) + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_words; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}
// start of a search call: status/round counters = 0, result slabs = 0xFF (ids -1, scores NaN pattern), and the
// users copy when the caller's batch already lives in device memory
struct SearchInitArgs {
  int32_t* status; int B;
  int32_t* round_n; int32_t* round_exp; int64_t n_round;     // 5 * max_batch
  uint32_t* out_item_w; uint32_t* out_sc_w; uint32_t* out_nodes_w; int64_t n_out;   // B*k (out_item: 2 words each)
  const float* users_src; float* users_dst; int64_t n_users; // nullptr src: copied by the host (H2D)
};
__global__ void search_init_kernel(SearchInitArgs a) {
  const int64_t t0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, step = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = t0; i < a.B; i += step) a.status[i] = 0;
  for (int64_t i = t0; i < a.n_round; i += step) { a.round_n[i] = 0; a.round_exp[i] = 0; }
  for (int64_t i = t0; i < a.n_out; i += step) {
    a.out_item_w[2 * i] = 0xFFFFFFFFu; a.out_item_w[2 * i + 1] = 0xFFFFFFFFu;
    a.out_sc_w[i] = 0xFFFFFFFFu; a.out_nodes_w[i] = 0xFFFFFFFFu;
  }
  if (a.users_src) for (int64_t i = t0; i < a.n_users; i += step) a.users_dst[i] = a.users_src[i];
}

// ---- mark: set the bits of list[q][0..n) (set_difference on a list that is already unique:
// build_opt_graph.py:119-120,132-133).  One thread per (q, i).
__global__ void mark_kernel(const int32_t* __restrict__ list, int64_t stride, int n,
                            uint32_t* __restrict__ bitmap, int64_t n_words,
                            const int32_t* __restrict__ status, int B) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= (int64_t)B * n) return;
  const int q = (int)(t / n), i = (int)(t % n);
  if (status[q] != 0) return;
  const int32_t v = list[(int64_t)q * stride + i];
  atomicOr(bitmap + (int64_t)q * n_words + (v >> 5), 1u << (v & 31));
}

// ---- op-level BitmapRefDifference: ONE warp walks all groups in order ---------------------
template <typename T>
__global__ void bitmap_diff_op_kernel(const T* __restrict__ vals, const int64_t* __restrict__ rs,
                                      int64_t n_groups, uint32_t* __restrict__ bm, int64_t n_words,
                                      T* __restrict__ out, int64_t* __restrict__ out_rs,
                                      int* __restrict__ bad_flag) {
  const unsigned lane = threadIdx.x & 31;
  int64_t o = 0;
  int bad = 0;
  if (lane == 0) out_rs[0] = 0;
  for (int64_t g = 0; g < n_groups; ++g) {
    const int64_t b = rs[g], e = rs[g + 1];
    for (int64_t p0 = b; p0 < e; p0 += 32) {
      const int64_t p = p0 + lane;
      const bool valid = p < e;
      T v = valid ? vals[p] : (T)-1;
      o += filter_step<T>(v, valid, bm, n_words, out, o, &bad);
    }
    if (lane == 0) out_rs[g + 1] = o;
  }
  if (bad) *bad_flag = 1;
}

// ---- op-level GroupGather: count + fill, one warp per group --------------------------------
template <typename T>
__global__ void group_gather_count_kernel(const int64_t* __restrict__ prs, int64_t n_prs,
                                          const int64_t* __restrict__ iv,
                                          const int64_t* __restrict__ irs, int64_t n_groups,
                                          int64_t* __restrict__ group_len, int* __restrict__ bad_flag) {
  const int64_t g = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const unsigned lane = threadIdx.x & 31;
  if (g >= n_groups) return;
  long long sum = 0;
  for (int64_t j = irs[g] + lane; j < irs[g + 1]; j += 32) {
    const int64_t idx = iv[j];
    if (idx < 0 || idx + 1 >= n_prs) { *bad_flag = 1; continue; }
    sum += prs[idx + 1] - prs[idx];
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(FULL, sum, d);
  if (lane == 0) group_len[g] = sum;
}

template <typename T>
__global__ void group_gather_fill_kernel(const T* __restrict__ pv, const int64_t* __restrict__ prs,
                                         const int64_t* __restrict__ iv,
                                         const int64_t* __restrict__ irs, int64_t n_groups,
                                         const int64_t* __restrict__ out_rs, T* __restrict__ out) {
  const int64_t g = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const unsigned lane = threadIdx.x & 31;
  if (g >= n_groups) return;
  int64_t o = out_rs[g];
  for (int64_t j = irs[g]; j < irs[g + 1]; ++j) {
    const int64_t idx = iv[j];
    const int64_t b = prs[idx], e = prs[idx + 1];
    for (int64_t k = b + lane; k < e; k += 32) out[o + (k - b)] = pv[k];
    o += e - b;
  }
}

// unique=true: first-occurrence order per group; one warp per group, scratch bitmap-free
// O(len * distinct/32) compare scan (op-level convenience, not on the exec.pb path).
template <typename T>
__global__ void group_gather_unique_kernel(const T* __restrict__ pv, const int64_t* __restrict__ prs,
                                           const int64_t* __restrict__ iv,
                                           const int64_t* __restrict__ irs, int64_t n_groups,
                                           const int64_t* __restrict__ cap_rs, T* __restrict__ tmp,
                                           int64_t* __restrict__ group_len) {
  const int64_t g = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const unsigned lane = threadIdx.x & 31;
  if (g >= n_groups) return;
  T* out = tmp + cap_rs[g];
  int64_t o = 0;
  for (int64_t j = irs[g]; j < irs[g + 1]; ++j) {
    const int64_t idx = iv[j];
    for (int64_t k = prs[idx]; k < prs[idx + 1]; ++k) {
      const T v = pv[k];
      bool seen = false;
      for (int64_t q = lane; q < o; q += 32) seen |= (out[q] == v);
      seen = __any_sync(FULL, seen);
      if (!seen) {
        if (lane == 0) out[o] = v;
        ++o;
        __syncwarp();
      }
    }
  }
  if (lane == 0) group_len[g] = o;
}

// ---- K1: row gather, 16-byte vector loads, one warp per row chunk -------------------------
// row_bytes must be a multiple of 16 for the vector path; the scalar path handles the rest.
__global__ void gather_rows_vec_kernel(const uint4* __restrict__ table, int64_t n_rows,
                                       int vec_per_row, const int32_t* __restrict__ ids, int64_t n,
                                       uint4* __restrict__ out, int* __restrict__ bad_flag) {
  const int64_t total = n * vec_per_row;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = t / vec_per_row;
    const int c = (int)(t - r * vec_per_row);
    const int64_t id = ids[r];
    if (id < 0 || id >= n_rows) { *bad_flag = 1; continue; }
    uint4 v;
    const uint4* src = table + id * vec_per_row + c;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(src));
    out[t] = v;
  }
}
__global__ void gather_rows_bytes_kernel(const uint8_t* __restrict__ table, int64_t n_rows,
                                         int64_t row_bytes, const int32_t* __restrict__ ids,
                                         int64_t n, uint8_t* __restrict__ out,
                                         int* __restrict__ bad_flag) {
  const int64_t total = n * row_bytes;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = t / row_bytes;
    const int64_t id = ids[r];
    if (id < 0 || id >= n_rows) { *bad_flag = 1; continue; }
    out[t] = table[id * row_bytes + (t - r * row_bytes)];
  }
}

// ---- K5: stable top-k --------------------------------------------------------------------
// Order = TopKV2's stable_comp: larger value first, equal values -> smaller position first
// (-0.0 == +0.0 as in the float compare).  One CTA per row/query.
// The row is the concatenation of segment A (a_n elements) and segment B (b_n elements):
// exactly tf.concat([...]) + TopKV2 + tf.gather(ids, indices) of build_opt_graph.py:52-66,125-127.
__device__ __forceinline__ uint32_t order_key(float f) {
  uint32_t u = __float_as_uint(f);
  if ((u << 1) == 0) u = 0;  // -0.0 -> +0.0
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// Shard exchange fused into the final top-k: the CTA that produced a query's per-shard top-k stores the
// (score, item id) records and the query's status straight into EVERY rank's receive window -- peer HBM mapped
// through CUDA IPC / peer access, i.e. plain st.global over NVLink -- and the last CTA to finish publishes
// "shard `rank` has delivered sequence `seq`" in every window (release at system scope).
constexpr int NANN_MAX_SHARDS = 16;
struct ShardPush {
  int world, rank, k, B;                            // world == 0: no exchange
  float* sc[NANN_MAX_SHARDS];                       // window slot of peer p: scores [B][world][k]
  int64_t* ids[NANN_MAX_SHARDS];                    //                        ids    [B][world][k]
  int32_t* st[NANN_MAX_SHARDS];                     //                        status [B][world]
  unsigned long long* arrive[NANN_MAX_SHARDS];      // peer p's flag "rank delivered seq" for this slot
  const unsigned long long* my_done;                // own window: done[p] = 1 + last sequence rank p has merged
  unsigned long long seq, need_done;                // slot is free once every done[p] >= need_done
  unsigned int* counter;                            // local: CTAs finished (wraps to 0 at B)
  int* error;                                       // local: set when a wait timed out (peer died / calls out of step)
};

struct TopkArgs {
  const float* a_sc; const int32_t* a_ids; int64_t a_stride; int a_n;        // fixed length
  const float* b_sc; int64_t b_sc_stride; const int32_t* b_ids; int64_t b_ids_stride;
  const int32_t* b_n_ptr; int b_n_fixed;                                      // per-row or fixed
  int k;
  float* out_sc; int32_t* out_ids; int32_t* out_pos; int64_t out_stride; int64_t out_offset;
  const int64_t* item_ids; int64_t* out_item_ids; int64_t out_item_stride;    // optional final gather
  int32_t* status;       // per-row status (nullable): rows with status!=0 are skipped
  int reject_single;     // b_n==1 -> InvalidArgument (the tf.squeeze scalar case)
  // ragged form (BatchTopKOnRT): row r = b_sc[b_row_off[r] .. b_row_off[r+1]), k = min(len, k_ptr[r] or k),
  // outputs at out_row_off[r]; group-local positions as int64; ascending flips the value order
  const int64_t* b_row_off; const int64_t* k_ptr; const int64_t* out_row_off; int64_t* out_pos64; int ascending;
  // eval traversal (model.py:255-271): list a has a per-row length, k is clamped to the row length, the
  // number of results per row is reported
  const int32_t* a_n_ptr; int clamp; int32_t* out_n;
  ShardPush push;        // final top-k of a sharded search: results go to the peers' windows as well
};

// system-scope flag helpers of the shard exchange
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long shard_global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#ifndef NANN_SHARD_TIMEOUT_NS
#define NANN_SHARD_TIMEOUT_NS 20000000000ull   // 20 s: a peer that never delivers must not hang the GPU
#endif
// spin until *flag >= want; false on timeout
__device__ __forceinline__ bool shard_wait_flag(const unsigned long long* flag, unsigned long long want) {
  if (ld_acquire_sys_u64(flag) >= want) return true;
  const unsigned long long t0 = shard_global_ns();
  for (;;) {
    __nanosleep(200);
    if (ld_acquire_sys_u64(flag) >= want) return true;
    if (shard_global_ns() - t0 > NANN_SHARD_TIMEOUT_NS) return false;
  }
}

constexpr int TOPK_THREADS = 512;
constexpr int TOPK_MAX_K = 4096;

__device__ __forceinline__ float topk_score_at(const TopkArgs& a, int64_t row, int i, int a_n) {
  if (a.b_row_off) return a.b_sc[a.b_row_off[row] + i];
  return i < a_n ? a.a_sc[row * a.a_stride + i] : a.b_sc[row * a.b_sc_stride + (i - a_n)];
}
__device__ __forceinline__ uint32_t topk_key_at(const TopkArgs& a, int64_t row, int i, int a_n) {
  const uint32_t key = order_key(topk_score_at(a, row, i, a_n));
  return a.ascending ? ~key : key;
}

__device__ __forceinline__ void topk_row(const TopkArgs& a, const int64_t row, unsigned long long* sel) {
  __shared__ int hist[256];
  __shared__ uint32_t s_prefix;
  __shared__ int s_remaining, s_cnt, s_warp_cnt[TOPK_THREADS / 32], s_eq_taken;
  const int tid = threadIdx.x;
  if (a.status && a.status[row] != 0) return;
  const int b_n = a.b_row_off ? (int)(a.b_row_off[row + 1] - a.b_row_off[row])
                              : (a.b_n_ptr ? a.b_n_ptr[row] : a.b_n_fixed);
  const int a_n = a.a_n_ptr ? a.a_n_ptr[row] : a.a_n;
  const int n = a_n + b_n;
  int k = a.k;
  if (a.clamp) k = k < n ? k : n;
  if (a.b_row_off) {  // BatchTopKOnRT: k[g] (or scalar k) clamped to the group length (:117-121)
    const long long kk = a.k_ptr ? a.k_ptr[row] : (long long)a.k;
    k = (int)(kk < 0 ? 0 : (kk < n ? kk : n));
  }
  if (n < k || (a.reject_single && b_n == 1)) {
    if (tid == 0 && a.status) a.status[row] = NANN_INVALID_ARGUMENT;
    return;
  }
  if (a.out_n && tid == 0) a.out_n[row] = k;
  if (k == 0) return;

  // --- radix select (MSD, 8 bits per pass): key of the k-th best -> s_prefix
  uint32_t prefix = 0, mask = 0;
  int remaining = k;
  for (int pass = 3; pass >= 0; --pass) {
    if (tid < 256) hist[tid] = 0;
    __syncthreads();
    const int shift = pass * 8;
    for (int i = tid; i < n; i += TOPK_THREADS) {
      const uint32_t key = topk_key_at(a, row, i, a_n);
      if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255], 1);
    }
    __syncthreads();
    if (tid < 32) {  // warp 0: suffix scan from bin 255 downwards
      int carry = 0, found_bin = -1, found_before = 0;
      for (int blk = 7; blk >= 0; --blk) {
        const int bin = blk * 32 + (31 - tid);  // lane 0 -> highest bin of the block
        const int c = hist[bin];
        int incl = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          int t = __shfl_up_sync(FULL, incl, d);
          if (tid >= d) incl += t;
        }
        const int before = carry + incl - c;  // elements in strictly higher bins
        const bool hit = (before < remaining) && (remaining <= before + c);
        const unsigned hm = __ballot_sync(FULL, hit);
        if (hm && found_bin < 0) {
          const int src = __ffs(hm) - 1;
          found_bin = __shfl_sync(FULL, bin, src);
          found_before = __shfl_sync(FULL, before, src);
        }
        carry += __shfl_sync(FULL, incl, 31);
      }
      if (tid == 0) {
        s_prefix = prefix | ((uint32_t)found_bin << shift);
        s_remaining = remaining - found_before;
      }
    }
    __syncthreads();
    prefix = s_prefix;
    remaining = s_remaining;
    mask |= 255u << shift;
    __syncthreads();
  }
  const uint32_t thr = prefix;   // key of the k-th element
  const int eq_take = remaining; // how many elements with key == thr belong to the top k
  const int n_gt = k - eq_take;

  // --- collect: all keys > thr (any order), then the first eq_take keys == thr by position
  if (tid == 0) { s_cnt = 0; s_eq_taken = 0; }
  __syncthreads();
  for (int i = tid; i < n; i += TOPK_THREADS) {
    const uint32_t key = topk_key_at(a, row, i, a_n);
    if (key > thr) {
      const int slot = atomicAdd(&s_cnt, 1);
      sel[slot] = ((unsigned long long)(~key) << 32) | (uint32_t)i;
    }
  }
  __syncthreads();
  {
    const int lane = tid & 31, wid = tid >> 5;
    for (int base = 0; base < n; base += TOPK_THREADS) {
      if (s_eq_taken >= eq_take) break;  // uniform: read after the barrier below / initial sync
      const int i = base + tid;
      const bool eq = (i < n) && (topk_key_at(a, row, i, a_n) == thr);
      const unsigned bm = __ballot_sync(FULL, eq);
      if (lane == 0) s_warp_cnt[wid] = __popc(bm);
      __syncthreads();
      int before = s_eq_taken;
      for (int w = 0; w < wid; ++w) before += s_warp_cnt[w];
      const int rank = before + __popc(bm & lanemask_lt());
      if (eq && rank < eq_take) sel[n_gt + rank] = ((unsigned long long)(~thr) << 32) | (uint32_t)i;
      __syncthreads();
      if (tid == 0) {
        int tot = 0;
        for (int w = 0; w < TOPK_THREADS / 32; ++w) tot += s_warp_cnt[w];
        s_eq_taken += tot;
      }
      __syncthreads();
    }
  }
  // --- sort the k selected composites ascending = (value desc, position asc); bitonic in smem
  int kpad = 1;
  while (kpad < k) kpad <<= 1;
  for (int i = k + tid; i < kpad; i += TOPK_THREADS) sel[i] = ~0ull;
  __syncthreads();
  for (int size = 2; size <= kpad; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = tid; t < (kpad >> 1); t += TOPK_THREADS) {
        const int lo = ((t / stride) * (stride << 1)) + (t % stride);
        const int hi = lo + stride;
        const bool up = ((lo & size) == 0);
        const unsigned long long x = sel[lo], y = sel[hi];
        if ((x > y) == up) { sel[lo] = y; sel[hi] = x; }
      }
      __syncthreads();
    }
  }
  // --- emit
  for (int r = tid; r < k; r += TOPK_THREADS) {
    const int pos = (int)(uint32_t)(sel[r] & 0xffffffffull);
    const float sc = topk_score_at(a, row, pos, a_n);
    const int64_t o = (a.out_row_off ? a.out_row_off[row] : row * a.out_stride + a.out_offset) + r;
    if (a.out_sc) a.out_sc[o] = sc;
    if (a.out_pos) a.out_pos[o] = pos;
    if (a.out_pos64) a.out_pos64[o] = pos;
    int32_t id = pos;
    if (a.a_ids || a.b_ids)
      id = pos < a_n ? a.a_ids[row * a.a_stride + pos] : a.b_ids[row * a.b_ids_stride + (pos - a_n)];
    if (a.out_ids) a.out_ids[o] = id;
    if (a.out_item_ids) a.out_item_ids[row * a.out_item_stride + r] = a.item_ids[id];
  }
}

__global__ void __launch_bounds__(TOPK_THREADS)
topk_kernel(TopkArgs a) {
  extern __shared__ unsigned long long sel[];  // [kpad] composite keys
  const int64_t row = blockIdx.x;
  topk_row(a, row, sel);
  if (a.push.world == 0) return;
  // ---- shard exchange epilogue: this query's records -> every rank's window, then the delivery flags
  // (the window slot is free: shard_backpressure_kernel ran before this kernel.  The wait is NOT done here: CTAs that
  // spin on the peers' done flags while they fill every SM would keep this GPU's own merge kernel -- which produces one
  // of those flags -- from ever being scheduled.)
  const ShardPush& P = a.push;
  const int tid = threadIdx.x;
  __syncthreads();           // orders the emit loop's writes before the re-reads below (same thread anyway)
  {
    const int32_t st = a.status ? a.status[row] : 0;
    for (int r = tid; r < P.k; r += TOPK_THREADS) {
      const float sc = st == 0 ? a.out_sc[row * a.out_stride + a.out_offset + r] : __int_as_float(-1);
      const int64_t id = st == 0 ? a.out_item_ids[row * a.out_item_stride + r] : (int64_t)-1;
      const int64_t o = (row * P.world + P.rank) * P.k + r;
      for (int p = 0; p < P.world; ++p) { P.sc[p][o] = sc; P.ids[p][o] = id; }
    }
    if (tid < P.world) P.st[tid][row * P.world + P.rank] = st;
  }
  __threadfence_system();    // every thread: its stores are ordered before the counter increment below
  __syncthreads();
  if (tid == 0) {
    const unsigned int prev = atomicInc(P.counter, (unsigned int)(P.B - 1));
    if (prev == (unsigned int)(P.B - 1)) {   // last CTA: all B queries are in every window
      __threadfence_system();
      for (int p = 0; p < P.world; ++p) st_release_sys_u64(P.arrive[p], P.seq + 1);
    }
  }
}

// ---- shard merge: one CTA per query sorts the world*k_in records of its window row (stable: score desc, ties ->
// lower shard, then lower per-shard rank == nann_merge_topk) and emits the best k_out; the last CTA tells every
// peer that this rank is done with the slot.
struct ShardMergeArgs {
  int world, rank, k_in, k_out, B;
  const float* sc; const int64_t* ids; const int32_t* st;       // own window slot
  float* out_sc; int64_t* out_ids; int32_t* out_status;          // [B][k_out], [B]
  unsigned long long* done[NANN_MAX_SHARDS];                    // peer p's done[rank]
  unsigned long long seq;
  unsigned int* counter;
  const int* error;                                              // a wait timed out: every query fails
};
constexpr int SHARD_MERGE_THREADS = 256;
__global__ void __launch_bounds__(SHARD_MERGE_THREADS)
shard_merge_kernel(ShardMergeArgs a) {
  extern __shared__ unsigned long long mkeys[];   // [npad]
  const int64_t row = blockIdx.x;
  const int tid = threadIdx.x;
  const int n = a.world * a.k_in;
  int st = *a.error ? NANN_DEADLINE_EXCEEDED : 0;
  for (int g = 0; g < a.world && st == 0; ++g) st = a.st[row * a.world + g];   // first failing shard decides
  if (st != 0) {             // what a failed single-shard search leaves behind: ids -1, scores 0xFFFFFFFF
    for (int r = tid; r < a.k_out; r += SHARD_MERGE_THREADS) {
      a.out_sc[row * a.k_out + r] = __int_as_float(-1);
      a.out_ids[row * a.k_out + r] = -1;
    }
  } else {
    int npad = 1;
    while (npad < n) npad <<= 1;
    for (int i = tid; i < npad; i += SHARD_MERGE_THREADS)
      mkeys[i] = i < n ? (((unsigned long long)(~order_key(a.sc[row * n + i])) << 32) | (uint32_t)i) : ~0ull;
    __syncthreads();
    for (int size = 2; size <= npad; size <<= 1) {
      for (int stride = size >> 1; stride > 0; stride >>= 1) {
        for (int t = tid; t < (npad >> 1); t += SHARD_MERGE_THREADS) {
          const int lo = ((t / stride) * (stride << 1)) + (t % stride);
          const int hi = lo + stride;
          const bool up = ((lo & size) == 0);
          const unsigned long long x = mkeys[lo], y = mkeys[hi];
          if ((x > y) == up) { mkeys[lo] = y; mkeys[hi] = x; }
        }
        __syncthreads();
      }
    }
    for (int r = tid; r < a.k_out; r += SHARD_MERGE_THREADS) {
      const int pos = (int)(uint32_t)(mkeys[r] & 0xffffffffull);
      a.out_sc[row * a.k_out + r] = a.sc[row * n + pos];
      a.out_ids[row * a.k_out + r] = a.ids[row * n + pos];
    }
  }
  if (tid == 0 && a.out_status) a.out_status[row] = st;
  __syncthreads();           // all reads of the window row are done
  if (tid == 0) {
    __threadfence();
    const unsigned int prev = atomicInc(a.counter, (unsigned int)(a.B - 1));
    if (prev == (unsigned int)(a.B - 1)) {
      __threadfence_system();
      for (int p = 0; p < a.world; ++p) st_release_sys_u64(a.done[p], a.seq + 1);
    }
  }
}
// one warp, before the final top-k of a sharded search: lane p waits until rank p has merged the sequence that used
// this window slot before (done[p] >= need_done), so that the push may overwrite it
__global__ void shard_backpressure_kernel(const unsigned long long* my_done, int world, unsigned long long need_done, int* error) {
  const int p = threadIdx.x;
  if (p < world && !shard_wait_flag(my_done + p, need_done)) *error = 1;
}
// one warp: lane p waits for shard p's delivery of `seq` into this rank's window slot
__global__ void shard_wait_kernel(const unsigned long long* arrive, int world, unsigned long long seq, int* error) {
  const int p = threadIdx.x;
  if (p < world && !shard_wait_flag(arrive + p, seq + 1)) *error = 1;
}

// ---- K6: G-way shard merge: rows of G*k_in (score,id) in allgather layout [G][B][k_in] ------
// Same stable top-k with position = g*k_in + r, i.e. ties -> lower shard, then lower rank.
__global__ void merge_pack_kernel(const float* __restrict__ scores, int G, int B, int k_in,
                                  float* __restrict__ packed /* [B][G*k_in] */) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t tot = (int64_t)G * B * k_in;
  if (t >= tot) return;
  const int r = (int)(t % k_in);
  const int64_t gb = t / k_in;
  const int b = (int)(gb % B), g = (int)(gb / B);
  packed[((int64_t)b * G + g) * k_in + r] = scores[t];
}
__global__ void merge_emit_kernel(const int32_t* __restrict__ pos, const int64_t* __restrict__ ids,
                                  int G, int B, int k_in, int k_out, int64_t* __restrict__ out_ids) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= (int64_t)B * k_out) return;
  const int b = (int)(t / k_out);
  const int p = pos[t];
  const int g = p / k_in, r = p % k_in;
  out_ids[t] = ids[((int64_t)g * B + b) * k_in + r];
}

}  // namespace nann
