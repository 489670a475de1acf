// lib_bloom.inl -- BloomFilterDifference (UO/bitmap_op/bitmap_ops.cc:264-432): the visited-set filter of the
// reference's ragged-batch family with a 4-hash Bloom filter instead of an exact bitmap.
//
// The reference walks all values of all groups in order, one shared flags array: a value is emitted when at least one
// of its four bits was still clear (and all four are set afterwards).  "Was still clear" depends on every value before
// it, so the parallel form is: bit b belongs to the FIRST position that touches it (atomicMin over positions) provided
// it was clear before the call; a value is emitted iff it owns at least one of its bits.  Same output, same flags.
// Fingerprint64 = farmhash (third-party, pinned in tensorflow/workspace.bzl:250-257), restated from the published
// algorithm for inputs of at most 32 bytes (decimal strings of int64 have at most 20).

namespace nann {

#define BF_K0 0xc3a5c85c97cb3127ULL
#define BF_K1 0xb492b66fbe98f273ULL
#define BF_K2 0x9ae16a3b2f90404fULL
__device__ __forceinline__ uint64_t bf_fetch64(const char* p) {
  uint64_t v = 0;
#pragma unroll
  for (int i = 7; i >= 0; --i) v = (v << 8) | (uint8_t)p[i];
  return v;
}
__device__ __forceinline__ uint32_t bf_fetch32(const char* p) {
  return (uint32_t)(uint8_t)p[0] | ((uint32_t)(uint8_t)p[1] << 8) | ((uint32_t)(uint8_t)p[2] << 16) | ((uint32_t)(uint8_t)p[3] << 24);
}
__device__ __forceinline__ uint64_t bf_rot(uint64_t v, int s) { return s == 0 ? v : ((v >> s) | (v << (64 - s))); }
__device__ __forceinline__ uint64_t bf_len16(uint64_t u, uint64_t v, uint64_t mul) {
  uint64_t a = (u ^ v) * mul;
  a ^= (a >> 47);
  uint64_t b = (v ^ a) * mul;
  b ^= (b >> 47);
  return b * mul;
}
__device__ uint64_t bf_fingerprint64(const char* s, int len) {
  if (len <= 16) {
    if (len >= 8) {
      const uint64_t mul = BF_K2 + (uint64_t)len * 2, a = bf_fetch64(s) + BF_K2, b = bf_fetch64(s + len - 8);
      const uint64_t c = bf_rot(b, 37) * mul + a, d = (bf_rot(a, 25) + b) * mul;
      return bf_len16(c, d, mul);
    }
    if (len >= 4) {
      const uint64_t mul = BF_K2 + (uint64_t)len * 2, a = bf_fetch32(s);
      return bf_len16((uint64_t)len + (a << 3), bf_fetch32(s + len - 4), mul);
    }
    if (len > 0) {
      const uint8_t a = (uint8_t)s[0], b = (uint8_t)s[len >> 1], c = (uint8_t)s[len - 1];
      const uint32_t y = (uint32_t)a + ((uint32_t)b << 8), z = (uint32_t)len + ((uint32_t)c << 2);
      const uint64_t m = y * BF_K2 ^ z * BF_K0;
      return (m ^ (m >> 47)) * BF_K2;
    }
    return BF_K2;
  }
  const uint64_t mul = BF_K2 + (uint64_t)len * 2, a = bf_fetch64(s) * BF_K1, b = bf_fetch64(s + 8);
  const uint64_t c = bf_fetch64(s + len - 8) * mul, d = bf_fetch64(s + len - 16) * BF_K2;
  return bf_len16(bf_rot(a + b, 43) + bf_rot(c, 30) + d, a + bf_rot(b + BF_K2, 18) + c, mul);
}
// std::to_string(node) -> buf (no terminator needed); returns the length
__device__ __forceinline__ int bf_to_string(long long node, char* buf) {
  char tmp[24];
  int n = 0;
  unsigned long long u = node < 0 ? 0ull - (unsigned long long)node : (unsigned long long)node;
  do { tmp[n++] = (char)('0' + (int)(u % 10)); u /= 10; } while (u);
  int len = 0;
  if (node < 0) buf[len++] = '-';
  while (n) buf[len++] = tmp[--n];
  return len;
}

struct BloomParams { long long bucket, bucket_size; long long primes[4]; };

// bit ids of value p (:340-351); bits of the filter that were clear before the call are claimed by the first position
template <typename T>
__global__ void bloom_claim_kernel(const T* __restrict__ vals, int64_t n, BloomParams P, const uint32_t* __restrict__ flags,
                                   uint32_t* __restrict__ bit_ids /* [n][4] */, int* __restrict__ first /* [bucket_size*32] */) {
  const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (p >= n) return;
  char buf[32];
  const int len = bf_to_string((long long)vals[p], buf);
  uint64_t raw = bf_fingerprint64(buf, len);
  if (P.bucket > 0) raw = raw % (uint64_t)P.bucket;
  const int mult[4] = {1, 3, 5, 7};
#pragma unroll
  for (int l = 0; l < 4; ++l) {
    const uint64_t lp = (uint64_t)P.primes[l];
    const uint64_t tmp = ((raw * (uint64_t)mult[l]) % lp + lp) % lp;
    const uint32_t b = (uint32_t)(tmp % (uint64_t)(P.bucket_size * 32));
    bit_ids[p * 4 + l] = b;
    if (!(flags[b >> 5] & (1u << (b & 31)))) atomicMin(first + b, (int)p);
  }
}
__global__ void bloom_keep_kernel(int64_t n, const uint32_t* __restrict__ flags, const uint32_t* __restrict__ bit_ids,
                                  const int* __restrict__ first, int32_t* __restrict__ keep) {
  const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (p >= n) return;
  int k = 0;
#pragma unroll
  for (int l = 0; l < 4; ++l) {
    const uint32_t b = bit_ids[p * 4 + l];
    k |= (!(flags[b >> 5] & (1u << (b & 31)))) && first[b] == (int)p;
  }
  keep[p] = k;
}
template <typename T>
__global__ void bloom_emit_kernel(const T* __restrict__ vals, int64_t n, const int32_t* __restrict__ keep, const int64_t* __restrict__ pos,
                                  const uint32_t* __restrict__ bit_ids, uint32_t* __restrict__ flags, T* __restrict__ out,
                                  const int64_t* __restrict__ rs, int64_t n_rs, int64_t* __restrict__ out_rs) {
  const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (p < n_rs) out_rs[p] = pos[rs[p]];
  if (p >= n) return;
  if (keep[p]) out[pos[p]] = vals[p];
#pragma unroll
  for (int l = 0; l < 4; ++l) { const uint32_t b = bit_ids[p * 4 + l]; atomicOr(flags + (b >> 5), 1u << (b & 31)); }
}

static bool bloom_is_prime(long long x) {                      // bitmap_ops.cc:395-402
  for (long long i = (long long)(std::sqrt((double)x) + 1e-6); i > 1; i--)
    if ((x % i) == 0) return false;
  return true;
}

template <typename T>
static nann_status bloom_impl(const T* vals, int64_t n_v, const int64_t* rs, int64_t n_rs, int32_t* flags, int64_t n_flags,
                              int64_t bucket, int64_t bucket_size, nann_alloc_fn alloc, void* ctx, void* stream) {
  NANN_TRY(require_device());
  if (!alloc) return fail(NANN_INVALID_ARGUMENT, "alloc callback is NULL");
  if (bucket < 0 || bucket_size < 1) return fail(NANN_INVALID_ARGUMENT, "bucket >= 0 and bucket_size >= 1 required (attrs, bitmap_ops.cc:272-273)");
  if (n_flags < bucket_size)
    return fail(NANN_INVALID_ARGUMENT, "idx_flag has %lld words, bucket_size is %lld (the reference would write out of bounds)",
                (long long)n_flags, (long long)bucket_size);
  if (bucket_size * 32 > 0x7fffffffll || n_v > 0x7fffffffll) return fail(NANN_UNIMPLEMENTED, "more than 2^31-1 filter bits or values");
  cudaStream_t st = (cudaStream_t)stream;
  DevIn<T> d_v;
  DevIn<int64_t> d_rs;
  NANN_TRY(d_v.init(vals, n_v, st));
  NANN_TRY(d_rs.init(rs, n_rs, st));
  int code = 0;
  NANN_TRY(validate_ragged(n_v, d_rs.d, n_rs, st, &code));
  if (code) return fail(NANN_INVALID_ARGUMENT, "Invalid RaggedTensor input0 a, code: %d", code);   // :307-309
  if (n_rs == 1) return deliver_void(alloc, ctx);                                                   // :312-322
  BloomParams P{};
  P.bucket = bucket; P.bucket_size = bucket_size;
  static const int mod_param[4] = {29, 47, 67, 83};
  for (int i = 0; i < 4; ++i) {                                                                     // :404-421
    long long t = (long long)mod_param[i] * bucket_size * 32;
    while (t > 0 && !bloom_is_prime(t)) --t;
    P.primes[i] = t;
  }
  DevOut<int32_t> d_flags;
  NANN_TRY(d_flags.init(flags, n_flags, st, /*copy_in=*/true));
  const int64_t n_bits = bucket_size * 32;
  DevBuf<uint32_t> d_bits;
  DevBuf<int> d_first;
  DevBuf<int32_t> d_keep;
  DevBuf<int64_t> d_pos, d_ors;
  DevBuf<T> d_out;
  NANN_TRY(d_bits.alloc(std::max<int64_t>(n_v, 1) * 4));
  NANN_TRY(d_first.alloc(n_bits));
  NANN_TRY(d_keep.alloc(std::max<int64_t>(n_v, 1)));
  NANN_TRY(d_pos.alloc(n_v + 1));
  NANN_TRY(d_ors.alloc(n_rs));
  NANN_TRY(d_out.alloc(std::max<int64_t>(n_v, 1)));
  NANN_LAUNCH(fill_words_kernel, 148, 256, 0, st, (uint32_t*)d_first.d, n_bits, 0x7fffffffu);
  const unsigned blocks = (unsigned)ceil_div(std::max<int64_t>(std::max(n_v, n_rs), 1), 256);
  if (n_v > 0) {
    NANN_LAUNCH(bloom_claim_kernel<T>, blocks, 256, 0, st, d_v.d, n_v, P, (const uint32_t*)d_flags.d, d_bits.d, d_first.d);
    NANN_LAUNCH(bloom_keep_kernel, blocks, 256, 0, st, n_v, (const uint32_t*)d_flags.d, d_bits.d, d_first.d, d_keep.d);
  }
  NANN_TRY(device_scan(d_keep.d, n_v, d_pos.d, st));
  NANN_LAUNCH(bloom_emit_kernel<T>, blocks, 256, 0, st, d_v.d, n_v, d_keep.d, d_pos.d, d_bits.d, (uint32_t*)d_flags.d, d_out.d,
              d_rs.d, n_rs, d_ors.d);
  int64_t total = 0;
  NANN_CUDA(cudaMemcpyAsync(&total, d_pos.d + n_v, 8, cudaMemcpyDeviceToHost, st));
  NANN_CUDA(cudaStreamSynchronize(st));
  NANN_TRY(deliver<T>(alloc, ctx, 0, d_out.d, total, st));
  NANN_TRY(deliver<int64_t>(alloc, ctx, 1, d_ors.d, n_rs, st));
  NANN_TRY(d_flags.finish(st));
  NANN_CUDA(cudaStreamSynchronize(st));
  return NANN_OK;
}

}  // namespace nann

extern "C" {
nann_status nann_bloom_filter_difference_i32(const int32_t* v, int64_t n_v, const int64_t* rs, int64_t n_rs, int32_t* flags,
                                             int64_t n_flags, int64_t bucket, int64_t bucket_size, nann_alloc_fn alloc,
                                             void* ctx, void* stream) {
  return bloom_impl<int32_t>(v, n_v, rs, n_rs, flags, n_flags, bucket, bucket_size, alloc, ctx, stream);
}
nann_status nann_bloom_filter_difference_i64(const int64_t* v, int64_t n_v, const int64_t* rs, int64_t n_rs, int32_t* flags,
                                             int64_t n_flags, int64_t bucket, int64_t bucket_size, nann_alloc_fn alloc,
                                             void* ctx, void* stream) {
  return bloom_impl<int64_t>(v, n_v, rs, n_rs, flags, n_flags, bucket, bucket_size, alloc, ctx, stream);
}
}  // extern "C"
