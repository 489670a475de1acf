// Micro-benchmark: tcgen05.mma fed by a TMA ring (producer warp / MMA warp / mbarrier full+empty per slot),
// the skeleton of the scorer's layer-2 loop.  Reports cycles per MMA for a few issue structures.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_ring mma_ring.cu && ./mma_ring
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t a) {
  return (uint64_t)((a >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)64 << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ uint32_t idesc(int M, int N) { return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred;
}
__device__ __forceinline__ void mma(uint32_t d, uint64_t ad, uint64_t bd, uint32_t id) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(ad), "l"(bd), "r"(id), "r"(1u) : "memory");
}
__device__ __forceinline__ void commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void wait_bar(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void wait_bar_warp(uint32_t bar, uint32_t parity) {
  if ((threadIdx.x & 31) == 0) wait_bar(bar, parity);
  __syncwarp();
}

constexpr int A_BYTES = 32768, STAGE = 32768, NSLOT = 5;

// MODE 0: single-thread MMA warp (lane 0 does everything)      MODE 1: warp-uniform + elect
// PER : MMAs per ring stage (8 or 4 alternate when PER == 12: 8 then 4)
template <int MODE, int NMMA>
__global__ void __launch_bounds__(128, 1) k(long long* out, int stages, const uint8_t* src) {
  extern __shared__ uint8_t raw[];
  uint8_t* sm = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full[NSLOT], empty[NSLOT], fin;
  __shared__ uint32_t slot_;
  for (int i = threadIdx.x; i < A_BYTES / 4; i += 128) ((uint32_t*)sm)[i] = 0x3c003c00u;
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot_)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < NSLOT; ++i) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[i])));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&empty[i])));
    }
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&fin)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot_;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t a = smem_u32(sm), ring = smem_u32(sm + A_BYTES);
  const uint32_t id = idesc(128, 256);
  if (warp == 0) {           // producer
    if (MODE == 0) {
      if (lane == 0)
        for (int i = 0; i < stages; ++i) {
          const int s = i % NSLOT; const uint32_t ph = (i / NSLOT) & 1;
          wait_bar(smem_u32(&empty[s]), ph ^ 1);
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[s])), "r"((uint32_t)STAGE) : "memory");
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(ring + s * STAGE), "l"(src + (size_t)((i * 7 + blockIdx.x) % 64) * STAGE), "r"((uint32_t)STAGE), "r"(smem_u32(&full[s])) : "memory");
        }
    } else {
      for (int i = 0; i < stages; ++i) {
        const int s = i % NSLOT; const uint32_t ph = (i / NSLOT) & 1;
        wait_bar_warp(smem_u32(&empty[s]), ph ^ 1);
        if (elect_one()) {
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[s])), "r"((uint32_t)STAGE) : "memory");
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(ring + s * STAGE), "l"(src + (size_t)((i * 7 + blockIdx.x) % 64) * STAGE), "r"((uint32_t)STAGE), "r"(smem_u32(&full[s])) : "memory");
        }
      }
    }
  } else if (warp == 1) {    // MMA issuer
    const long long t0 = clock64();
    if (MODE == 0) {
      if (lane == 0) {
        for (int i = 0; i < stages; ++i) {
          const int s = i % NSLOT; const uint32_t ph = (i / NSLOT) & 1;
          wait_bar(smem_u32(&full[s]), ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t b = ring + s * STAGE;
#pragma unroll
          for (int j = 0; j < NMMA; ++j) mma(tmem + (i & 1) * 256, desc_sw128(a + (j & 1) * 16384 + (j >> 1 & 3) * 32), desc_sw128(b + (j >> 1 & 3) * 32), id);
          commit(smem_u32(&empty[s]));
        }
        commit(smem_u32(&fin));
      }
    } else {
      for (int i = 0; i < stages; ++i) {
        const int s = i % NSLOT; const uint32_t ph = (i / NSLOT) & 1;
        wait_bar_warp(smem_u32(&full[s]), ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t b = ring + s * STAGE;
        if (elect_one()) {
#pragma unroll
          for (int j = 0; j < NMMA; ++j) mma(tmem + (i & 1) * 256, desc_sw128(a + (j & 1) * 16384 + (j >> 1 & 3) * 32), desc_sw128(b + (j >> 1 & 3) * 32), id);
          commit(smem_u32(&empty[s]));
        }
      }
      if (elect_one()) commit(smem_u32(&fin));
    }
    if (lane == 0) {
      wait_bar(smem_u32(&fin), 0);
      if (blockIdx.x == 0) out[0] = clock64() - t0;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

template <int MODE, int NMMA>
void run(const char* name, int grid, const uint8_t* src) {
  long long* d;
  cudaMalloc(&d, 8);
  const int stages = 200, smem = A_BYTES + NSLOT * STAGE + 1024;
  cudaFuncSetAttribute(k<MODE, NMMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k<MODE, NMMA><<<grid, 128, smem>>>(d, stages, src);
  k<MODE, NMMA><<<grid, 128, smem>>>(d, stages, src);
  long long h = 0;
  cudaError_t e = cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("%-40s grid=%3d  %7.1f cyc/stage  %6.1f cyc/mma  ring %5.1f B/cyc (%s)\n", name, grid, (double)h / stages, (double)h / stages / NMMA,
         (double)STAGE * stages / (double)h, cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  uint8_t* src;
  cudaMalloc(&src, 64 * STAGE);
  cudaMemset(src, 0, 64 * STAGE);
  for (int grid : {1, 148}) {
    run<0, 8>("single-thread, 8 MMA/stage", grid, src);
    run<1, 8>("elect uniform, 8 MMA/stage", grid, src);
    run<0, 4>("single-thread, 4 MMA/stage", grid, src);
    run<1, 4>("elect uniform, 4 MMA/stage", grid, src);
    run<1, 2>("elect uniform, 2 MMA/stage", grid, src);
    run<1, 16>("elect uniform, 16 MMA/stage", grid, src);
  }
  return 0;
}
