// umma_rate_test.cu — hardware probe (developer tool, B200): cycles per SS-mode tcgen05.mma (M = 128, K = 16 bf16) as a
// function of N when a single thread issues them back to back with fixed descriptors (no per-MMA address arithmetic).
// Answers: is a small-N MMA bound by N (floor = N/2 cycles) or by streaming the 128 x 32 B A operand from shared memory?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/umma_rate_test tools/umma_rate_test.cu && ./tools/umma_rate_test
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t row_bytes) {
  const uint64_t layout = row_bytes == 128 ? 2ull : (row_bytes == 64 ? 4ull : 6ull);
  uint64_t d = (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((8u * row_bytes) >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= layout << 61;
  return d;
}

template <int kUnroll>
__global__ void rate(int n_mma_n, int a_row_bytes, int iters, int n_issuers, unsigned long long* out) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar[4];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(base)[i] = 0;   // zeros: values do not matter
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  // issuer w = lane 0 of warp w (w < n_issuers): its own accumulator columns and its own A / B tiles
  if ((tid & 31) == 0 && warp < n_issuers) {
    const uint64_t ad = make_desc(smem_u32(base) + warp * 16384, (uint32_t)a_row_bytes);
    const uint64_t bd = make_desc(smem_u32(base) + 65536 + warp * 8192, 32u);
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n_mma_n >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t d = tmem + warp * 128;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int u = 0; u < kUnroll; ++u)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(1u) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[warp])) : "memory");
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(smem_u32(&bar[warp])) : "memory");
    out[blockIdx.x * 4 + warp] = (unsigned long long)(clock64() - t0);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

int main() {
  unsigned long long* d_out;
  cudaMalloc(&d_out, 148 * 4 * sizeof(unsigned long long));
  const int smem = 1024 + 100 * 1024;
  cudaFuncSetAttribute(rate<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 512, unroll = 8;
  for (int issuers : {1, 2, 4})
    for (int rb : {128, 64, 32})
      for (int n : {16, 32, 64, 128, 256}) {
        if (issuers * n > 512) continue;
        rate<8><<<148, 128, smem>>>(n, rb, iters, issuers, d_out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 2; }
        unsigned long long h[4];
        cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
        printf("{\"probe\": \"umma_rate\", \"issuers\": %d, \"a_row_bytes\": %d, \"N\": %d, \"cycles_per_mma_per_issuer\": %.1f, \"cycles_per_mma_sm\": %.1f}\n",
               issuers, rb, n, (double)h[0] / (iters * unroll), (double)h[0] / (iters * unroll * issuers));
      }
  return 0;
}
