// umma_shift_test.cu — hardware probe (developer tool, run on a B200):  can a K-major swizzled UMMA shared-memory
// descriptor start at an arbitrary ROW of a 1024-byte-aligned strip (start address = base + s * row_bytes), so that
// the taps of a convolution become shifted views of one TMA-loaded halo strip?  Also probes K-slices inside a row
// (start + 32*j) for the 64 B / 128 B swizzle modes and an optional base_offset field.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/umma_shift_test tools/umma_shift_test.cu && ./tools/umma_shift_test
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

static constexpr int kRows = 160;  // strip rows in smem (>= 128 + max shift)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t row_bytes, uint32_t base_off) {
  const uint64_t layout = row_bytes == 128 ? 2ull : (row_bytes == 64 ? 4ull : 6ull);
  uint64_t d = (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((8u * row_bytes) >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base_off & 7u) << 49;
  d |= layout << 61;
  return d;
}

// A strip: logical [kRows][row_bytes/2] bf16, value f(row, k); stored with the TMA/UMMA swizzle of `row_bytes`.
// B: [16 n][16 k] bf16 identity, SWIZZLE_32B rows of 32 B (n-major rows, K contiguous).
__global__ void probe(int row_bytes, int shift, int kslice, int use_base_off, float* out /*128x16*/) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  unsigned char* sA = base;                       // kRows * row_bytes <= 20480
  unsigned char* sB = base + 24576;               // 16 * 32 B = 512 B (1024-aligned)
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int kcols = row_bytes / 2;
  const uint32_t mask = row_bytes == 128 ? 7u : (row_bytes == 64 ? 3u : 1u);
  for (int i = tid; i < kRows * kcols; i += blockDim.x) {
    const int r = i / kcols, k = i % kcols;
    const float v = (float)(((r * 7 + k * 3) % 31) - 15);
    uint32_t off = (uint32_t)(r * row_bytes + k * 2);
    off ^= ((off >> 7) & mask) << 4;
    *reinterpret_cast<__nv_bfloat16*>(sA + off) = __float2bfloat16(v);
  }
  for (int i = tid; i < 16 * 16; i += blockDim.x) {
    const int n = i / 16, k = i % 16;
    uint32_t off = (uint32_t)(n * 32 + k * 2);
    off ^= ((off >> 7) & 1u) << 4;
    *reinterpret_cast<__nv_bfloat16*>(sB + off) = __float2bfloat16(n == k ? 1.f : 0.f);
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> visible to the tensor core
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  if (tid == 0) {
    const uint32_t a_addr = smem_u32(sA) + (uint32_t)(shift * row_bytes + kslice * 32);
    const uint32_t boff = use_base_off ? ((a_addr >> 7) & 7u) : 0u;
    const uint64_t ad = make_desc(a_addr, (uint32_t)row_bytes, boff);
    const uint64_t bd = make_desc(smem_u32(sB), 32u, 0u);
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((16u >> 3) << 17) | ((128u >> 4) << 24);
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(0u)
        : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  uint32_t done = 0;
  int spins = 0;
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
    if (++spins > (1 << 22)) __trap();
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t r[16];
  const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  const int m = warp * 32 + lane;
  for (int q = 0; q < 16; ++q) out[m * 16 + q] = __uint_as_float(r[q]);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem) : "memory");
}

int main() {
  float* d_out;
  cudaMalloc(&d_out, 128 * 16 * sizeof(float));
  std::vector<float> h(128 * 16);
  const int smem = 1024 + 24576 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  int total_bad = 0;
  for (int use_bo = 0; use_bo < 2; ++use_bo)
    for (int rb : {32, 64, 128})
      for (int ks = 0; ks < rb / 32; ++ks)
        for (int s = 0; s < 12; ++s) {
          cudaMemset(d_out, 0xff, 128 * 16 * sizeof(float));
          probe<<<1, 128, smem>>>(rb, s, ks, use_bo, d_out);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("CUDA error %s (rb=%d s=%d ks=%d bo=%d)\n", cudaGetErrorString(e), rb, s, ks, use_bo); return 2; }
          cudaMemcpy(h.data(), d_out, h.size() * sizeof(float), cudaMemcpyDeviceToHost);
          int bad = 0;
          for (int m = 0; m < 128; ++m)
            for (int n = 0; n < 16; ++n) {
              const int r = s + m, k = ks * 16 + n;
              const float want = (float)(((r * 7 + k * 3) % 31) - 15);
              if (h[m * 16 + n] != want) ++bad;
            }
          printf("{\"probe\": \"umma_shift\", \"base_off_field\": %d, \"row_bytes\": %d, \"kslice\": %d, \"shift\": %d, \"mismatches\": %d}\n", use_bo, rb, ks, s, bad);
          if (!use_bo) total_bad += bad;
        }
  printf("{\"probe\": \"umma_shift\", \"total_mismatches_without_base_offset\": %d}\n", total_bad);
  return 0;
}
