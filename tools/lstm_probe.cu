// lstm_probe.cu — where does a recurrent step's time go?  Times lstm_recurrent4_kernel<kDbg> (csrc/lstm.cu) with parts
// switched off on synthetic buffers (developer tool, B200).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -o tools/lstm_probe tools/lstm_probe.cu
#include "../dcs-net_b200/csrc/lstm.cu"
#include <stdarg.h>
#include <vector>
// stubs for the symbols lstm.cu links against inside the library
namespace dcs {
std::atomic<uint64_t> g_launches{0};
int set_error(int code, const char* fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); fprintf(stderr, "\n"); return code; }
}
extern "C" int dcs_cconv2d_fwd(const dcs_cconv_params*, void*) { return -1; }
extern "C" int dcs_cconv2d_tc_fwd(const dcs_cconv_params*, void*) { return -1; }

template <int kDbg>
static float run(const float* pre, const float* whh, float* hout, int B, int S) {
  const int64_t rows2 = 2ll * B * S;
  dim3 grid(4 * B / 4, 2);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaFuncSetAttribute(dcs::lstm_recurrent4_kernel<kDbg>, cudaFuncAttributeMaxDynamicSharedMemorySize, dcs::kRec4Smem);
  for (int i = 0; i < 2; ++i) dcs::lstm_recurrent4_kernel<kDbg><<<grid, 256, dcs::kRec4Smem>>>(pre, 2 * rows2 * 256, rows2 * 256, 256, whh, hout, B, S);
  cudaEventRecord(e0);
  for (int i = 0; i < 5; ++i) dcs::lstm_recurrent4_kernel<kDbg><<<grid, 256, dcs::kRec4Smem>>>(pre, 2 * rows2 * 256, rows2 * 256, 256, whh, hout, B, S);
  cudaEventRecord(e1);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); exit(2); }
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms / 5;
}

static float run_mma(const float* pre, const float* whh_frag, float* hout, int B, int S) {
  const int64_t rows2 = 2ll * B * S;
  dim3 grid(4 * B / 4, 2);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaFuncSetAttribute(dcs::lstm_recurrent4_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dcs::kRec4Smem);
  for (int i = 0; i < 2; ++i) dcs::lstm_recurrent4_mma_kernel<<<grid, 256, dcs::kRec4Smem>>>(pre, 2 * rows2 * 256, rows2 * 256, 256, whh_frag, hout, B, S);
  cudaEventRecord(e0);
  for (int i = 0; i < 5; ++i) dcs::lstm_recurrent4_mma_kernel<<<grid, 256, dcs::kRec4Smem>>>(pre, 2 * rows2 * 256, rows2 * 256, 256, whh_frag, hout, B, S);
  cudaEventRecord(e1);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); exit(2); }
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms / 5;
}

int main(int argc, char** argv) {
  const int B = 64, S = 500;
  const int64_t rows2 = 2ll * B * S;
  float *pre, *whh, *hout;
  cudaMalloc(&pre, 4 * rows2 * 256 * sizeof(float));
  cudaMalloc(&whh, 2 * 4 * 256 * 64 * sizeof(float));
  cudaMalloc(&hout, 4ll * B * S * 128 * sizeof(float));
  std::vector<float> h(4 * rows2 * 256);
  for (size_t i = 0; i < h.size(); ++i) h[i] = 0.01f * (float)((i * 2654435761u) % 201) - 1.f;
  cudaMemcpy(pre, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice);
  cudaMemcpy(whh, h.data(), 2 * 4 * 256 * 64 * sizeof(float), cudaMemcpyHostToDevice);
  if (argc > 1) {  // "mma": only the tensor-core recurrent kernel (for ncu)
    printf("{\"probe\": \"lstm_mma\", \"ms\": %.4f}\n", run_mma(pre, whh, hout, B, S));
    return 0;
  }
  printf("{\"probe\": \"lstm\", \"B\": %d, \"S\": %d", B, S);
  printf(", \"mma_ms\": %.4f", run_mma(pre, whh, hout, B, S));
  printf(", \"full_ms\": %.4f", run<0>(pre, whh, hout, B, S));
  printf(", \"no_pre_loads_ms\": %.4f", run<1>(pre, whh, hout, B, S));
  printf(", \"no_h_stores_ms\": %.4f", run<2>(pre, whh, hout, B, S));
  printf(", \"no_transcendentals_ms\": %.4f", run<4>(pre, whh, hout, B, S));
  printf(", \"no_fma_ms\": %.4f", run<8>(pre, whh, hout, B, S));
  printf(", \"no_shuffles_ms\": %.4f", run<16>(pre, whh, hout, B, S));
  printf(", \"no_loads_stores_ms\": %.4f", run<3>(pre, whh, hout, B, S));
  printf(", \"only_sync_ms\": %.4f", run<31>(pre, whh, hout, B, S));
  printf("}\n");
  return 0;
}
