// Micro-probe: MUFU.EX2 throughput per SM for f32, f16x2 and bf16x2 operands (decides whether the attention softmax can
// use packed exponentials).  nvcc -gencode arch=compute_100a,code=sm_100a -o mufu_probe mufu_probe.cu ; ./mufu_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(float* out, int iters) {
  uint32_t r[8];
  for (int i = 0; i < 8; ++i) r[i] = 0x3c003c00u + threadIdx.x + i;   // two small halves / one float pattern
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+r"(r[i]));
      if (MODE == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(r[i]));
      if (MODE == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(r[i]));
    }
  }
  uint32_t s = 0;
  for (int i = 0; i < 8; ++i) s ^= r[i];
  if (s == 0x12345678u) out[0] = 1.f;
}

template <int MODE>
void run(const char* name, int per_instr) {
  int dev_sms = 0, clk = 0;
  cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  float* out;
  cudaMalloc(&out, 4);
  const int iters = 20000, threads = 1024, blocks = dev_sms * 2;
  k<MODE><<<blocks, threads>>>(out, 100);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  cudaEventRecord(a);
  k<MODE><<<blocks, threads>>>(out, iters);
  cudaEventRecord(b);
  cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  const double instr = double(blocks) * threads * iters * 8;
  const double per_sm_per_s = instr / dev_sms / (ms * 1e-3);
  printf("[mufu_probe] %-8s %.3f ms: %.2f thread-instr/ns/SM = %.1f exps/clk/SM at %d MHz nominal (err %s)\n", name, ms,
         per_sm_per_s * 1e-9, per_sm_per_s * per_instr / (clk * 1e3), clk / 1000, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  run<0>("f32", 1);
  run<1>("f16x2", 2);
  run<2>("bf16x2", 2);
  return 0;
}
