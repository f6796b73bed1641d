// NOTE: the softmax-stream modes of this probe spill (512-thread launch bound = 128 registers for 128 live scores); their
// "one MUFU per 16 clk per warp" result is an artefact -- see exp_sched_probe.cu for the re-measurement.  The LDTM modes stand.
// Micro-probe (not part of the product): tcgen05.ld throughput per SM as a function of how many warps load at once, alone and
// mixed with the softmax instruction stream (FFMA2 + MUFU.EX2 + FADD2 + F2FP), in SM clocks (clock64), so the answer does not
// depend on the clock the box happens to run at.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_probe tmem_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ float ex2f(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint64_t pk(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) { uint64_t d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ uint32_t packbf(float a, float b) { uint32_t r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a)); return r; }

// MODE bit 0: every warp loads 128 columns (4 x LDTM.x32) per iteration; bit 1: every warp runs the softmax stream on 128 values
template <int MODE>
__global__ void __launch_bounds__(512, 1) probe(unsigned long long* out, int iters, int nwarps, float c, float m) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 128 % 512;
  float v[128];
#pragma unroll
  for (int i = 0; i < 128; ++i) v[i] = 0.001f * (threadIdx.x + i);
  uint64_t l01 = pk(0.f, 0.f), l23 = pk(0.f, 0.f);
  uint32_t acc = 0;
  const uint64_t c2 = pk(c, c), m2 = pk(-m, -m);
  __syncthreads();
  const unsigned long long t0 = clock64();
  if (warp < nwarps) {
    for (int it = 0; it < iters; ++it) {
      if (MODE & 1) {
        uint32_t a[32], b[32], d[32], e[32];
        tmem_ld32(base, a);
        tmem_ld32(base + 32, b);
        tmem_ld32(base + 64, d);
        tmem_ld32(base + 96, e);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (!(MODE & 2)) {
#pragma unroll
          for (int i = 0; i < 32; ++i) acc ^= a[i] ^ b[i] ^ d[i] ^ e[i];
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            v[i] += __uint_as_float(a[i] & 1u);
            v[32 + i] += __uint_as_float(b[i] & 1u);
            v[64 + i] += __uint_as_float(d[i] & 1u);
            v[96 + i] += __uint_as_float(e[i] & 1u);
          }
        }
      }
      if (MODE & 2) {
#pragma unroll
        for (int i = 0; i < 128; i += 2) {
          uint64_t x = fma2(pk(v[i], v[i + 1]), c2, m2);
          float x0, x1;
          upk(x, x0, x1);
          const float p0 = ex2f(x0), p1 = ex2f(x1);
          if (i & 2) l23 = add2(l23, pk(p0, p1)); else l01 = add2(l01, pk(p0, p1));
          acc ^= packbf(p0, p1);
          v[i] = p0 * 0.25f;
          v[i + 1] = p1 * 0.25f;
        }
      }
    }
  }
  const unsigned long long t1 = clock64();
  __syncthreads();
  float s0, s1, s2, s3;
  upk(l01, s0, s1);
  upk(l23, s2, s3);
  float s = s0 + s1 + s2 + s3;
  for (int i = 0; i < 128; ++i) s += v[i];
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  if (s == 12345.678f || acc == 0x12345678u) out[1] = (unsigned long long)s;
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512) : "memory");
  }
}

// Consumer-lag experiment: 128 exponentials per thread per iteration; the row-sum / pack of pair k is forced (by a data
// dependency through an opaque zero) to wait for the MUFU results of pair k + LAG.  LAG = 0 is "consume at once".
template <int LAG>
__global__ void __launch_bounds__(512, 1) lag_probe(unsigned long long* out, int iters, int nwarps, float c, float m, uint32_t zero) {
  const int warp = threadIdx.x >> 5;
  float v[128];
#pragma unroll
  for (int i = 0; i < 128; ++i) v[i] = 0.001f * (threadIdx.x + i);
  uint64_t l01 = pk(0.f, 0.f), l23 = pk(0.f, 0.f);
  uint32_t acc = 0;
  const uint64_t c2 = pk(c, c), m2 = pk(-m, -m);
  __syncthreads();
  const unsigned long long t0 = clock64();
  if (warp < nwarps) {
    for (int it = 0; it < iters; ++it) {
      float p[128];
#pragma unroll
      for (int i = 0; i < 64 + LAG; ++i) {
        if (i < 64) {
          uint64_t x;
          asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(x) : "l"(pk(v[2 * i], v[2 * i + 1])), "l"(c2), "l"(m2));
          float x0, x1;
          upk(x, x0, x1);
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p[2 * i]) : "f"(x0));
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p[2 * i + 1]) : "f"(x1));
        }
        if (i >= LAG) {
          const int k = i - LAG;
          float a = p[2 * k], b = p[2 * k + 1];
          if (LAG > 0 && k + LAG < 64) a = __uint_as_float(__float_as_uint(a) | (__float_as_uint(p[2 * (k + LAG) + 1]) & zero));
          uint64_t s;
          asm("add.rn.f32x2 %0, %1, %2;" : "=l"(s) : "l"(k & 1 ? l23 : l01), "l"(pk(a, b)));
          if (k & 1) l23 = s; else l01 = s;
          uint32_t pkd;
          asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pkd) : "f"(b), "f"(a));
          acc ^= pkd;
          v[2 * k] = a * 0.25f;
          v[2 * k + 1] = b * 0.25f;
        }
      }
    }
  }
  const unsigned long long t1 = clock64();
  __syncthreads();
  float s0, s1, s2, s3;
  upk(l01, s0, s1);
  upk(l23, s2, s3);
  float s = s0 + s1 + s2 + s3;
  for (int i = 0; i < 128; ++i) s += v[i];
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  if (s == 12345.678f || acc == 0x12345678u) out[1] = (unsigned long long)s;
}

template <int LAG>
void run_lag(int nwarps) {
  unsigned long long* out;
  cudaMalloc(&out, 16);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int iters = 2000;
  lag_probe<LAG><<<sms, 512>>>(out, 50, nwarps, 1.0f, 0.5f, 0u);
  lag_probe<LAG><<<sms, 512>>>(out, iters, nwarps, 1.0f, 0.5f, 0u);
  cudaError_t e = cudaDeviceSynchronize();
  unsigned long long h[2] = {0, 0};
  cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
  const double clk = double(h[0]) / iters;
  printf("[tmem_probe] softmax stream, consumer lag %2d pairs, %2d warps: %8.1f clk per 128 exps/thread | ex2 %5.2f /clk/SM (%s)\n", LAG,
         nwarps, clk, 4096.0 * nwarps / clk, cudaGetErrorString(e));
  cudaFree(out);
}

// Packed exponentials: PACKED = 0: two ex2.approx.ftz.f32 per pair; 1: one ex2.approx.ftz.bf16x2 per pair (input packed by a
// cvt.rn.bf16x2.f32, output is already the packed bf16 P; the row sum unpacks it); 2: one ex2.approx.f16x2 per pair.
template <int PACKED>
__global__ void __launch_bounds__(512, 1) packed_probe(unsigned long long* out, int iters, int nwarps, float c, float m) {
  const int warp = threadIdx.x >> 5;
  float v[128];
#pragma unroll
  for (int i = 0; i < 128; ++i) v[i] = 0.001f * (threadIdx.x + i);
  uint64_t l01 = pk(0.f, 0.f), l23 = pk(0.f, 0.f);
  uint32_t acc = 0;
  const uint64_t c2 = pk(c, c), m2 = pk(-m, -m);
  __syncthreads();
  const unsigned long long t0 = clock64();
  if (warp < nwarps) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 64; ++i) {
        uint64_t x;
        asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(x) : "l"(pk(v[2 * i], v[2 * i + 1])), "l"(c2), "l"(m2));
        float x0, x1, p0, p1;
        upk(x, x0, x1);
        uint32_t pkd;
        if (PACKED == 0) {
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p0) : "f"(x0));
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p1) : "f"(x1));
          asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pkd) : "f"(p1), "f"(p0));
        } else if (PACKED == 1) {
          uint32_t xin;
          asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(xin) : "f"(x1), "f"(x0));
          asm("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(pkd) : "r"(xin));
          p0 = __uint_as_float(pkd << 16);
          p1 = __uint_as_float(pkd & 0xffff0000u);
        } else {
          uint32_t xin;
          asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(xin) : "f"(x1), "f"(x0));
          asm("ex2.approx.f16x2 %0, %1;" : "=r"(pkd) : "r"(xin));
          asm("{.reg .f16 lo, hi; mov.b32 {lo, hi}, %2; cvt.f32.f16 %0, lo; cvt.f32.f16 %1, hi;}" : "=f"(p0), "=f"(p1) : "r"(pkd));
        }
        uint64_t s;
        asm("add.rn.f32x2 %0, %1, %2;" : "=l"(s) : "l"(i & 1 ? l23 : l01), "l"(pk(p0, p1)));
        if (i & 1) l23 = s; else l01 = s;
        acc ^= pkd;
        v[2 * i] = p0 * 0.25f;
        v[2 * i + 1] = p1 * 0.25f;
      }
    }
  }
  const unsigned long long t1 = clock64();
  __syncthreads();
  float s0, s1, s2, s3;
  upk(l01, s0, s1);
  upk(l23, s2, s3);
  float s = s0 + s1 + s2 + s3;
  for (int i = 0; i < 128; ++i) s += v[i];
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  if (s == 12345.678f || acc == 0x12345678u) out[1] = (unsigned long long)s;
}

template <int PACKED>
void run_packed(const char* name, int nwarps) {
  unsigned long long* out;
  cudaMalloc(&out, 16);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int iters = 2000;
  packed_probe<PACKED><<<sms, 512>>>(out, 50, nwarps, 1.0f, 0.5f);
  packed_probe<PACKED><<<sms, 512>>>(out, iters, nwarps, 1.0f, 0.5f);
  cudaError_t e = cudaDeviceSynchronize();
  unsigned long long h[2] = {0, 0};
  cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
  const double clk = double(h[0]) / iters;
  printf("[tmem_probe] packed-exp %-22s %2d warps: %8.1f clk per 128 exps/thread | exps %5.2f /clk/SM (%s)\n", name, nwarps, clk,
         4096.0 * nwarps / clk, cudaGetErrorString(e));
  cudaFree(out);
}

template <int MODE>
void run(const char* name, int nwarps) {
  unsigned long long* out;
  cudaMalloc(&out, 16);
  cudaMemset(out, 0, 16);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int iters = 2000;
  probe<MODE><<<sms, 512>>>(out, 50, nwarps, 1.0f, 0.5f);
  probe<MODE><<<sms, 512>>>(out, iters, nwarps, 1.0f, 0.5f);
  cudaError_t e = cudaDeviceSynchronize();
  unsigned long long h[2] = {0, 0};
  cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
  const double clk_per_iter = double(h[0]) / iters;
  // per iteration per warp: 16 KB of TMEM (MODE&1), 128*32 exponentials (MODE&2)
  printf("[tmem_probe] %-28s %2d warps: %8.1f clk per iteration", name, nwarps, clk_per_iter);
  if (MODE & 1) printf(" | TMEM read %6.1f B/clk/SM", 16384.0 * nwarps / clk_per_iter);
  if (MODE & 2) printf(" | ex2 %5.2f /clk/SM", 4096.0 * nwarps / clk_per_iter);
  printf(" (%s)\n", cudaGetErrorString(e));
  cudaFree(out);
}

int main() {
  for (int w : {1, 4, 8, 16}) run<1>("LDTM 4 x x32 only", w);
  for (int w : {4, 8, 16}) run<2>("softmax stream only", w);
  for (int w : {4, 8, 16}) run<3>("LDTM + softmax stream", w);
  for (int w : {4, 8, 16}) {
    run_packed<0>("2 x ex2.f32", w);
    run_packed<1>("1 x ex2.bf16x2", w);
    run_packed<2>("1 x ex2.f16x2", w);
  }
  for (int w : {4, 8}) {
    run_lag<0>(w);
    run_lag<2>(w);
    run_lag<4>(w);
    run_lag<6>(w);
    run_lag<10>(w);
  }
  return 0;
}
