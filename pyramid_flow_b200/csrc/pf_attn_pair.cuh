// pf_attn_pair.cuh — shared by the two-q-tile attention kernels (pf_attn2.cu: one thread per full row; pf_attn3.cu: two threads
// per row): launch arguments, packed fp32x2 helpers and the exponential loops.
#pragma once

#include "../../include/pf_b200.h"
#include "pf_common.cuh"

namespace pf {

constexpr int A2_BM = 128;
constexpr int A2_BN = 128;
constexpr int A2_HD = 64;
constexpr int A2_KSTAGES = 4;
constexpr int A2_VSTAGES = 3;
constexpr int A2_THREADS = 384;
constexpr int A2_TILE_BYTES = A2_BN * A2_HD * 2;  // 16 KB
constexpr int A2_SMEM_BYTES = (2 + A2_KSTAGES + A2_VSTAGES) * A2_TILE_BYTES + 1024;
constexpr uint32_t A2_TMEM_COLS = 512;
constexpr uint32_t A2_TM_TILE = 256, A2_TM_S = 0, A2_TM_O = 128, A2_TM_P = 192;
constexpr int A2_REGS_SOFTMAX = 232, A2_REGS_OTHER = 40;   // 256*232 + 128*40 = 384*168: exactly the CTA's launch allocation (more would block setmaxnreg.inc forever)

struct Attn2Args {
  __nv_bfloat16* out;
  long long ldo;
  int batch, heads, seq, q_tiles, q_tile_begin, n_pairs;
  float scale_log2;
  const int* seg;
  const int* time;
  const int* psched;
  int sched_stride;
  const int* pmask_idx;        // [batch, n_pairs, 2 * sched_stride]: block of (entry, tile X) in pmask_bits, -1 = none
  const uint4* pmask_bits;     // [blocks, 128 rows]: 128 allow bits of the row over the kv tile (bit i of word w = column 32 w + i)
  // sequence parallel: output rows go straight into the owning rank's buffer (pf_b200.h)
  __nv_bfloat16* peer_out[PF_MAX_PEERS];
  int peer_count, peer_chunk_rows, peer_col_begin;
  unsigned long long* trace;   // debug: per-CTA phase stamps (pf_debug_attn_cta_trace), NULL = off
  long long trace_cap;
  int b_delay;                 // clocks q tile B's softmax warps wait once, after their first S tile arrived, before their first exponential (pf_attn2.cu)
  unsigned long long* timeline;   // debug: per-iteration clock stamps of CTA (0, 0, 0) of the timeline instantiation (pf_debug_attn_trace)
  uint32_t zero;   // always 0, opaque to ptxas: lets the exponential loop express "wait for a later MUFU" as a data dependency
};

// ---- packed fp32x2 helpers (sm_100: FFMA2 / FADD2) -------------------------------------------------------------------
__device__ __forceinline__ uint64_t f2_pack(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void f2_unpack(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t f2_add(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t f2_sub(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// named barriers 1 / 2 = the "XU token" of softmax warpgroup 0 / 1 (ping-pong, see the kernel comment)
__device__ __forceinline__ void a2_token_wait(int id) {
  if (id == 1) asm volatile("bar.sync 1, 256;" ::: "memory");
  else asm volatile("bar.sync 2, 256;" ::: "memory");
}
__device__ __forceinline__ void a2_token_pass(int id) {
  if (id == 1) asm volatile("bar.arrive 1, 256;" ::: "memory");
  else asm volatile("bar.arrive 2, 256;" ::: "memory");
}

// One tcgen05.ld for a whole 128-column fp32 row (4 x 32 registers): a warp's four x32 loads are served one after the other
// (~110 clk each, tools/probes/tmem_probe.cu: 437 clk per 16 KB per warp), one x128 load has a single latency.
__device__ __forceinline__ void tmem_ld128(uint32_t taddr, uint32_t (&v0)[32], uint32_t (&v1)[32], uint32_t (&v2)[32],
                                           uint32_t (&v3)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x128.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63, %64, %65, %66, %67, %68, %69, %70, %71, %72, %73, %74, %75, %76, %77, %78, %79, %80, %81, %82, %83, %84, %85, %86, %87, %88, %89, %90, %91, %92, %93, %94, %95, %96, %97, %98, %99, %100, %101, %102, %103, %104, %105, %106, %107, %108, %109, %110, %111, %112, %113, %114, %115, %116, %117, %118, %119, %120, %121, %122, %123, %124, %125, %126, %127}, [%128];"
      : "=r"(v0[0]), "=r"(v0[1]), "=r"(v0[2]), "=r"(v0[3]), "=r"(v0[4]), "=r"(v0[5]), "=r"(v0[6]), "=r"(v0[7]), "=r"(v0[8]), "=r"(v0[9]), "=r"(v0[10]), "=r"(v0[11]), "=r"(v0[12]), "=r"(v0[13]), "=r"(v0[14]), "=r"(v0[15]), "=r"(v0[16]), "=r"(v0[17]), "=r"(v0[18]), "=r"(v0[19]), "=r"(v0[20]), "=r"(v0[21]), "=r"(v0[22]), "=r"(v0[23]), "=r"(v0[24]), "=r"(v0[25]), "=r"(v0[26]), "=r"(v0[27]), "=r"(v0[28]), "=r"(v0[29]), "=r"(v0[30]), "=r"(v0[31]), "=r"(v1[0]), "=r"(v1[1]), "=r"(v1[2]), "=r"(v1[3]), "=r"(v1[4]), "=r"(v1[5]), "=r"(v1[6]), "=r"(v1[7]), "=r"(v1[8]), "=r"(v1[9]), "=r"(v1[10]), "=r"(v1[11]), "=r"(v1[12]), "=r"(v1[13]), "=r"(v1[14]), "=r"(v1[15]), "=r"(v1[16]), "=r"(v1[17]), "=r"(v1[18]), "=r"(v1[19]), "=r"(v1[20]), "=r"(v1[21]), "=r"(v1[22]), "=r"(v1[23]), "=r"(v1[24]), "=r"(v1[25]), "=r"(v1[26]), "=r"(v1[27]), "=r"(v1[28]), "=r"(v1[29]), "=r"(v1[30]), "=r"(v1[31]), "=r"(v2[0]), "=r"(v2[1]), "=r"(v2[2]), "=r"(v2[3]), "=r"(v2[4]), "=r"(v2[5]), "=r"(v2[6]), "=r"(v2[7]), "=r"(v2[8]), "=r"(v2[9]), "=r"(v2[10]), "=r"(v2[11]), "=r"(v2[12]), "=r"(v2[13]), "=r"(v2[14]), "=r"(v2[15]), "=r"(v2[16]), "=r"(v2[17]), "=r"(v2[18]), "=r"(v2[19]), "=r"(v2[20]), "=r"(v2[21]), "=r"(v2[22]), "=r"(v2[23]), "=r"(v2[24]), "=r"(v2[25]), "=r"(v2[26]), "=r"(v2[27]), "=r"(v2[28]), "=r"(v2[29]), "=r"(v2[30]), "=r"(v2[31]), "=r"(v3[0]), "=r"(v3[1]), "=r"(v3[2]), "=r"(v3[3]), "=r"(v3[4]), "=r"(v3[5]), "=r"(v3[6]), "=r"(v3[7]), "=r"(v3[8]), "=r"(v3[9]), "=r"(v3[10]), "=r"(v3[11]), "=r"(v3[12]), "=r"(v3[13]), "=r"(v3[14]), "=r"(v3[15]), "=r"(v3[16]), "=r"(v3[17]), "=r"(v3[18]), "=r"(v3[19]), "=r"(v3[20]), "=r"(v3[21]), "=r"(v3[22]), "=r"(v3[23]), "=r"(v3[24]), "=r"(v3[25]), "=r"(v3[26]), "=r"(v3[27]), "=r"(v3[28]), "=r"(v3[29]), "=r"(v3[30]), "=r"(v3[31])
      : "r"(taddr)
      : "memory");
}

__device__ __forceinline__ float a2_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float a2_max3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
__device__ __forceinline__ float a2_max32(const uint32_t (&v)[32]) {
  float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    m0 = a2_max3(m0, __uint_as_float(v[i + 0]), __uint_as_float(v[i + 1]));
    m1 = a2_max3(m1, __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
    m2 = a2_max3(m2, __uint_as_float(v[i + 4]), __uint_as_float(v[i + 5]));
    m3 = a2_max3(m3, __uint_as_float(v[i + 6]), __uint_as_float(v[i + 7]));
  }
  return fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
}
__device__ __forceinline__ void a2_mask32(uint32_t (&v)[32], uint32_t bits) {
#pragma unroll
  for (int i = 0; i < 32; ++i)
    if (!((bits >> i) & 1u)) v[i] = 0xff800000u;  // -inf
}

// p = 2^(s*c - m_ref) for 32 scores -> 16 packed bf16x2, row sum into two packed accumulators (4 chains).
// POLY of every 4 pairs take the FMA-pipe path: x = n + f (round to nearest, f in [-0.5, 0.5]), 2^f by a cubic (rel. error
// 6e-4, bf16 P carries 4e-3), n added into the exponent field (LEA).  x <= 8 by construction (lazy-rescale invariant) and is
// clamped at -126 from below; masked tiles (scores of -inf) always take the MUFU path.
// volatile forms: ptxas keeps volatile asm statements in program order, which is how the exponential loop below pins the
// distance between a MUFU and its consumers
__device__ __forceinline__ float a2_ex2_ordered(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint64_t f2_add_ordered(uint64_t a, uint64_t b) {
  uint64_t d;
  asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint32_t pack_bf16x2_ordered(float lo, float hi) {
  uint32_t r;
  asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// Same result as two a2_exp32 calls on (va, vb), software-pipelined by hand: the row sum / bf16 pack of pair i is issued
// A2_EXP_LAG pairs (2 x A2_EXP_LAG MUFU.EX2 = 16 x A2_EXP_LAG XU clocks) after its exponentials.  Left to ptxas the consumers sat
// two MUFUs behind their producers; a warp then stalls on the MUFU latency at every pair, the XU queue drains, and one warp
// keeps the XU pipe only ~50 % busy (tools/probes/tmem_probe.cu: 7.9 ex2/clk/SM with one such warp per SMSP, 11.3 with two;
// attn2 in lockstep: XU 59 %, 0.34 IPC per SMSP, profiles/r02_attn2_lockstep_ncu.txt).  The polynomial pairs' FMA-pipe work
// is not pinned and fills issue slots between MUFUs.
constexpr int A2_EXP_LAG = 0;   // measured: no gain from a forced lag (tools/probes/tmem_probe.cu: a warp's MUFU rate is capped at one per 16 clk whatever the consumer distance)
template <int POLY>
__device__ __forceinline__ void a2_exp64(const uint32_t (&va)[32], const uint32_t (&vb)[32], uint32_t (&pka)[16],
                                         uint32_t (&pkb)[16], uint64_t c2, uint64_t nm2, uint64_t& l01, uint64_t& l23,
                                         uint32_t zero) {
  const uint64_t magic = f2_pack(12582912.f, 12582912.f);   // 1.5 * 2^23
  const uint64_t k3 = f2_pack(0.0555041086648216f, 0.0555041086648216f);
  const uint64_t k2 = f2_pack(0.2402264923172690f, 0.2402264923172690f);
  const uint64_t k1 = f2_pack(0.6931471805599453f, 0.6931471805599453f);
  const uint64_t one = f2_pack(1.f, 1.f);
  float p[64];
#pragma unroll
  for (int i = 0; i < 32 + A2_EXP_LAG; ++i) {
    if (i < 32) {
      const uint32_t s0 = i < 16 ? va[2 * i] : vb[2 * i - 32], s1 = i < 16 ? va[2 * i + 1] : vb[2 * i - 31];
      const uint64_t x = f2_fma(f2_pack(__uint_as_float(s0), __uint_as_float(s1)), c2, nm2);
      float x0, x1;
      f2_unpack(x, x0, x1);
      if ((i & 3) < POLY) {
        const uint64_t xc = f2_pack(fmaxf(x0, -126.f), fmaxf(x1, -126.f));
        const uint64_t t = f2_add(xc, magic);
        const uint64_t f = f2_sub(xc, f2_sub(t, magic));
        uint64_t q = f2_fma(f, k3, k2);
        q = f2_fma(q, f, k1);
        q = f2_fma(q, f, one);
        float q0, q1, t0, t1;
        f2_unpack(q, q0, q1);
        f2_unpack(t, t0, t1);
        p[2 * i] = __int_as_float(__float_as_int(q0) + (__float_as_int(t0) << 23));
        p[2 * i + 1] = __int_as_float(__float_as_int(q1) + (__float_as_int(t1) << 23));
      } else {
        p[2 * i] = a2_ex2(x0);
        p[2 * i + 1] = a2_ex2(x1);
      }
    }
    if (i >= A2_EXP_LAG) {
      const int k = i - A2_EXP_LAG;
      float a = p[2 * k];
      // ptxas schedules by data dependencies only (volatile asm does not pin SASS order): OR-ing in `later & 0` makes pair k's
      // consumers depend on the MUFU of pair k + LAG, so they are placed LAG pairs behind and the XU queue stays full
      if (A2_EXP_LAG > 0 && k + A2_EXP_LAG < 32 && ((k + A2_EXP_LAG) & 3) >= POLY)
        a = __uint_as_float(__float_as_uint(a) | (__float_as_uint(p[2 * (k + A2_EXP_LAG) + 1]) & zero));
      if (k & 1) l23 = f2_add(l23, f2_pack(a, p[2 * k + 1]));
      else l01 = f2_add(l01, f2_pack(a, p[2 * k + 1]));
      if (k < 16) pka[k] = pack_bf16x2(a, p[2 * k + 1]);
      else pkb[k - 16] = pack_bf16x2(a, p[2 * k + 1]);
    }
  }
}

template <int POLY>
__device__ __forceinline__ void a2_exp32(const uint32_t (&v)[32], uint32_t (&pk)[16], uint64_t c2, uint64_t nm2, uint64_t& l01,
                                         uint64_t& l23) {
  const uint64_t magic = f2_pack(12582912.f, 12582912.f);   // 1.5 * 2^23
  const uint64_t k3 = f2_pack(0.0555041086648216f, 0.0555041086648216f);
  const uint64_t k2 = f2_pack(0.2402264923172690f, 0.2402264923172690f);
  const uint64_t k1 = f2_pack(0.6931471805599453f, 0.6931471805599453f);
  const uint64_t one = f2_pack(1.f, 1.f);
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    uint64_t x = f2_fma(f2_pack(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])), c2, nm2);
    float p0, p1;
    if ((i & 3) < POLY) {
      float x0, x1;
      f2_unpack(x, x0, x1);
      x = f2_pack(fmaxf(x0, -126.f), fmaxf(x1, -126.f));
      const uint64_t t = f2_add(x, magic);
      const uint64_t f = f2_sub(x, f2_sub(t, magic));
      uint64_t p = f2_fma(f, k3, k2);
      p = f2_fma(p, f, k1);
      p = f2_fma(p, f, one);
      float t0, t1;
      f2_unpack(p, p0, p1);
      f2_unpack(t, t0, t1);
      p0 = __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23));
      p1 = __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23));
    } else {
      float x0, x1;
      f2_unpack(x, x0, x1);
      p0 = a2_ex2(x0);
      p1 = a2_ex2(x1);
    }
    if (i & 1) l23 = f2_add(l23, f2_pack(p0, p1));
    else l01 = f2_add(l01, f2_pack(p0, p1));
    pk[i] = pack_bf16x2(p0, p1);
  }
}

}  // namespace pf
