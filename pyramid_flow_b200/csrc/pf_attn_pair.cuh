// pf_attn_pair.cuh — the two-q-tile attention kernel (pf_attn2.cu): launch arguments, packed fp32x2 helpers, the row max and the
// exponential loop.
#pragma once

#include "../../include/pf_b200.h"
#include "pf_common.cuh"
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

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

// p = 2^(s*c - m_ref) for 64 scores (two 32-column TMEM loads) -> 32 packed bf16x2, row sum into two packed accumulators
// (4 chains).  Scale-and-subtract and the sums are packed (FFMA2 / FADD2); ptxas places each pair's sum / pack one pair behind its
// two MUFU.EX2, which keeps one warp at 8.96 clk per MUFU and two warps per SMSP at 95 % of the XU's rate
// (tools/probes/exp_sched_probe.cu, profiles/r02_exp_sched_probe.txt).
__device__ __forceinline__ void a2_exp64(const uint32_t (&va)[32], const uint32_t (&vb)[32], uint32_t (&pka)[16],
                                         uint32_t (&pkb)[16], uint64_t c2, uint64_t nm2, uint64_t& l01, uint64_t& l23) {
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const uint32_t s0 = i < 16 ? va[2 * i] : vb[2 * i - 32], s1 = i < 16 ? va[2 * i + 1] : vb[2 * i - 31];
    const uint64_t x = f2_fma(f2_pack(__uint_as_float(s0), __uint_as_float(s1)), c2, nm2);
    float x0, x1;
    f2_unpack(x, x0, x1);
    const float p0 = a2_ex2(x0), p1 = a2_ex2(x1);
    if (i & 1) l23 = f2_add(l23, f2_pack(p0, p1));
    else l01 = f2_add(l01, f2_pack(p0, p1));
    if (i < 16) pka[i] = pack_bf16x2(p0, p1);
    else pkb[i - 16] = pack_bf16x2(p0, p1);
  }
}


// Host: the 128 x 128 allow bits of one (q tile, kv tile) block: bit i of word w of row r = q row qt*128 + r may attend kv column
// kt*128 + 32 w + i  <=>  both inside the sequence, same segment, time_kv <= time_q (mask definition F:318-350).  A plan of the
// 768p run holds a few hundred such blocks; four columns per SSE2 compare (a scalar loop was 0.2 ms per block).
inline void attn_build_mask_block(const int32_t* sg, const int32_t* tm, int seq, int qt, int kt, uint32_t* blk) {
  alignas(16) int32_t sgk[128], tmk[128];
  uint32_t valid[4] = {0u, 0u, 0u, 0u};
  for (int c = 0; c < 128; ++c) {
    const int kv = kt * 128 + c;
    const bool in = kv < seq;
    sgk[c] = in ? sg[kv] : 0;
    tmk[c] = in ? tm[kv] : 0;
    if (in) valid[c >> 5] |= 1u << (c & 31);
  }
  for (int r = 0; r < 128; ++r) {
    const int q = qt * 128 + r;
    uint32_t w[4] = {0u, 0u, 0u, 0u};
    if (q < seq) {
      const int32_t sq = sg[q], tq = tm[q];
#if defined(__SSE2__)
      const __m128i sq4 = _mm_set1_epi32(sq), tq4 = _mm_set1_epi32(tq);
      for (int c = 0; c < 128; c += 4) {
        const __m128i eq = _mm_cmpeq_epi32(_mm_load_si128(reinterpret_cast<const __m128i*>(sgk + c)), sq4);
        const __m128i gt = _mm_cmpgt_epi32(_mm_load_si128(reinterpret_cast<const __m128i*>(tmk + c)), tq4);
        const uint32_t m = static_cast<uint32_t>(_mm_movemask_ps(_mm_castsi128_ps(_mm_andnot_si128(gt, eq))));
        w[c >> 5] |= m << (c & 31);
      }
#else
      for (int c = 0; c < 128; ++c) w[c >> 5] |= static_cast<uint32_t>(sgk[c] == sq && tmk[c] <= tq) << (c & 31);
#endif
      for (int k = 0; k < 4; ++k) w[k] &= valid[k];
    }
    for (int k = 0; k < 4; ++k) blk[r * 4 + k] = w[k];
  }
}

}  // namespace pf
