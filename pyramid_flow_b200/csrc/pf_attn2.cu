// pf_attn2.cu — masked joint attention forward, two q tiles per CTA (head_dim 64), the default attention kernel.
//
// Same contract as pf_attn.cu (out = softmax(q.k*scale | (seg_q == seg_kv) && (time_q >= time_kv)) . v, reference B:363-365,
// B:596-598, mask F:318-350), restructured after measuring the one-tile kernel (profiles/r01_attn_v2_ncu.txt: XU pipe 62 %,
// tensor pipe 31 %, every softmax row split over two threads that exchange their partial max through shared memory):
//
//   * one CTA per SM owns TWO adjacent 128-row q tiles of one (batch, head) and walks the union of their kv tile lists once:
//     every K/V tile is loaded from L2 once for 256 q rows (half the L2->smem traffic of the one-tile kernel);
//   * softmax warpgroup X (warps 4X..4X+3, 128 threads) owns q tile X: ONE THREAD = ONE FULL ROW of 128 scores in registers.
//     The row max is exact and thread-local: no partner exchange, no shared-memory round trip, no stale reference;
//   * scale-and-subtract and the row sums are packed (fma.rn.f32x2 / add.rn.f32x2 = SASS FFMA2 / FADD2); every exponential is a
//     MUFU.EX2 -- the XU pipe bounds attention at head_dim 64 (16384 ex2 per 128x128 tile at 16/clk/SM = 1024 clk against
//     512 clk of MMA);
//   * O is rescaled in TMEM by the owning thread only when the row max moved by more than 2^8 since the last rescale (the
//     exponent argument is therefore always <= 8: no overflow for any input); the test is branch-free;
//   * the second q tile's softmax warps start `b_delay` clocks late, once per CTA (PF_OPT_ATTN_TILE_PHASE): nothing couples
//     the two warpgroups but the K/V ring, so started together they STAY together -- both exponentiate (sharing the XU), then
//     both load / reduce / store with the XU idle; started out of phase they stay out of phase, and one tile's TMEM loads, row
//     max and P stores run under the other tile's MUFU stream;
//   * warp 8/9 = MMA issuers of tile A/B (one elected lane each), warp 10 = TMA producer (Q once, K through a 4-stage and V
//     through a 3-stage mbarrier ring shared by both tiles), warp 11 allocates TMEM; setmaxnreg moves registers from
//     warps 8-11 (40) to the softmax warpgroups (232).
// What was tried on top and measured slower or equal on B200 is listed in DESIGN.md §3b (polynomial exponentials on the FMA
// pipe, a strict ping-pong token between the tiles, two threads per row, an exponential-only phase, a software-pipelined loop).
// TMEM (512 columns): tile X at column 256 X: S fp32 [0,128) | O fp32 [128,192) | P bf16x2 [192,256).
#include <algorithm>
#include <vector>

#include "pf_attn_pair.cuh"

namespace pf {

// TL = 1: timeline instantiation (per-iteration clock64 stamps of CTA (0, 0, 0): softmax thread 0 of each q tile and the two MMA
// issuers, tools/gpu_check.py attn_timeline); the TL = 0 kernel carries none of it
template <int TL>
__global__ void __launch_bounds__(A2_THREADS, 1)
attn2_fwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                 const __grid_constant__ CUtensorMap tm_v, const Attn2Args a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_q = smem;                                   // 2 tiles
  uint8_t* smem_k = smem + 2 * A2_TILE_BYTES;
  uint8_t* smem_v = smem_k + A2_KSTAGES * A2_TILE_BYTES;

  __shared__ __align__(8) uint64_t bar_q[2], bar_s_full[2], bar_s_free[2], bar_p_full[2], bar_pv_done[2];
  __shared__ __align__(8) uint64_t k_full[A2_KSTAGES], k_empty[A2_KSTAGES], v_full[A2_VSTAGES], v_empty[A2_VSTAGES];
  __shared__ uint32_t tmem_slot;
  __shared__ unsigned long long cta_stamp[4];
  const unsigned long long cta_t0 = a.trace ? clock64() : 0ull;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int pair = blockIdx.x;                              // pair 0 = the last two q tiles (longest kv lists first)
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int bh = b * a.heads + h;
  const int qt_hi = a.q_tiles - 1 - 2 * pair;               // tile B (X = 1)
  const int qt_lo = qt_hi - 1;                              // tile A (X = 0); missing for the first tile of an odd count
  const bool act_lo = qt_lo >= a.q_tile_begin;
  const int n_act = act_lo ? 2 : 1;
  const int* sched = a.psched + (static_cast<size_t>(b) * a.n_pairs + pair) * a.sched_stride;
  const int n_kv = sched[0];

  if (warp == 10 && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
  }
  if (warp == 8 && lane == 0) {
    for (int x = 0; x < 2; ++x) {
      mbar_init(&bar_q[x], 1);
      mbar_init(&bar_s_full[x], 1);
      mbar_init(&bar_s_free[x], 128);
      mbar_init(&bar_p_full[x], 128);
      mbar_init(&bar_pv_done[x], 1);
    }
    for (int i = 0; i < A2_KSTAGES; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], n_act);
    }
    for (int i = 0; i < A2_VSTAGES; ++i) {
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], n_act);
    }
    fence_barrier_init();
  }
  if (warp == 11) {
    tmem_alloc(&tmem_slot, A2_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (a.trace && threadIdx.x == 128) cta_stamp[0] = clock64();

  if (warp >= 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(A2_REGS_OTHER));
    if (warp == 10) {
      if (elect_one()) {
        // ===== TMA producer =====
        if (act_lo) {
          mbar_arrive_expect_tx(&bar_q[0], A2_TILE_BYTES);
          tma_load_3d(smem_q, &tm_q, &bar_q[0], 0, qt_lo * A2_BM, bh);
        }
        mbar_arrive_expect_tx(&bar_q[1], A2_TILE_BYTES);
        tma_load_3d(smem_q + A2_TILE_BYTES, &tm_q, &bar_q[1], 0, qt_hi * A2_BM, bh);
        int ks = 0, vs = 0;
        uint32_t kph = 0, vph = 0;
        for (int j = 0; j < n_kv; ++j) {
          const int kt = sched[1 + j] >> 4;
          mbar_wait(&k_empty[ks], kph ^ 1);
          mbar_arrive_expect_tx(&k_full[ks], A2_TILE_BYTES);
          tma_load_3d(smem_k + ks * A2_TILE_BYTES, &tm_k, &k_full[ks], 0, kt * A2_BN, bh);
          mbar_wait(&v_empty[vs], vph ^ 1);
          mbar_arrive_expect_tx(&v_full[vs], A2_TILE_BYTES);
          tma_load_3d(smem_v + vs * A2_TILE_BYTES, &tm_v, &v_full[vs], 0, kt * A2_BN, bh);
          if (++ks == A2_KSTAGES) { ks = 0; kph ^= 1; }
          if (++vs == A2_VSTAGES) { vs = 0; vph ^= 1; }
        }
      }
    } else if (warp == 8 || warp == 9) {
      const int X = warp - 8;
      if ((X == 1 || act_lo) && elect_one()) {
        // ===== MMA issuer of q tile X (both issuers walk the same kv list; a K/V stage is released when both committed) =====
        constexpr uint32_t idesc_qk = make_idesc_bf16(A2_BM, A2_BN, 0, 0);  // A = Q (K-major), B = K (K-major)
        constexpr uint32_t idesc_pv = make_idesc_bf16(A2_BM, A2_HD, 0, 1);  // A = P (TMEM),    B = V (MN-major)
        const uint32_t t_s = tmem_base + X * A2_TM_TILE + A2_TM_S;
        const uint32_t t_o = tmem_base + X * A2_TM_TILE + A2_TM_O;
        const uint32_t t_p = tmem_base + X * A2_TM_TILE + A2_TM_P;
        mbar_wait(&bar_q[X], 0);
        const uint64_t dq = make_smem_desc_kmajor_sw128(smem_u32(smem_q + X * A2_TILE_BYTES));
        int ks = 0, vs = 0;
        uint32_t kph = 0, vph = 0;
        auto issue_qk = [&]() {
          mbar_wait(&k_full[ks], kph);
          tc_fence_after();
          const uint64_t dk = make_smem_desc_kmajor_sw128(smem_u32(smem_k + ks * A2_TILE_BYTES));
#pragma unroll
          for (int kk = 0; kk < A2_HD / 16; ++kk) umma_ss(t_s, dq + 2 * kk, dk + 2 * kk, idesc_qk, kk != 0);
          umma_commit(&k_empty[ks]);
          umma_commit(&bar_s_full[X]);
          if (++ks == A2_KSTAGES) { ks = 0; kph ^= 1; }
        };
        const bool tl_on = TL && a.timeline != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
        auto tl = [&](int j, int slot) {
          if (TL && tl_on && j < 64) a.timeline[((2 + X) * 64 + j) * 12 + slot] = clock64();
        };
        issue_qk();
        for (int j = 0; j < n_kv; ++j) {
          if (j + 1 < n_kv) {
            mbar_wait(&bar_s_free[X], j & 1);   // S(j) lives in the softmax threads' registers
            tl(j, 0);
            issue_qk();                         // S(j+1) runs on the tensor pipe under softmax(j)
            tl(j, 1);
          }
          mbar_wait(&bar_p_full[X], j & 1);
          tl(j, 2);
          mbar_wait(&v_full[vs], vph);
          tl(j, 3);
          tc_fence_after();
          // V tile [128 kv x 64 hd], 128-byte rows: MN-major, 8-row k groups 1024 B apart, 16 kv rows (2048 B) per MMA
          const uint32_t sv = smem_u32(smem_v + vs * A2_TILE_BYTES);
#pragma unroll
          for (int kk = 0; kk < A2_BN / 16; ++kk) {
            const uint64_t dv = make_smem_desc(sv + kk * 2048, A2_BN * 128, 1024);
            umma_ts(t_o, t_p + kk * 8, dv, idesc_pv, (j | kk) != 0);
          }
          umma_commit(&v_empty[vs]);
          umma_commit(&bar_pv_done[X]);
          tl(j, 4);
          if (++vs == A2_VSTAGES) { vs = 0; vph ^= 1; }
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(A2_REGS_SOFTMAX));
    // ===== softmax + lazy O rescale + epilogue: warpgroup X owns q tile X, thread = one full row =====
    const int X = warp >> 2;
    const int quarter = warp & 3;
    if (X == 1 || act_lo) {
      const int qt = X ? qt_hi : qt_lo;
      const int row = quarter * 32 + lane;
      const int qpos = qt * A2_BM + row;
      const bool q_valid = qpos < a.seq;
      const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
      const uint32_t t_s = tmem_base + lane_base + X * A2_TM_TILE + A2_TM_S;
      const uint32_t t_o = tmem_base + lane_base + X * A2_TM_TILE + A2_TM_O;
      const uint32_t t_p = tmem_base + lane_base + X * A2_TM_TILE + A2_TM_P;
      const float c = a.scale_log2;
      const uint64_t c2 = f2_pack(c, c);
      float m_run = -INFINITY;   // reference max (raw score units) the accumulators are scaled by; -inf: nothing finite yet
      uint64_t l01 = f2_pack(0.f, 0.f), l23 = f2_pack(0.f, 0.f);
      int entry = sched[1];
      const int* mask_idx = a.pmask_idx + (static_cast<size_t>(b) * a.n_pairs + pair) * 2 * a.sched_stride;
      const bool tl_on = TL && a.timeline != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && quarter == 0 && lane == 0;
      auto tl = [&](int j, int slot) {   // TL = 1 only: per-iteration clock stamps of CTA (0, 0, 0) (tools/gpu_check.py attn_timeline)
        if (TL && tl_on && j < 64) a.timeline[(X * 64 + j) * 12 + slot] = clock64();
      };
      if (X == 1 && act_lo && a.b_delay > 0) {
        // De-phase the two q tiles once per CTA.  Nothing couples the two softmax warpgroups but the K/V ring, so they start
        // together and STAY together: both exponentiate (sharing the XU), then both load / reduce / store with the XU idle.
        // Started part of a tile apart they stay apart just the same, and one tile's TMEM loads, max and P stores then run
        // under the other tile's MUFU stream (measured: PF_OPT_ATTN_TILE_PHASE in tools/gpu_check.py attn_phase_sweep).
        mbar_wait(&bar_s_full[X], 0);
        const long long t_begin = clock64();
        while (clock64() - t_begin < a.b_delay) {
        }
      }

      for (int j = 0; j < n_kv; ++j) {
        tl(j, 0);
        const int fl = (entry >> (2 * X)) & 3;              // bit0: this tile has allowed pairs here, bit1: element mask
        const bool own = (fl & 1) != 0;
        const bool masked = !own || (fl & 2) != 0;
        if (j + 1 < n_kv) entry = __ldg(sched + 2 + j);
        // element mask of a partial tile: 128 allow bits per q row, precomputed on the host (pf_attn_build_pair_masks) --
        // one 16-byte load per thread, issued here so its latency hides under the wait for S.  (Building the bits from the
        // seg/time arrays in the kernel cost ~20k clk per partial tile = 20-35 % of the whole kernel:
        // profiles/r02_attn_cta_phases.txt.)
        uint32_t allow0 = 0u, allow1 = 0u, allow2 = 0u, allow3 = 0u;
        if (own && masked) {
          const int blk = __ldg(mask_idx + 2 * j + X);
          const uint4 w = __ldg(a.pmask_bits + static_cast<size_t>(blk) * A2_BM + row);
          allow0 = w.x;
          allow1 = w.y;
          allow2 = w.z;
          allow3 = w.w;
        }
        bool pv_ok = true;
        if (j > 0) pv_ok = mbar_test(&bar_pv_done[X], (j - 1) & 1);    // probed early, consumed before the P store
        mbar_wait(&bar_s_full[X], j & 1);
        tc_fence_after();
        tl(j, 1);
        if (a.trace && j == 0 && threadIdx.x == 128) cta_stamp[1] = clock64();

        // ---- the row's 128 scores: TMEM -> registers, then the tensor pipe may overwrite S with S(j+1)
        uint32_t v0[32], v1[32], v2[32], v3[32];
        tmem_ld32(t_s, v0);      // (one x128 load instead of four x32 measured the same: 2.97 vs 2.93 ms)
        tmem_ld32(t_s + 32, v1);
        tmem_ld32(t_s + 64, v2);
        tmem_ld32(t_s + 96, v3);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&bar_s_free[X]);
        tl(j, 2);
        if (masked) {
          a2_mask32(v0, allow0);
          a2_mask32(v1, allow1);
          a2_mask32(v2, allow2);
          a2_mask32(v3, allow3);
        }
        const float m_tile = fmaxf(fmaxf(a2_max32(v0), a2_max32(v1)), fmaxf(a2_max32(v2), a2_max32(v3)));

        // ---- lazy rescale: move the reference only when the row max grew by more than 2^8 (exponent argument <= 8)
        // (branch-free, same values as the nested ifs it replaces: four data-dependent branches with their FSETP -> BRA latencies
        // sat between the row max and the first exponential: 2.92 -> 2.84 ms per launch at the bench shape)
        const bool first = m_run == -INFINITY;                  // everything accumulated so far is exactly zero
        const bool grow = m_tile > m_run;
        const bool need = grow && !first && (m_tile - m_run) * c > 8.f;
        const float alpha = need ? a2_ex2(fmaxf((m_run - m_tile) * c, -126.f)) : 1.f;
        m_run = (grow && (first || need)) ? m_tile : m_run;
        const float m_ref = (m_run == -INFINITY) ? 0.f : m_run * c;
        const uint64_t nm2 = f2_pack(-m_ref, -m_ref);
        {
          float a0, a1;
          f2_unpack(l01, a0, a1);
          l01 = f2_pack(a0 * alpha, a1 * alpha);
          f2_unpack(l23, a0, a1);
          l23 = f2_pack(a0 * alpha, a1 * alpha);
        }

        tl(j, 3);
        // ---- first half of the row
        uint32_t pk0[16], pk1[16];
        a2_exp64(v0, v1, pk0, pk1, c2, nm2, l01, l23);
        tl(j, 4);
        // ---- P(j-1) consumed and O(j-1) produced before P is overwritten / O is rescaled
        if (j > 0) {
          if (!pv_ok) mbar_wait(&bar_pv_done[X], (j - 1) & 1);
          tc_fence_after();
          if (__any_sync(0xffffffffu, need)) {
#pragma unroll 1
            for (int cc = 0; cc < 64; cc += 16) {
              uint32_t o[16];
              tmem_ld16(t_o + cc, o);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
              tmem_st16(t_o + cc, o);
            }
          }
        }
        tmem_st16(t_p, pk0);
        tmem_st16(t_p + 16, pk1);
        tl(j, 5);
        // ---- second half
        a2_exp64(v2, v3, pk0, pk1, c2, nm2, l01, l23);
        tl(j, 6);
        tmem_st16(t_p + 32, pk0);
        tmem_st16(t_p + 48, pk1);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&bar_p_full[X]);
        tl(j, 7);
      }

      if (a.trace && threadIdx.x == 128) cta_stamp[2] = clock64();
      // ---- epilogue: O / l -> bf16 -> out[b, qpos, h*64 .. +64]
      float s0, s1, s2, s3;
      f2_unpack(l01, s0, s1);
      f2_unpack(l23, s2, s3);
      const float l_run = (s0 + s1) + (s2 + s3);
      mbar_wait(&bar_pv_done[X], (n_kv - 1) & 1);
      tc_fence_after();
      const float inv = (l_run > 0.f) ? 1.f / l_run : 0.f;
      __nv_bfloat16* dst;
      if (a.peer_count > 1) {
        const int r = min(qpos / a.peer_chunk_rows, a.peer_count - 1);
        dst = a.peer_out[r] + static_cast<size_t>(qpos - r * a.peer_chunk_rows) * a.ldo + a.peer_col_begin + h * A2_HD;
      } else {
        dst = a.out + (static_cast<size_t>(b) * a.seq + qpos) * a.ldo + h * A2_HD;
      }
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t o[32];
        tmem_ld32(t_o + hh * 32, o);
        tmem_ld_wait();
        if (q_valid) {
          uint4* d4 = reinterpret_cast<uint4*>(dst + hh * 32);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint4 u;
            u.x = pack_bf16x2(__uint_as_float(o[8 * i + 0]) * inv, __uint_as_float(o[8 * i + 1]) * inv);
            u.y = pack_bf16x2(__uint_as_float(o[8 * i + 2]) * inv, __uint_as_float(o[8 * i + 3]) * inv);
            u.z = pack_bf16x2(__uint_as_float(o[8 * i + 4]) * inv, __uint_as_float(o[8 * i + 5]) * inv);
            u.w = pack_bf16x2(__uint_as_float(o[8 * i + 6]) * inv, __uint_as_float(o[8 * i + 7]) * inv);
            d4[i] = u;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 11) {
    tc_fence_after();
    tmem_dealloc(tmem_base, A2_TMEM_COLS);
  }
  if (a.trace && threadIdx.x == 128) {       // thread 128 = first thread of the upper tile's warpgroup (always active)
    const long long idx = (static_cast<long long>(blockIdx.z) * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    if (idx < a.trace_cap) {
      unsigned smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      unsigned long long* r = a.trace + idx * 8;
      r[0] = cta_t0;
      r[1] = clock64();
      r[2] = static_cast<unsigned long long>(n_kv);
      r[3] = smid;
      r[4] = cta_stamp[0];
      r[5] = cta_stamp[1];
      r[6] = cta_stamp[2];
      r[7] = 0;
    }
  }
}

static unsigned long long* g_a2_trace = nullptr;
static long long g_a2_trace_cap = 0;
static unsigned long long* g_a2_timeline = nullptr;

void attn2_set_trace(unsigned long long* p, long long cap) {
  g_a2_trace = p;
  g_a2_trace_cap = cap;
}
void attn2_set_timeline(unsigned long long* p) { g_a2_timeline = p; }

template <int TL>
static int attn2_launch_t(const CUtensorMap* tm, const Attn2Args& a, dim3 grid, cudaStream_t stream) {
  auto kern = attn2_fwd_kernel<TL>;
  if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(kern), A2_SMEM_BYTES, "attn2_fwd_kernel")) return rc;
  kern<<<grid, A2_THREADS, A2_SMEM_BYTES, stream>>>(tm[0], tm[1], tm[2], a);
  return check_launch("pf_attn_fwd_masked(pair kernel)");
}

int warmup_attn2() {
  int rc = ensure_dyn_smem(reinterpret_cast<const void*>(attn2_fwd_kernel<0>), A2_SMEM_BYTES, "attn2_fwd_kernel<0>");
  if (!rc) rc = ensure_dyn_smem(reinterpret_cast<const void*>(attn2_fwd_kernel<1>), A2_SMEM_BYTES, "attn2_fwd_kernel<1>");
  return rc;
}

// called by pf_attn_fwd_masked (pf_attn.cu) after argument validation
int attn2_launch(const pf_attn_desc* d, cudaStream_t stream) {
  CUtensorMap tm[3];
  const void* ptrs[3] = {d->q, d->k, d->v};
  for (int i = 0; i < 3; ++i) {
    const uint64_t dims[3] = {A2_HD, static_cast<uint64_t>(d->seq), static_cast<uint64_t>(d->batch) * d->heads};
    const uint64_t strides[2] = {A2_HD * 2, static_cast<uint64_t>(d->seq) * A2_HD * 2};
    const uint32_t box[3] = {A2_HD, A2_BN, 1};
    int rc = encode_tensor_map(&tm[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, ptrs[i], dims, strides, box,
                               CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  Attn2Args a{};
  a.out = static_cast<__nv_bfloat16*>(d->out);
  a.ldo = d->ldo;
  a.batch = d->batch;
  a.heads = d->heads;
  a.seq = d->seq;
  a.q_tiles = (d->seq + A2_BM - 1) / A2_BM;
  a.q_tile_begin = d->q_row_begin / A2_BM;
  a.n_pairs = (a.q_tiles + 1) / 2;
  a.scale_log2 = d->scale * 1.4426950408889634f;
  a.seg = d->seg;
  a.time = d->time;
  a.psched = d->pair_sched;
  a.sched_stride = d->sched_stride;
  a.pmask_idx = d->pair_mask_index;
  a.pmask_bits = static_cast<const uint4*>(d->pair_mask_bits);
  a.trace = g_a2_trace;
  a.trace_cap = g_a2_trace_cap;
  a.timeline = g_a2_timeline;
  a.b_delay = get_option(PF_OPT_ATTN_TILE_PHASE);
  a.peer_count = d->peer_count;
  a.peer_chunk_rows = d->peer_chunk_rows;
  a.peer_col_begin = d->peer_col_begin;
  for (int i = 0; i < PF_MAX_PEERS; ++i) a.peer_out[i] = static_cast<__nv_bfloat16*>(d->peer_out[i]);
  if (d->peer_count > 1) {
    if (d->batch != 1 || d->peer_count > PF_MAX_PEERS || d->peer_chunk_rows <= 0 ||
        static_cast<long long>(d->peer_chunk_rows) * d->peer_count < d->seq || d->peer_col_begin % 8 != 0) {
      set_error("pf_attn_fwd_masked: bad peer layout (batch %d, count %d, chunk rows %d, seq %d)", d->batch, d->peer_count,
                d->peer_chunk_rows, d->seq);
      return -1;
    }
    for (int i = 0; i < d->peer_count; ++i)
      if (d->peer_out[i] == nullptr) {
        set_error("pf_attn_fwd_masked: peer_out[%d] is null", i);
        return -1;
      }
  }
  // pair p covers tiles q_tiles-2-2p and q_tiles-1-2p: launch the pairs whose upper tile is >= q_tile_begin
  const int pairs = (a.q_tiles - a.q_tile_begin + 1) / 2;
  dim3 grid(pairs, d->heads, d->batch);
  return a.timeline != nullptr ? attn2_launch_t<1>(tm, a, grid, stream) : attn2_launch_t<0>(tm, a, grid, stream);
}

}  // namespace pf

extern "C" int pf_attn_build_pair_schedule(const int32_t* tile_sched, int32_t batch, int32_t seq, int32_t sched_stride,
                                           int32_t* out) {
  using namespace pf;
  PF_REQUIRE(tile_sched && out && batch > 0 && seq > 0, "pf_attn_build_pair_schedule: bad arguments");
  const int q_tiles = (seq + 127) / 128;
  PF_REQUIRE(sched_stride >= 1 + q_tiles, "pf_attn_build_pair_schedule: stride %d too small", sched_stride);
  const int n_pairs = (q_tiles + 1) / 2;
  for (int b = 0; b < batch; ++b) {
    for (int p = 0; p < n_pairs; ++p) {
      const int hi = q_tiles - 1 - 2 * p, lo = hi - 1;
      const int32_t* rh = tile_sched + (static_cast<size_t>(b) * q_tiles + hi) * sched_stride;
      const int32_t* rl = lo >= 0 ? tile_sched + (static_cast<size_t>(b) * q_tiles + lo) * sched_stride : nullptr;
      int32_t* row = out + (static_cast<size_t>(b) * n_pairs + p) * sched_stride;
      const int nh = rh[0], nl = rl ? rl[0] : 0;
      int ih = 0, il = 0, cnt = 0;
      while (ih < nh || il < nl) {
        const int eh = ih < nh ? rh[1 + ih] : 0x7fffffff, el = il < nl ? rl[1 + il] : 0x7fffffff;
        const int kh = eh >> 1, kl = el >> 1;
        const int kt = std::min(kh, kl);
        int fl = 0, fh = 0;
        if (kl == kt) { fl = 1 | ((el & 1) << 1); ++il; }
        if (kh == kt) { fh = 1 | ((eh & 1) << 1); ++ih; }
        row[1 + cnt] = (kt << 4) | fl | (fh << 2);
        ++cnt;
      }
      row[0] = cnt;
      for (int i = 1 + cnt; i < sched_stride; ++i) row[i] = 0;
    }
  }
  return 0;
}

extern "C" int64_t pf_attn_build_pair_masks(const int32_t* seg, const int32_t* time, const int32_t* pair_sched, int32_t batch,
                                            int32_t seq, int32_t sched_stride, int32_t* mask_index, uint32_t* mask_bits,
                                            int64_t capacity_blocks) {
  using namespace pf;
  if (!seg || !time || !pair_sched || !mask_index || batch <= 0 || seq <= 0) {
    set_error("pf_attn_build_pair_masks: bad arguments");
    return -1;
  }
  const int q_tiles = (seq + 127) / 128;
  const int n_pairs = (q_tiles + 1) / 2;
  int64_t blocks = 0;
  for (int b = 0; b < batch; ++b) {
    const int32_t* sg = seg + static_cast<size_t>(b) * seq;
    const int32_t* tm = time + static_cast<size_t>(b) * seq;
    for (int p = 0; p < n_pairs; ++p) {
      const int32_t* row = pair_sched + (static_cast<size_t>(b) * n_pairs + p) * sched_stride;
      int32_t* mi = mask_index + (static_cast<size_t>(b) * n_pairs + p) * 2 * sched_stride;
      for (int i = 0; i < 2 * sched_stride; ++i) mi[i] = -1;
      const int hi = q_tiles - 1 - 2 * p;
      for (int e = 0; e < row[0]; ++e) {
        const int ent = row[1 + e], kt = ent >> 4;
        for (int x = 0; x < 2; ++x) {
          const int fl = (ent >> (2 * x)) & 3;
          if (fl != 3) continue;                       // needs bits only when the tile owns the entry AND is partial
          const int qt = x ? hi : hi - 1;
          if (mask_bits != nullptr && blocks < capacity_blocks) {
            attn_build_mask_block(sg, tm, seq, qt, kt, mask_bits + static_cast<size_t>(blocks) * 128 * 4);
          }
          mi[2 * e + x] = static_cast<int32_t>(blocks);
          ++blocks;
        }
      }
    }
  }
  return blocks;
}
