// pf_attn4.cu -- masked joint attention forward, two q tiles per CTA, one thread per score row, exponentials in an "E phase".
//
// Same shell and contract as pf_attn2.cu (one CTA per SM owns two adjacent 128-row q tiles of one (batch, head), walks the union
// of their kv tile lists once; S = Q.K^T as SS MMAs, O += P.V as TS MMAs with P in TMEM; host-built pair schedule and row masks;
// exact thread-local row max, lazy O rescale).  What changes is the order of the softmax instructions inside a warp, after
// reading pf_attn2's SASS (tools/sass_sched.py) and measuring one warp's exponential stream (tools/probes/exp_sched_probe.cu):
//
//   * ptxas placed each pair's consumers (FADD2 row sum, F2FP pack) one pair behind its two MUFU.EX2.  A MUFU result arrives
//     ~48 clk after issue; the consumer's scoreboard wait stalls the (in-order) warp, and its MUFU rate drops to one per 16 clk
//     -- half the XU's rate -- which is what capped pf_attn2 at 70 % XU / 2906 clk per pair of kv tiles.
//   * here the 128 scores of a row become their exponentials IN PLACE in an "E phase" that contains no consumer of a MUFU
//     result (scale/subtract FFMA2 -> MUFU.EX2, or the FMA-pipe Cody-Waite + cubic for POLY8 of every 8 pairs), closed by a
//     never-taken exit on which all 128 results are live (the one construct ptxas neither schedules across nor sinks below);
//     the row sums and bf16 packs follow in a "C phase".  A warp then issues MUFU.EX2 every 8 clk.
//   * the two softmax warpgroups (q tile A / B) alternate their E phases with a token (named barriers): one warp alone now
//     saturates its SMSP's XU, so tile B's TMEM loads, max, C phase and P stores run under tile A's E phase and vice versa.
#include "pf_attn_pair.cuh"

namespace pf {

// E phase of one row: v <- 2^(v*c - m_ref) in place.  POLY8 of every 8 pairs take the FMA pipe (x = n + f, cubic for 2^f, n
// added into the exponent field); x <= 8 by the lazy-rescale invariant and is clamped at -126 from below.  Returns a predicate
// that is never true but depends on this row's data (so the exit it guards cannot be hoisted or folded).
template <int POLY8>
__device__ __forceinline__ bool a4_exp_inplace(uint32_t (&v0)[32], uint32_t (&v1)[32], uint32_t (&v2)[32], uint32_t (&v3)[32],
                                               uint64_t c2, uint64_t nm2) {
  const uint64_t magic = f2_pack(12582912.f, 12582912.f);   // 1.5 * 2^23
  const uint64_t k3 = f2_pack(0.0555041086648216f, 0.0555041086648216f);
  const uint64_t k2 = f2_pack(0.2402264923172690f, 0.2402264923172690f);
  const uint64_t k1 = f2_pack(0.6931471805599453f, 0.6931471805599453f);
  const uint64_t one = f2_pack(1.f, 1.f);
  bool never = false;
#pragma unroll
  for (int i = 0; i < 64; ++i) {
    uint32_t& s0 = i < 16 ? v0[2 * i] : i < 32 ? v1[2 * i - 32] : i < 48 ? v2[2 * i - 64] : v3[2 * i - 96];
    uint32_t& s1 = i < 16 ? v0[2 * i + 1] : i < 32 ? v1[2 * i - 31] : i < 48 ? v2[2 * i - 63] : v3[2 * i - 95];
    const uint64_t x = f2_fma(f2_pack(__uint_as_float(s0), __uint_as_float(s1)), c2, nm2);
    float x0, x1;
    f2_unpack(x, x0, x1);
    if (i == 63) never = (x0 == 3.0e38f);
    // polynomial pairs spread evenly over every 8 (Bresenham): POLY8 = 2 -> pairs 3 and 7, 3 -> 2, 5, 7, ...
    if ((((i & 7) + 1) * POLY8) / 8 != ((i & 7) * POLY8) / 8) {
      const uint64_t xc = f2_pack(fmaxf(x0, -126.f), fmaxf(x1, -126.f));
      const uint64_t t = f2_add(xc, magic);
      const uint64_t f = f2_sub(xc, f2_sub(t, magic));
      uint64_t q = f2_fma(f, k3, k2);
      q = f2_fma(q, f, k1);
      q = f2_fma(q, f, one);
      float q0, q1, t0, t1;
      f2_unpack(q, q0, q1);
      f2_unpack(t, t0, t1);
      s0 = static_cast<uint32_t>(__float_as_int(q0) + (__float_as_int(t0) << 23));
      s1 = static_cast<uint32_t>(__float_as_int(q1) + (__float_as_int(t1) << 23));
    } else {
      s0 = __float_as_uint(a2_ex2(x0));
      s1 = __float_as_uint(a2_ex2(x1));
    }
  }
  return never;
}

// the never-taken exit of the E phase: every exponential is live here
__device__ __forceinline__ void a4_dump(__nv_bfloat16* out, const uint32_t (&v0)[32], const uint32_t (&v1)[32],
                                     const uint32_t (&v2)[32], const uint32_t (&v3)[32]) {
  volatile uint32_t* p = reinterpret_cast<volatile uint32_t*>(out);
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    p[i] = v0[i];
    p[32 + i] = v1[i];
    p[64 + i] = v2[i];
    p[96 + i] = v3[i];
  }
}

// C phase of 32 exponentials: row sum into two packed accumulators (4 chains), bf16 pack
__device__ __forceinline__ void a4_sum_pack(const uint32_t (&v)[32], uint32_t (&pk)[16], uint64_t& l01, uint64_t& l23) {
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float p0 = __uint_as_float(v[2 * i]), p1 = __uint_as_float(v[2 * i + 1]);
    if (i & 1) l23 = f2_add(l23, f2_pack(p0, p1));
    else l01 = f2_add(l01, f2_pack(p0, p1));
    pk[i] = pack_bf16x2(p0, p1);
  }
}

// TL = 1: timeline instantiation (per-iteration clock64 stamps of CTA (0, 0, 0): softmax thread 0 of each q tile and the two MMA
// issuers); the TL = 0 kernels carry none of it
template <int POLY8, int PINGPONG, int TL>
__global__ void __launch_bounds__(A2_THREADS, 1)
attn4_fwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
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
      const int seg_q = q_valid ? a.seg[static_cast<size_t>(b) * a.seq + qpos] : -0x7fffffff;
      const int time_q = q_valid ? a.time[static_cast<size_t>(b) * a.seq + qpos] : -0x7fffffff;
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
      // Ping-pong: the exponential phase (the XU-bound part) of the two warpgroups is strictly alternated with a token passed
      // through named barriers.  Left alone the two q tiles fall into lockstep (both S tiles become ready together), contend
      // for the XU during their exps and leave it idle while both load / reduce / store: measured XU pipe 59 % busy, the same
      // as the one-tile kernel (profiles/r02_attn2_lockstep_ncu.txt).  With the token one warpgroup's TMEM loads, max and P
      // store run under the other's exponentials.
      const bool pingpong = PINGPONG && act_lo;
      if (pingpong && X == 1) a2_token_pass(1);
      const bool tl_on = TL && a.timeline != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && quarter == 0 && lane == 0;
      auto tl = [&](int j, int slot) {
        if (TL && tl_on && j < 64) a.timeline[(X * 64 + j) * 12 + slot] = clock64();
      };

      for (int j = 0; j < n_kv; ++j) {
        tl(j, 0);
        const int kt = entry >> 4;
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
        float alpha = 1.f;
        bool need = false;
        if (m_tile > m_run) {
          if (m_run == -INFINITY) {
            m_run = m_tile;                      // everything accumulated so far is exactly zero
          } else if ((m_tile - m_run) * c > 8.f) {
            need = true;
            alpha = a2_ex2(fmaxf((m_run - m_tile) * c, -126.f));
            m_run = m_tile;
          }
        }
        const float m_ref = (m_run == -INFINITY) ? 0.f : m_run * c;
        const uint64_t nm2 = f2_pack(-m_ref, -m_ref);
        if (need) {
          float a0, a1;
          f2_unpack(l01, a0, a1);
          l01 = f2_pack(a0 * alpha, a1 * alpha);
          f2_unpack(l23, a0, a1);
          l23 = f2_pack(a0 * alpha, a1 * alpha);
        }

        // ---- E phase: every score becomes its exponential IN PLACE; nothing in this block consumes a MUFU result, so the warp
        // issues MUFU.EX2 back to back (one per 8 clk = the XU's rate) with the FMA-pipe polynomial pairs in the gaps
        tl(j, 3);
        if (pingpong) a2_token_wait(1 + X);
        tl(j, 4);
        bool never;
        if (POLY8 > 0 && !masked) never = a4_exp_inplace<POLY8>(v0, v1, v2, v3, c2, nm2);
        else never = a4_exp_inplace<0>(v0, v1, v2, v3, c2, nm2);
        // Basic-block boundary that ptxas neither schedules across nor sinks a MUFU below: a never-taken exit on which all 128
        // results are live.  Without it ptxas interleaves each pair's row-sum / pack right behind its two MUFUs; the consumer
        // then waits ~48 clk for a result the XU delivers 8 clk apart, and the warp's MUFU rate halves (tools/probes/
        // exp_sched_probe.cu, tools/sass_sched.py).
        if (never) {
          a4_dump(a.out, v0, v1, v2, v3);
          asm volatile("trap;");
        }
        tl(j, 5);
        if (pingpong && !(X == 1 && j == n_kv - 1)) a2_token_pass(2 - X);   // the other warpgroup's exponentials may start
        // ---- C phase, first half of the row: row sum + bf16 pack
        uint32_t pk0[16], pk1[16];
        a4_sum_pack(v0, pk0, l01, l23);
        a4_sum_pack(v1, pk1, l01, l23);
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
        tl(j, 6);
        tmem_st16(t_p, pk0);
        tmem_st16(t_p + 16, pk1);
        tl(j, 7);
        // ---- second half
        a4_sum_pack(v2, pk0, l01, l23);
        a4_sum_pack(v3, pk1, l01, l23);
        tmem_st16(t_p + 32, pk0);
        tmem_st16(t_p + 48, pk1);
        tmem_st_wait();
        tc_fence_before();
        tl(j, 8);
        mbar_arrive(&bar_p_full[X]);
        tl(j, 9);
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


template <int POLY8, int PINGPONG, int TL = 0>
static int attn4_launch_t(const CUtensorMap* tm, const Attn2Args& a, dim3 grid, cudaStream_t stream) {
  auto kern = attn4_fwd_kernel<POLY8, PINGPONG, TL>;
  if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(kern), A2_SMEM_BYTES, "attn4_fwd_kernel")) return rc;
  kern<<<grid, A2_THREADS, A2_SMEM_BYTES, stream>>>(tm[0], tm[1], tm[2], a);
  return check_launch("pf_attn_fwd_masked(pair kernel, E/C phases)");
}

int warmup_attn4() {
  int rc = 0;
#define PF_WARM4(P, Q) if (!rc) rc = ensure_dyn_smem(reinterpret_cast<const void*>(attn4_fwd_kernel<P, Q, 0>), A2_SMEM_BYTES, "attn4_fwd_kernel")
  PF_WARM4(0, 0); PF_WARM4(1, 0); PF_WARM4(2, 0); PF_WARM4(3, 0); PF_WARM4(4, 0);
  PF_WARM4(0, 1); PF_WARM4(1, 1); PF_WARM4(2, 1); PF_WARM4(3, 1); PF_WARM4(4, 1);
#undef PF_WARM4
  return rc;
}

// poly8 = exponential pairs per 8 on the FMA pipe (0..4)
int attn4_launch_raw(const CUtensorMap* tm, const Attn2Args& a, dim3 grid, int poly8, int pingpong, cudaStream_t stream) {
  if (a.timeline != nullptr) {   // debug: the instrumented instantiations (poly8 0 or 3 only)
    if (poly8 == 0) return pingpong ? attn4_launch_t<0, 1, 1>(tm, a, grid, stream) : attn4_launch_t<0, 0, 1>(tm, a, grid, stream);
    return pingpong ? attn4_launch_t<3, 1, 1>(tm, a, grid, stream) : attn4_launch_t<3, 0, 1>(tm, a, grid, stream);
  }
  switch (poly8 * 2 + (pingpong ? 1 : 0)) {
    case 0: return attn4_launch_t<0, 0>(tm, a, grid, stream);
    case 1: return attn4_launch_t<0, 1>(tm, a, grid, stream);
    case 2: return attn4_launch_t<1, 0>(tm, a, grid, stream);
    case 3: return attn4_launch_t<1, 1>(tm, a, grid, stream);
    case 4: return attn4_launch_t<2, 0>(tm, a, grid, stream);
    case 5: return attn4_launch_t<2, 1>(tm, a, grid, stream);
    case 6: return attn4_launch_t<3, 0>(tm, a, grid, stream);
    case 7: return attn4_launch_t<3, 1>(tm, a, grid, stream);
    case 8: return attn4_launch_t<4, 0>(tm, a, grid, stream);
    case 9: return attn4_launch_t<4, 1>(tm, a, grid, stream);
  }
  set_error("pf_attn_fwd_masked: bad poly8 %d", poly8);
  return -1;
}

}  // namespace pf
