// pf_attn5.cu -- masked joint attention forward, two q tiles per CTA, one thread per score row, software-pipelined softmax.
//
// Same shell and contract as pf_attn2.cu (one CTA per SM owns two adjacent 128-row q tiles of one (batch, head) and walks the
// union of their kv tile lists once; S = Q.K^T as SS MMAs, O += P.V as TS MMAs with P in TMEM; host-built pair schedule and row
// masks; exact thread-local row max, lazy O rescale).  What changes is WHEN a softmax thread talks to TMEM, after a per-phase
// clock64 timeline of pf_attn2/4 on B200 (profiles/r02_attn4_timeline.txt, tools/probes/exp_sched_probe.cu):
//
//   * one warp per SMSP runs the exponential stream at 8.96 clk per MUFU.EX2 and two warps at 8.46 (95 % of the XU's rate):
//     the stream itself was never the limit (an earlier probe that said "one MUFU per 16 clk per warp" was spilling);
//   * per kv tile each softmax warp spent ~2150 clk exponentiating (XU shared with the other q tile's warp) and ~750-900 clk
//     with the XU IDLE: wait for S, 4 x tcgen05.ld (220-440 clk), row max (64 FMNMX3), P stores, barrier round trips --
//     and both warps of an SMSP do this at the same moment (they fall into lockstep), so the XU pipe sat at 65-70 %;
//   * strict alternation of the two q tiles (ping-pong token) does not help: a warp's non-exponential phases run 2x slower
//     while its SMSP neighbour streams MUFUs (every tcgen05.ld / st / mbarrier operation shares the MIO path with them).
//
// Here the row is processed in four 32-column chunks, and the registers a chunk frees are refilled at once with the same
// columns of S(j+1) (tcgen05.ld is asynchronous and completes under the next chunk's MUFUs); the max of the next tile is
// folded in chunk by chunk, and P(j) goes to TMEM chunk by chunk.  Between the last MUFU of tile j and the first of tile j+1
// remain one tcgen05.ld latency, 16 FMNMX3 and the lazy-rescale test.
#include "pf_attn_pair.cuh"

namespace pf {

template <int TL>
__global__ void __launch_bounds__(A2_THREADS, 1)
attn5_fwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
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
      const int* mask_idx = a.pmask_idx + (static_cast<size_t>(b) * a.n_pairs + pair) * 2 * a.sched_stride;
      const bool tl_on = TL && a.timeline != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && quarter == 0 && lane == 0;
      auto tl = [&](int j, int slot) {
        if (TL && tl_on && j < 64) a.timeline[(X * 64 + j) * 12 + slot] = clock64();
      };
      // element mask of a partial tile: 128 allow bits per q row, precomputed on the host (pf_attn_build_pair_masks)
      auto load_allow = [&](int jj, int ent, uint4& w) -> bool {   // returns "tile jj needs the element mask"
        const int fl = (ent >> (2 * X)) & 3;                       // bit0: this q tile has allowed pairs here, bit1: partial
        w = make_uint4(0u, 0u, 0u, 0u);
        if (fl == 3) {
          const int blk = __ldg(mask_idx + 2 * jj + X);
          w = __ldg(a.pmask_bits + static_cast<size_t>(blk) * A2_BM + row);
        }
        return fl != 1;                                            // not owned (all -inf) or partial
      };

      // ---- prologue: S(0) -> registers, its row max
      uint32_t v0[32], v1[32], v2[32], v3[32];
      float m_tile;
      int e1 = n_kv > 1 ? __ldg(sched + 2) : 0;   // schedule entry of tile j + 1
      {
        uint4 w;
        const bool msk = load_allow(0, sched[1], w);
        mbar_wait(&bar_s_full[X], 0);
        tc_fence_after();
        if (a.trace && threadIdx.x == 128) cta_stamp[1] = clock64();
        if (X == 1 && act_lo && a.b_delay > 0) {   // de-phase the two q tiles once per CTA (see pf_attn2.cu)
          const long long t_begin = clock64();
          while (clock64() - t_begin < a.b_delay) {
          }
        }
        tmem_ld32(t_s, v0);
        tmem_ld32(t_s + 32, v1);
        tmem_ld32(t_s + 64, v2);
        tmem_ld32(t_s + 96, v3);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&bar_s_free[X]);
        if (msk) {
          a2_mask32(v0, w.x);
          a2_mask32(v1, w.y);
          a2_mask32(v2, w.z);
          a2_mask32(v3, w.w);
        }
        m_tile = fmaxf(fmaxf(a2_max32(v0), a2_max32(v1)), fmaxf(a2_max32(v2), a2_max32(v3)));
      }

      // One kv tile: v0..v3 hold the (masked) scores of tile j and m_tile their row max.  The exponentials run chunk by chunk
      // (32 columns: scale/subtract, MUFU.EX2, row sum, bf16 pack, P chunk -> TMEM); as soon as a chunk's registers are free the
      // same columns of S(j+1) are loaded into them (tcgen05.ld is asynchronous: it completes under the next chunk's MUFU
      // stream), and the max of the chunk loaded one step earlier is folded in.  After the last chunk only one TMEM load and
      // 16 FMNMX3 stand between this tile's exponentials and the next tile's: the loads, the max and the P stores that cost
      // pf_attn2 ~850 clk per kv tile with the XU idle (profiles/r02_attn4_timeline.txt) now run under the MUFU stream.
#pragma unroll 1
      for (int j = 0; j < n_kv; ++j) {
        tl(j, 0);
        const bool has_next = j + 1 < n_kv;
        const int e2 = (j + 2 < n_kv) ? __ldg(sched + 3 + j) : 0;
        uint4 w = make_uint4(0u, 0u, 0u, 0u);
        bool msk_n = false;
        if (has_next) msk_n = load_allow(j + 1, e1, w);
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
        uint32_t pka[16], pkb[16];
        // ---- chunk 0
        a2_exp32<0>(v0, pka, c2, nm2, l01, l23);
        tl(j, 1);
        // Both barriers complete about now (P.V(j-1) ~450 clk after the end of tile j-1, S(j+1) ~450 clk after S(j) was
        // released): probe them here, consume the predicates one chunk later -- an mbarrier round trip through the MIO queue
        // costs 100-300 clk behind the MUFU stream, and stalled here both q tiles' warps would leave the XU idle together.
        // (unconditional: for j = 0 the parity-1 probe of the fresh barrier is true at once, and the probe of S(n_kv) is unused)
        const bool pv_ok = mbar_test(&bar_pv_done[X], (j + 1) & 1);
        const bool sfull_ok = mbar_test(&bar_s_full[X], (j + 1) & 1);
        // ---- chunk 1
        a2_exp32<0>(v1, pkb, c2, nm2, l01, l23);
        tl(j, 2);
        if (j > 0) {                             // P(j-1) consumed and O(j-1) produced before P is overwritten / O is rescaled
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
        tmem_st16(t_p, pka);
        tmem_st16(t_p + 16, pkb);
        if (has_next) {
          if (!sfull_ok) mbar_wait(&bar_s_full[X], (j + 1) & 1);
          tc_fence_after();
          tmem_ld32(t_s, v0);                    // S(j+1) columns 0..63 into the registers chunks 0 and 1 just freed
          tmem_ld32(t_s + 32, v1);
        }
        tl(j, 3);
        // ---- chunk 2
        a2_exp32<0>(v2, pka, c2, nm2, l01, l23);
        tmem_st16(t_p + 32, pka);
        float mx = -INFINITY;
        if (has_next) {
          tmem_ld_wait();                        // issued one chunk (>= 256 MUFU clocks) ago
          if (msk_n) {
            a2_mask32(v0, w.x);
            a2_mask32(v1, w.y);
          }
          mx = fmaxf(a2_max32(v0), a2_max32(v1));
          tmem_ld32(t_s + 64, v2);
        }
        tl(j, 4);
        // ---- chunk 3
        a2_exp32<0>(v3, pkb, c2, nm2, l01, l23);
        tmem_st16(t_p + 48, pkb);
        if (has_next) {
          tmem_ld_wait();
          if (msk_n) a2_mask32(v2, w.z);
          mx = fmaxf(mx, a2_max32(v2));
          tmem_ld32(t_s + 96, v3);
        }
        tl(j, 5);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&bar_p_full[X]);
        tl(j, 6);
        if (has_next) {
          tmem_ld_wait();
          tc_fence_before();
          mbar_arrive(&bar_s_free[X]);           // S(j+1) lives in registers: the tensor pipe may write S(j+2)
          tl(j, 7);
          if (msk_n) a2_mask32(v3, w.w);
          m_tile = fmaxf(mx, a2_max32(v3));
          e1 = e2;
        }
        tl(j, 8);
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


int warmup_attn5() {
  int rc = ensure_dyn_smem(reinterpret_cast<const void*>(attn5_fwd_kernel<0>), A2_SMEM_BYTES, "attn5_fwd_kernel<0>");
  if (!rc) rc = ensure_dyn_smem(reinterpret_cast<const void*>(attn5_fwd_kernel<1>), A2_SMEM_BYTES, "attn5_fwd_kernel<1>");
  return rc;
}

int attn5_launch_raw(const CUtensorMap* tm, const Attn2Args& a, dim3 grid, cudaStream_t stream) {
  if (int rc = warmup_attn5()) return rc;
  if (a.timeline != nullptr) attn5_fwd_kernel<1><<<grid, A2_THREADS, A2_SMEM_BYTES, stream>>>(tm[0], tm[1], tm[2], a);   // debug
  else attn5_fwd_kernel<0><<<grid, A2_THREADS, A2_SMEM_BYTES, stream>>>(tm[0], tm[1], tm[2], a);
  return check_launch("pf_attn_fwd_masked(pair kernel, pipelined softmax)");
}

}  // namespace pf
