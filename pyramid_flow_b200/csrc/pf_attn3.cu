// pf_attn3.cu — masked joint attention forward, two q tiles per CTA, TWO threads per score row (head_dim 64).
//
// Same shell as pf_attn2.cu (one CTA per SM owns two adjacent 128-row q tiles and walks the union of their kv lists once;
// S = Q.K^T as SS MMAs, O += P.V as TS MMAs with P in TMEM; host-built pair schedule and row masks), different softmax
// mapping, chosen from measurements of pf_attn2 on B200 (profiles/r02_attn_cta_phases.txt, tools/probes/tmem_probe.cu):
//   * one warp's MUFU.EX2 stream runs at one instruction per 16 clk however its consumers are scheduled (7.9 ex2/clk/SM with
//     one exponentiating warp per SMSP, 11.3 with two, 13.7 with four; the XU pipe's rate is 16/clk/SM);
//   * with one thread per full row a warp spends 128 x 16 = 2048 clk per kv tile in MUFU issue alone, and only two softmax
//     warps fit on an SMSP (232 registers each): measured 2906 clk per pair of tiles, XU pipe 70 % busy.
// Here a row's 128 scores are split over two threads (64 + 64 columns, 104 registers), so 16 softmax warps are resident -- four
// per SMSP -- and each spends 1024 clk per kv tile in MUFU issue.  The row max stays EXACT: the two halves publish their
// partial max in shared memory and meet on a 64-thread named barrier per tile (the two warps sit on the same SMSP and load S
// at the same moment, so the barrier costs one MIO round trip, not a wait).
//   warps 0-15  softmax: warp w -> q tile X = w >> 3, column half = (w >> 2) & 1, TMEM lane quarter = w & 3
//   warp 16/17  MMA issuers of tile A/B;  warp 18  TMA producer;  warp 19  TMEM allocation
// TMEM (512 columns): tile X at column 256 X: S fp32 [0,128) | O fp32 [128,192) | P bf16x2 [192,256).
#include <algorithm>

#include "pf_attn_pair.cuh"

namespace pf {

constexpr int A3_THREADS = 640;
constexpr int A3_REGS_SOFTMAX = 104, A3_REGS_OTHER = 40;   // 512*104 + 128*40 = 58368 <= 640*96 (launch allocation)

__device__ __forceinline__ void a3_pair_sync(int id) {
  // 64-thread named barriers 1..8: (q tile, lane quarter) -> the two warps holding the two halves of the same 32 rows
  switch (id) {
    case 0: asm volatile("bar.sync 1, 64;" ::: "memory"); break;
    case 1: asm volatile("bar.sync 2, 64;" ::: "memory"); break;
    case 2: asm volatile("bar.sync 3, 64;" ::: "memory"); break;
    case 3: asm volatile("bar.sync 4, 64;" ::: "memory"); break;
    case 4: asm volatile("bar.sync 5, 64;" ::: "memory"); break;
    case 5: asm volatile("bar.sync 6, 64;" ::: "memory"); break;
    case 6: asm volatile("bar.sync 7, 64;" ::: "memory"); break;
    default: asm volatile("bar.sync 8, 64;" ::: "memory"); break;
  }
}

// named barriers 9 / 10 = the XU token of q tile A / B (512 = the 256 waiting threads + the 256 passing ones)
__device__ __forceinline__ void a3_token_wait(int x) {
  if (x == 0) asm volatile("bar.sync 9, 512;" ::: "memory");
  else asm volatile("bar.sync 10, 512;" ::: "memory");
}
__device__ __forceinline__ void a3_token_pass(int x) {
  if (x == 0) asm volatile("bar.arrive 9, 512;" ::: "memory");
  else asm volatile("bar.arrive 10, 512;" ::: "memory");
}

template <int PINGPONG>
__global__ void __launch_bounds__(A3_THREADS, 1)
attn3_fwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                 const __grid_constant__ CUtensorMap tm_v, const Attn2Args a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_q = smem;                                   // 2 tiles
  uint8_t* smem_k = smem + 2 * A2_TILE_BYTES;
  uint8_t* smem_v = smem_k + A2_KSTAGES * A2_TILE_BYTES;

  __shared__ __align__(8) uint64_t bar_q[2], bar_s_full[2], bar_s_free[2], bar_p_full[2], bar_pv_done[2];
  __shared__ __align__(8) uint64_t k_full[A2_KSTAGES], k_empty[A2_KSTAGES], v_full[A2_VSTAGES], v_empty[A2_VSTAGES];
  __shared__ uint32_t tmem_slot;
  __shared__ float xch[2][2][2][A2_BM];   // [q tile][tile parity][column half][row]: partial row max / (epilogue) row sum

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int pair = blockIdx.x;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int bh = b * a.heads + h;
  const int qt_hi = a.q_tiles - 1 - 2 * pair;
  const int qt_lo = qt_hi - 1;
  const bool act_lo = qt_lo >= a.q_tile_begin;
  const int n_act = act_lo ? 2 : 1;
  const int* sched = a.psched + (static_cast<size_t>(b) * a.n_pairs + pair) * a.sched_stride;
  const int n_kv = sched[0];

  if (warp == 18 && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
  }
  if (warp == 16 && lane == 0) {
    for (int x = 0; x < 2; ++x) {
      mbar_init(&bar_q[x], 1);
      mbar_init(&bar_s_full[x], 1);
      mbar_init(&bar_s_free[x], 256);
      mbar_init(&bar_p_full[x], 256);
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
  if (warp == 19) {
    tmem_alloc(&tmem_slot, A2_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp >= 16) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(A3_REGS_OTHER));
    if (warp == 18) {
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
    } else if (warp == 16 || warp == 17) {
      const int X = warp - 16;
      if ((X == 1 || act_lo) && elect_one()) {
        // ===== MMA issuer of q tile X =====
        constexpr uint32_t idesc_qk = make_idesc_bf16(A2_BM, A2_BN, 0, 0);
        constexpr uint32_t idesc_pv = make_idesc_bf16(A2_BM, A2_HD, 0, 1);
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
        issue_qk();
        for (int j = 0; j < n_kv; ++j) {
          if (j + 1 < n_kv) {
            mbar_wait(&bar_s_free[X], j & 1);
            issue_qk();
          }
          mbar_wait(&bar_p_full[X], j & 1);
          mbar_wait(&v_full[vs], vph);
          tc_fence_after();
          const uint32_t sv = smem_u32(smem_v + vs * A2_TILE_BYTES);
#pragma unroll
          for (int kk = 0; kk < A2_BN / 16; ++kk) {
            const uint64_t dv = make_smem_desc(sv + kk * 2048, A2_BN * 128, 1024);
            umma_ts(t_o, t_p + kk * 8, dv, idesc_pv, (j | kk) != 0);
          }
          umma_commit(&v_empty[vs]);
          umma_commit(&bar_pv_done[X]);
          if (++vs == A2_VSTAGES) { vs = 0; vph ^= 1; }
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(A3_REGS_SOFTMAX));
    // ===== softmax: thread = (q tile X, row, column half) =====
    const int X = warp >> 3;
    const int half = (warp >> 2) & 1;
    const int quarter = warp & 3;
    if (X == 1 || act_lo) {
      const int qt = X ? qt_hi : qt_lo;
      const int row = quarter * 32 + lane;
      const int qpos = qt * A2_BM + row;
      const bool q_valid = qpos < a.seq;
      const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
      const uint32_t t_s = tmem_base + lane_base + X * A2_TM_TILE + A2_TM_S + half * 64;
      const uint32_t t_o = tmem_base + lane_base + X * A2_TM_TILE + A2_TM_O + half * 32;
      const uint32_t t_p = tmem_base + lane_base + X * A2_TM_TILE + A2_TM_P + half * 32;
      const int bar_id = X * 4 + quarter;
      const float c = a.scale_log2;
      const uint64_t c2 = f2_pack(c, c);
      float m_run = -INFINITY;
      uint64_t l01 = f2_pack(0.f, 0.f), l23 = f2_pack(0.f, 0.f);
      int entry = sched[1];
      const int* mask_idx = a.pmask_idx + (static_cast<size_t>(b) * a.n_pairs + pair) * 2 * a.sched_stride;
      // Ping-pong between the two q tiles: left alone, all 16 softmax warps run in lockstep (both S tiles become ready
      // together): 4 warps per SMSP share the XU during their exponentials (2048 clk) and leave it idle while all of them
      // load / reduce / exchange / store (~850 clk) -- measured 2950 clk per pair of tiles, the same as pf_attn2.  With the
      // token, tile A's two warps per SMSP exponentiate (2 x 64 MUFU.EX2 at the per-warp rate of one per 16 clk = the XU's
      // full rate) while tile B's warps do everything else, and vice versa.
      const bool pingpong = PINGPONG && act_lo;
      if (pingpong && X == 1) a3_token_pass(0);

      for (int j = 0; j < n_kv; ++j) {
        const int fl = (entry >> (2 * X)) & 3;
        const bool own = (fl & 1) != 0;
        const bool masked = !own || (fl & 2) != 0;
        if (j + 1 < n_kv) entry = __ldg(sched + 2 + j);
        uint32_t allow0 = 0u, allow1 = 0u;
        if (own && masked) {
          const int blk = __ldg(mask_idx + 2 * j + X);
          const uint2 w = __ldg(reinterpret_cast<const uint2*>(a.pmask_bits + static_cast<size_t>(blk) * A2_BM + row) + half);
          allow0 = w.x;
          allow1 = w.y;
        }
        bool pv_ok = true;
        if (j > 0) pv_ok = mbar_test(&bar_pv_done[X], (j - 1) & 1);
        mbar_wait(&bar_s_full[X], j & 1);
        tc_fence_after();

        uint32_t va[32], vb[32];
        tmem_ld32(t_s, va);
        tmem_ld32(t_s + 32, vb);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&bar_s_free[X]);
        if (masked) {
          a2_mask32(va, allow0);
          a2_mask32(vb, allow1);
        }
        // ---- exact row max: own 64 columns, then the partner's partial through shared memory
        const float m_part = fmaxf(a2_max32(va), a2_max32(vb));
        xch[X][j & 1][half][row] = m_part;
        a3_pair_sync(bar_id);
        const float m_tile = fmaxf(m_part, xch[X][j & 1][half ^ 1][row]);

        float alpha = 1.f;
        bool need = false;
        if (m_tile > m_run) {
          if (m_run == -INFINITY) {
            m_run = m_tile;
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
        uint32_t pk0[16], pk1[16];
        if (pingpong) a3_token_wait(X);
        a2_exp32<0>(va, pk0, c2, nm2, l01, l23);
        a2_exp32<0>(vb, pk1, c2, nm2, l01, l23);
        if (pingpong && !(X == 1 && j == n_kv - 1)) a3_token_pass(X ^ 1);
        if (j > 0) {
          if (!pv_ok) mbar_wait(&bar_pv_done[X], (j - 1) & 1);
          tc_fence_after();
          if (__any_sync(0xffffffffu, need)) {
#pragma unroll 1
            for (int cc = 0; cc < 32; cc += 16) {
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
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&bar_p_full[X]);
      }

      // ---- epilogue: combine the halves' row sums, O / l -> bf16 -> out[b, qpos, h*64 + half*32 .. +32]
      float s0, s1, s2, s3;
      f2_unpack(l01, s0, s1);
      f2_unpack(l23, s2, s3);
      const float l_part = (s0 + s1) + (s2 + s3);
      xch[X][n_kv & 1][half][row] = l_part;      // parity n_kv & 1: last read by the partner two tiles ago
      a3_pair_sync(bar_id);
      const float l_run = l_part + xch[X][n_kv & 1][half ^ 1][row];
      mbar_wait(&bar_pv_done[X], (n_kv - 1) & 1);
      tc_fence_after();
      const float inv = (l_run > 0.f) ? 1.f / l_run : 0.f;
      __nv_bfloat16* dst;
      if (a.peer_count > 1) {
        const int r = min(qpos / a.peer_chunk_rows, a.peer_count - 1);
        dst = a.peer_out[r] + static_cast<size_t>(qpos - r * a.peer_chunk_rows) * a.ldo + a.peer_col_begin + h * A2_HD + half * 32;
      } else {
        dst = a.out + (static_cast<size_t>(b) * a.seq + qpos) * a.ldo + h * A2_HD + half * 32;
      }
      uint32_t o[32];
      tmem_ld32(t_o, o);
      tmem_ld_wait();
      if (q_valid) {
        uint4* d4 = reinterpret_cast<uint4*>(dst);
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

  tc_fence_before();
  __syncthreads();
  if (warp == 19) {
    tc_fence_after();
    tmem_dealloc(tmem_base, A2_TMEM_COLS);
  }
}

int warmup_attn3() {
  int rc = ensure_dyn_smem(reinterpret_cast<const void*>(attn3_fwd_kernel<0>), A2_SMEM_BYTES, "attn3_fwd_kernel<0>");
  if (!rc) rc = ensure_dyn_smem(reinterpret_cast<const void*>(attn3_fwd_kernel<1>), A2_SMEM_BYTES, "attn3_fwd_kernel<1>");
  return rc;
}

int attn3_launch_raw(const CUtensorMap* tm, const Attn2Args& a, dim3 grid, int pingpong, cudaStream_t stream) {
  if (int rc = warmup_attn3()) return rc;
  if (pingpong) attn3_fwd_kernel<1><<<grid, A3_THREADS, A2_SMEM_BYTES, stream>>>(tm[0], tm[1], tm[2], a);
  else attn3_fwd_kernel<0><<<grid, A3_THREADS, A2_SMEM_BYTES, stream>>>(tm[0], tm[1], tm[2], a);
  return check_launch("pf_attn_fwd_masked(pair kernel, split rows)");
}

}  // namespace pf
