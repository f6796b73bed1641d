// pf_attn3q.cu — masked joint attention forward, THREE q tiles per CTA and 64-column kv steps (head_dim 64): the default kernel
// of launches without peer stores (pf_attn_desc.variant 0x20, or variant 0 under PF_OPT_ATTN_TRIPLE_KERNEL).  Same contract and
// the same math as pf_attn2.cu; 2.80 -> 2.60 ms per launch at the bench shape (profiles/r02_attn3q_validation.txt).
//
// Why: the timeline of pf_attn2 shows each softmax warp spending ~1450 clk per kv tile NOT feeding the XU
// (barrier round trip, TMEM loads, row max, P stores) and only two softmax warps per SMSP to cover for each other -- the XU
// pipe, which bounds attention at head_dim 64, stays at 68 %.  A 128-row q tile needs one softmax warp per SMSP (TMEM lane
// rule), so a third warp per SMSP means a third q tile per CTA, and TMEM (512 columns) holds three tiles only with 64-column
// kv steps: per tile S fp32 64 | O fp32 64 | P bf16x2 32 = 160 columns.
//
//   * one CTA per SM owns three adjacent 128-row q tiles of one (batch, head) and walks the union of their kv tile lists once
//     (host-built group schedule, pf_attn_build_group_schedule); every 128-row K/V tile is loaded once (TMA rings as in
//     pf_attn2) and consumed in two half steps of 64 kv rows;
//   * 12 softmax warps: warpgroup X (warps 4X..4X+3) owns q tile X, one thread = one row, 64 scores per half step in registers,
//     exact thread-local row max, lazy O rescale (branch-free test), every exponential a MUFU.EX2;
//   * warps 12/13/14 issue the MMAs of tile 0/1/2: S = Q.K_half^T (4 x M128 N64 K16, SS), O += P.V_half (4 x M128 N64 K16, TS);
//     warp 15 = TMA producer, also owns the TMEM allocation; 64 scores per thread need no setmaxnreg;
//   * tile X's softmax warps start 0.7 X b_delay clocks late, once per CTA (PF_OPT_ATTN_TILE_PHASE) -- measured: no effect here,
//     three warps per SMSP drift apart by themselves.
#include <algorithm>

#include "pf_attn_pair.cuh"

namespace pf {

constexpr int A3_G = 3;                                   // q tiles per CTA
constexpr int A3_THREADS = 512;                           // 12 softmax warps + 3 MMA issuers + TMA (also owns the TMEM allocation)
constexpr int A3_BH = 64;                                 // kv rows per half step
constexpr int A3_SMEM_BYTES = (A3_G + A2_KSTAGES + A2_VSTAGES) * A2_TILE_BYTES + 1024;
constexpr uint32_t A3_TM_TILE = 160, A3_TM_S = 0, A3_TM_O = 64, A3_TM_P = 128;
// (no setmaxnreg: a row of a half step is 64 scores, the softmax threads fit the 128 registers 512 threads can have)

struct Attn3Args {
  Attn2Args c;                 // common fields (psched / pmask_* hold the GROUP schedule and masks here)
  int n_groups;
};

__global__ void __launch_bounds__(A3_THREADS, 1)
attn3q_fwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                  const __grid_constant__ CUtensorMap tm_v, const Attn3Args aa) {
  const Attn2Args& a = aa.c;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_q = smem;                                   // 3 tiles
  uint8_t* smem_k = smem + A3_G * A2_TILE_BYTES;
  uint8_t* smem_v = smem_k + A2_KSTAGES * A2_TILE_BYTES;

  __shared__ __align__(8) uint64_t bar_q[A3_G], bar_s_full[A3_G], bar_s_free[A3_G], bar_p_full[A3_G], bar_pv_done[A3_G];
  __shared__ __align__(8) uint64_t k_full[A2_KSTAGES], k_empty[A2_KSTAGES], v_full[A2_VSTAGES], v_empty[A2_VSTAGES];
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int grp = blockIdx.x;                               // group 0 = the last three q tiles (longest kv lists first)
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int bh = b * a.heads + h;
  const int qt_top = a.q_tiles - 1 - A3_G * grp;            // tile X = 2; tile X covers q tile qt_top - (2 - X)
  // active tiles are a suffix: X >= x_first (a tile is missing below the sequence start or below q_tile_begin)
  const int x_first = max(0, (A3_G - 1) - (qt_top - a.q_tile_begin));
  const int n_act = A3_G - x_first;
  const int* sched = a.psched + (static_cast<size_t>(b) * aa.n_groups + grp) * a.sched_stride;
  const int n_kv = sched[0];
  const int n_steps = 2 * n_kv;

  if (warp == 15 && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
  }
  if (warp == 12 && lane == 0) {
    for (int x = 0; x < A3_G; ++x) {
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
  if (warp == 15) {
    tmem_alloc(&tmem_slot, A2_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp >= 12) {
    if (warp == 15) {
      if (elect_one()) {
        // ===== TMA producer =====
        for (int x = x_first; x < A3_G; ++x) {
          mbar_arrive_expect_tx(&bar_q[x], A2_TILE_BYTES);
          tma_load_3d(smem_q + x * A2_TILE_BYTES, &tm_q, &bar_q[x], 0, (qt_top - (A3_G - 1 - x)) * A2_BM, bh);
        }
        int ks = 0, vs = 0;
        uint32_t kph = 0, vph = 0;
        for (int j = 0; j < n_kv; ++j) {
          const int kt = sched[1 + j] >> 8;
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
    } else if (warp < 15) {
      const int X = warp - 12;
      if (X >= x_first && elect_one()) {
        // ===== MMA issuer of q tile X: every kv tile in two half steps (t = 2 j + half); all issuers walk the same list and a
        // K / V stage is released when every active issuer has committed its second half =====
        constexpr uint32_t idesc_qk = make_idesc_bf16(A2_BM, A3_BH, 0, 0);  // A = Q (K-major), B = K half (K-major), N = 64
        constexpr uint32_t idesc_pv = make_idesc_bf16(A2_BM, A2_HD, 0, 1);  // A = P (TMEM),    B = V half (MN-major), N = 64
        const uint32_t t_s = tmem_base + X * A3_TM_TILE + A3_TM_S;
        const uint32_t t_o = tmem_base + X * A3_TM_TILE + A3_TM_O;
        const uint32_t t_p = tmem_base + X * A3_TM_TILE + A3_TM_P;
        mbar_wait(&bar_q[X], 0);
        const uint64_t dq = make_smem_desc_kmajor_sw128(smem_u32(smem_q + X * A2_TILE_BYTES));
        int ks = 0, vs = 0;
        uint32_t kph = 0, vph = 0;
        auto issue_qk = [&](int t) {
          const int half = t & 1;
          if (half == 0) {
            mbar_wait(&k_full[ks], kph);
            tc_fence_after();
          }
          // K rows 64 half .. 64 half + 63 of the stage: 64 x 128 B further on (8192 B keeps the 128-byte swizzle phase)
          const uint64_t dk = make_smem_desc_kmajor_sw128(smem_u32(smem_k + ks * A2_TILE_BYTES + half * (A3_BH * 128)));
#pragma unroll
          for (int kk = 0; kk < A2_HD / 16; ++kk) umma_ss(t_s, dq + 2 * kk, dk + 2 * kk, idesc_qk, kk != 0);
          if (half == 1) {
            umma_commit(&k_empty[ks]);
            if (++ks == A2_KSTAGES) { ks = 0; kph ^= 1; }
          }
          umma_commit(&bar_s_full[X]);
        };
        issue_qk(0);
        for (int t = 0; t < n_steps; ++t) {
          if (t + 1 < n_steps) {
            mbar_wait(&bar_s_free[X], t & 1);   // S(t) lives in the softmax threads' registers
            issue_qk(t + 1);                    // S(t+1) runs on the tensor pipe under softmax(t)
          }
          mbar_wait(&bar_p_full[X], t & 1);
          const int half = t & 1;
          if (half == 0) mbar_wait(&v_full[vs], vph);
          tc_fence_after();
          // V tile [128 kv x 64 hd], 128-byte rows: MN-major, 16 kv rows (2048 B) per MMA; this half = kv rows 64 half ..
          const uint32_t sv = smem_u32(smem_v + vs * A2_TILE_BYTES) + half * (A3_BH * 128);
#pragma unroll
          for (int kk = 0; kk < A3_BH / 16; ++kk) {
            const uint64_t dv = make_smem_desc(sv + kk * 2048, A2_BN * 128, 1024);
            umma_ts(t_o, t_p + kk * 8, dv, idesc_pv, (t | kk) != 0);
          }
          if (half == 1) {
            umma_commit(&v_empty[vs]);
            if (++vs == A2_VSTAGES) { vs = 0; vph ^= 1; }
          }
          umma_commit(&bar_pv_done[X]);
        }
      }
    }
  } else {
    // ===== softmax + lazy O rescale + epilogue: warpgroup X owns q tile X, thread = one row, 64 columns per half step =====
    const int X = warp >> 2;
    const int quarter = warp & 3;
    if (X >= x_first) {
      const int qt = qt_top - (A3_G - 1 - X);
      const int row = quarter * 32 + lane;
      const int qpos = qt * A2_BM + row;
      const bool q_valid = qpos < a.seq;
      const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
      const uint32_t t_s = tmem_base + lane_base + X * A3_TM_TILE + A3_TM_S;
      const uint32_t t_o = tmem_base + lane_base + X * A3_TM_TILE + A3_TM_O;
      const uint32_t t_p = tmem_base + lane_base + X * A3_TM_TILE + A3_TM_P;
      const float c = a.scale_log2;
      const uint64_t c2 = f2_pack(c, c);
      float m_run = -INFINITY;   // reference max (raw score units) the accumulators are scaled by; -inf: nothing finite yet
      uint64_t l01 = f2_pack(0.f, 0.f), l23 = f2_pack(0.f, 0.f);
      int entry = sched[1];
      const int* mask_idx = a.pmask_idx + (static_cast<size_t>(b) * aa.n_groups + grp) * A3_G * a.sched_stride;
      if (X > x_first && a.b_delay > 0) {     // de-phase the q tiles once per CTA (see pf_attn2.cu)
        mbar_wait(&bar_s_full[X], 0);
        const long long t_begin = clock64();
        // three tiles share a half step of ~1700 clk: thirds of it, where pf_attn2's two tiles are b_delay apart
        const long long t_wait = static_cast<long long>(a.b_delay) * 7 / 10 * (X - x_first);
        while (clock64() - t_begin < t_wait) {
        }
      }
      uint32_t allow0 = 0u, allow1 = 0u, allow2 = 0u, allow3 = 0u;
      bool masked = false;

      for (int t = 0; t < n_steps; ++t) {
        const int j = t >> 1;
        const int half = t & 1;
        if (half == 0) {
          const int fl = (entry >> (2 * X)) & 3;            // bit0: this tile has allowed pairs here, bit1: element mask
          const bool own = (fl & 1) != 0;
          masked = !own || (fl & 2) != 0;
          if (j + 1 < n_kv) entry = __ldg(sched + 2 + j);
          allow0 = allow1 = allow2 = allow3 = 0u;
          if (own && masked) {
            const int blk = __ldg(mask_idx + A3_G * j + X);
            const uint4 w = __ldg(a.pmask_bits + static_cast<size_t>(blk) * A2_BM + row);
            allow0 = w.x;
            allow1 = w.y;
            allow2 = w.z;
            allow3 = w.w;
          }
        }
        bool pv_ok = true;
        if (t > 0) pv_ok = mbar_test(&bar_pv_done[X], (t - 1) & 1);    // probed early, consumed before the P store
        mbar_wait(&bar_s_full[X], t & 1);
        tc_fence_after();

        // ---- the row's 64 scores of this half step: TMEM -> registers, then the tensor pipe may overwrite S
        uint32_t v0[32], v1[32];
        tmem_ld32(t_s, v0);
        tmem_ld32(t_s + 32, v1);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&bar_s_free[X]);
        if (masked) {
          a2_mask32(v0, half ? allow2 : allow0);
          a2_mask32(v1, half ? allow3 : allow1);
        }
        const float m_tile = fmaxf(a2_max32(v0), a2_max32(v1));

        // ---- lazy rescale: move the reference only when the row max grew by more than 2^8 (exponent argument <= 8)
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

        uint32_t pk0[16], pk1[16];
        a2_exp64(v0, v1, pk0, pk1, c2, nm2, l01, l23);
        // ---- P(t-1) consumed and O(t-1) produced before P is overwritten / O is rescaled
        if (t > 0) {
          if (!pv_ok) mbar_wait(&bar_pv_done[X], (t - 1) & 1);
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
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&bar_p_full[X]);
      }

      // ---- epilogue: O / l -> bf16 -> out[b, qpos, h*64 .. +64]
      float s0, s1, s2, s3;
      f2_unpack(l01, s0, s1);
      f2_unpack(l23, s2, s3);
      const float l_run = (s0 + s1) + (s2 + s3);
      mbar_wait(&bar_pv_done[X], (n_steps - 1) & 1);
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
  if (warp == 15) {
    tc_fence_after();
    tmem_dealloc(tmem_base, A2_TMEM_COLS);
  }
}

int warmup_attn3q() {
  return ensure_dyn_smem(reinterpret_cast<const void*>(attn3q_fwd_kernel), A3_SMEM_BYTES, "attn3q_fwd_kernel");
}

// called by pf_attn_fwd_masked (pf_attn.cu) after argument validation
int attn3q_launch(const pf_attn_desc* d, cudaStream_t stream) {
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
  Attn3Args aa{};
  Attn2Args& a = aa.c;
  a.out = static_cast<__nv_bfloat16*>(d->out);
  a.ldo = d->ldo;
  a.batch = d->batch;
  a.heads = d->heads;
  a.seq = d->seq;
  a.q_tiles = (d->seq + A2_BM - 1) / A2_BM;
  a.q_tile_begin = d->q_row_begin / A2_BM;
  a.n_pairs = 0;
  aa.n_groups = (a.q_tiles + A3_G - 1) / A3_G;
  a.scale_log2 = d->scale * 1.4426950408889634f;
  a.seg = d->seg;
  a.time = d->time;
  a.psched = d->group_sched;
  a.sched_stride = d->sched_stride;
  a.pmask_idx = d->group_mask_index;
  a.pmask_bits = static_cast<const uint4*>(d->group_mask_bits);
  a.trace = nullptr;
  a.trace_cap = 0;
  a.timeline = nullptr;
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
  // group g covers tiles q_tiles-3-3g .. q_tiles-1-3g: launch the groups whose top tile is >= q_tile_begin
  const int groups = (a.q_tiles - a.q_tile_begin + A3_G - 1) / A3_G;
  dim3 grid(groups, d->heads, d->batch);
  if (int rc = warmup_attn3q()) return rc;
  attn3q_fwd_kernel<<<grid, A3_THREADS, A3_SMEM_BYTES, stream>>>(tm[0], tm[1], tm[2], aa);
  return check_launch("pf_attn_fwd_masked(three-q-tile kernel)");
}

}  // namespace pf

// Host helpers, generalised from the pair forms: group = q tiles per CTA (2 or 3), counted from the END of the sequence (group g
// = tiles q_tiles - group (g + 1) .. q_tiles - 1 - group g; the first group may miss its leading tiles).  Entry =
// (kv_tile << 8) | flags, flags = 2 bits per tile X at bit 2 X (X = 0 the lowest tile): bit0 the tile has an allowed pair in
// this kv tile, bit1 it needs the element mask.
extern "C" int pf_attn_build_group_schedule(const int32_t* tile_sched, int32_t batch, int32_t seq, int32_t sched_stride,
                                            int32_t group, int32_t* out) {
  using namespace pf;
  PF_REQUIRE(tile_sched && out && batch > 0 && seq > 0, "pf_attn_build_group_schedule: bad arguments");
  PF_REQUIRE(group >= 2 && group <= 4, "pf_attn_build_group_schedule: group %d not in [2, 4]", group);
  const int q_tiles = (seq + 127) / 128;
  PF_REQUIRE(sched_stride >= 1 + q_tiles, "pf_attn_build_group_schedule: stride %d too small", sched_stride);
  const int n_groups = (q_tiles + group - 1) / group;
  for (int b = 0; b < batch; ++b) {
    for (int g = 0; g < n_groups; ++g) {
      const int top = q_tiles - 1 - group * g;
      const int32_t* rows[4] = {nullptr, nullptr, nullptr, nullptr};
      int cnt_in[4] = {0, 0, 0, 0}, pos[4] = {0, 0, 0, 0};
      for (int x = 0; x < group; ++x) {
        const int qt = top - (group - 1 - x);
        if (qt >= 0) {
          rows[x] = tile_sched + (static_cast<size_t>(b) * q_tiles + qt) * sched_stride;
          cnt_in[x] = rows[x][0];
        }
      }
      int32_t* row = out + (static_cast<size_t>(b) * n_groups + g) * sched_stride;
      int cnt = 0;
      for (;;) {
        int kt = 0x7fffffff;
        for (int x = 0; x < group; ++x)
          if (pos[x] < cnt_in[x]) kt = std::min(kt, rows[x][1 + pos[x]] >> 1);
        if (kt == 0x7fffffff) break;
        int flags = 0;
        for (int x = 0; x < group; ++x)
          if (pos[x] < cnt_in[x] && (rows[x][1 + pos[x]] >> 1) == kt) {
            flags |= (1 | ((rows[x][1 + pos[x]] & 1) << 1)) << (2 * x);
            ++pos[x];
          }
        row[1 + cnt] = (kt << 8) | flags;
        ++cnt;
      }
      row[0] = cnt;
      for (int i = 1 + cnt; i < sched_stride; ++i) row[i] = 0;
    }
  }
  return 0;
}

extern "C" int64_t pf_attn_build_group_masks(const int32_t* seg, const int32_t* time, const int32_t* group_sched, int32_t batch,
                                             int32_t seq, int32_t sched_stride, int32_t group, int32_t* mask_index,
                                             uint32_t* mask_bits, int64_t capacity_blocks, const int32_t* pair_sched,
                                             const int32_t* pair_mask_index) {
  using namespace pf;
  if (!seg || !time || !group_sched || !mask_index || batch <= 0 || seq <= 0 || group < 2 || group > 4) {
    set_error("pf_attn_build_group_masks: bad arguments");
    return -1;
  }
  const int q_tiles = (seq + 127) / 128;
  const int n_groups = (q_tiles + group - 1) / group;
  const int n_pairs = (q_tiles + 1) / 2;
  const bool share = pair_sched != nullptr && pair_mask_index != nullptr;   // reuse the pair schedule's blocks: same (q tile, kv tile) masks
  int64_t blocks = 0;
  for (int b = 0; b < batch; ++b) {
    const int32_t* sg = seg + static_cast<size_t>(b) * seq;
    const int32_t* tm = time + static_cast<size_t>(b) * seq;
    for (int g = 0; g < n_groups; ++g) {
      const int32_t* row = group_sched + (static_cast<size_t>(b) * n_groups + g) * sched_stride;
      int32_t* mi = mask_index + (static_cast<size_t>(b) * n_groups + g) * group * sched_stride;
      for (int i = 0; i < group * sched_stride; ++i) mi[i] = -1;
      const int top = q_tiles - 1 - group * g;
      for (int e = 0; e < row[0]; ++e) {
        const int ent = row[1 + e], kt = ent >> 8;
        for (int x = 0; x < group; ++x) {
          const int fl = (ent >> (2 * x)) & 3;
          if (fl != 3) continue;                       // needs bits only when the tile owns the entry AND is partial
          const int qt = top - (group - 1 - x);
          if (share) {
            // q tile qt is tile x_p of pair p; its partial (qt, kt) block was numbered by pf_attn_build_pair_masks
            const int p = (q_tiles - 1 - qt) / 2;
            const int x_p = (qt == q_tiles - 1 - 2 * p) ? 1 : 0;
            const int32_t* prow = pair_sched + (static_cast<size_t>(b) * n_pairs + p) * sched_stride;
            int lo = 0, hi = prow[0] - 1, at = -1;
            while (lo <= hi) {
              const int mid = (lo + hi) / 2, k2 = prow[1 + mid] >> 4;
              if (k2 == kt) { at = mid; break; }
              if (k2 < kt) lo = mid + 1; else hi = mid - 1;
            }
            const int32_t blk = at < 0 ? -1 : pair_mask_index[(static_cast<size_t>(b) * n_pairs + p) * 2 * sched_stride + 2 * at + x_p];
            if (blk < 0) {
              set_error("pf_attn_build_group_masks: no pair block for q tile %d, kv tile %d (batch %d)", qt, kt, b);
              return -1;
            }
            mi[group * e + x] = blk;
            blocks = std::max<int64_t>(blocks, static_cast<int64_t>(blk) + 1);
            continue;
          }
          if (mask_bits != nullptr && blocks < capacity_blocks) {
            attn_build_mask_block(sg, tm, seq, qt, kt, mask_bits + static_cast<size_t>(blocks) * 128 * 4);
          }
          mi[group * e + x] = static_cast<int32_t>(blocks);
          ++blocks;
        }
      }
    }
  }
  return blocks;
}
