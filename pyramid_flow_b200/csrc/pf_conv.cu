// pf_conv.cu — causal 3-D convolution (k = 3x3x3 or 1x1x1, stride 1) as an im2col-free implicit GEMM on tcgen05.
//
// Replaces CausalConv3d -> nn.Conv3d (reference video_vae/modeling_causal_conv.py:116-146, cuDNN conv3d on NCDHW) for
// the VAE decoder.  Activations are channels-last bf16 [B, T, H, W, C]; an output tile is a TH x TW spatial patch of one
// frame (128 voxels = the UMMA M), and the K loop walks (tap, 64-channel chunk):
//   A tile  = ONE 5-D TMA box (64 ch, TW, TH, 1 frame, 1 batch) at the tap-shifted coordinate.  Out-of-bounds spatial
//             coordinates are zero-filled by TMA => the conv's spatial zero padding costs nothing; the causal temporal
//             padding is (kt-1) leading frames physically present in the input buffer (zeros for the first chunk, the
//             previous chunk's last frames afterwards — the reference's feature cache, C:126-143).
//   B tile  = weights re-laid out as [Cout, taps*Cin] (tap-major), a plain 2-D TMA box like the GEMM.
// The box lands in shared memory as [TH][TW][64] = 128 rows of 128 B with SWIZZLE_128B, i.e. exactly the K-major UMMA
// operand; no im2col buffer exists anywhere.  Pipeline / warp roles are the GEMM's (pf_gemm.cu).
// Epilogue: bias (+ residual) -> bf16/fp32 channels-last store, optionally through the depth-to-space addressing of
// CausalUpsample2x (R:616) / CausalTemporalUpsample2x (R:724-727) so the rearrange copy disappears.
#include <cstdlib>

#include "../../include/pf_b200.h"
#include "pf_common.cuh"

namespace pf {

struct ConvArgs {
  int b, t, h, w, cin;
  int cout;            // padded N actually computed (multiple of BN)
  int taps, kt, kh, kw;
  int kw_baseoff;      // kw-reuse kernel: set the A descriptor's base-offset field to the row offset (debug switch)
  int st, sh, sw;      // conv stride (t, h, w): 1, or 2 for the encoder's down-samplers; (b, t, h, w) are OUTPUT dims
  int th, tw, tiles_h, tiles_w;
  int n_tiles;
  const float* bias;
  int store_mode;      // 0 plain, 1 spatial depth-to-space (c p1 p2), 2 temporal depth-to-space (c p)
  void* out;
  int out_f32;
  int out_t_total, out_t_offset, out_h, out_w, out_c;   // geometry of the output buffer
  int store_channels;  // channels of the conv output that are stored (<= cout; the rest is padding)
  const __nv_bfloat16* residual;   // plain mode only; same geometry as out (bf16)
  int res_t_total, res_t_offset;
};

constexpr int CBM = 128;
constexpr int CBK = 64;
constexpr int CONV_THREADS = 256;

template <int BN>
struct ConvCfg {
  static constexpr int A_BYTES = CBM * CBK * 2;
  static constexpr int B_BYTES = BN * CBK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN >= 256) ? 4 : (BN >= 128) ? 6 : 8;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024;
  static constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
};

// Epilogue of one 128-voxel accumulator slice: this thread owns voxel (bb, tt, hh, ww), conv channels [n_base, n_base+BN).
template <int BN>
__device__ __forceinline__ void conv_epilogue_tile(const ConvArgs& g, uint32_t taddr, int bb, int tt, int hh, int ww,
                                                   bool valid, int n_base) {
#pragma unroll 1
  for (int c = 0; c < BN / 16; ++c) {
    uint32_t v[16];
    tmem_ld16(taddr + c * 16, v);
    tmem_ld_wait();
    const int n0 = n_base + c * 16;
    if (!valid || n0 >= g.store_channels) continue;
    float x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = __uint_as_float(v[i]) + (g.bias ? __ldg(g.bias + n0 + i) : 0.f);
    if (g.store_mode == 0) {
      const int to = tt + g.out_t_offset;
      if (to < 0 || to >= g.out_t_total) continue;
      const size_t vox = ((static_cast<size_t>(bb) * g.out_t_total + to) * g.out_h + hh) * g.out_w + ww;
      if (g.residual != nullptr) {
        const size_t rvox = ((static_cast<size_t>(bb) * g.res_t_total + (tt + g.res_t_offset)) * g.out_h + hh) * g.out_w + ww;
        const uint4* r4 = reinterpret_cast<const uint4*>(g.residual + rvox * g.out_c + n0);
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const uint4 rv = __ldg(r4 + u);
          const __nv_bfloat162* hv = reinterpret_cast<const __nv_bfloat162*>(&rv);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 f = __bfloat1622float2(hv[i]);
            x[8 * u + 2 * i] += f.x;
            x[8 * u + 2 * i + 1] += f.y;
          }
        }
      }
      const int nvalid = g.store_channels - n0;
      if (g.out_f32 == 2) {
        // uint8 image store of decode_latent (P:1238): clamp(v * 127.5 + 127.5, 0, 255), truncated like torch's .byte()
        uint8_t* dst = reinterpret_cast<uint8_t*>(g.out) + vox * g.out_c + n0;
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (i < nvalid) dst[i] = static_cast<uint8_t>(fminf(fmaxf(fmaf(x[i], 127.5f, 127.5f), 0.f), 255.f));
      } else if (g.out_f32) {
        float* dst = reinterpret_cast<float*>(g.out) + vox * g.out_c + n0;
        if (nvalid >= 16 && (g.out_c & 3) == 0) {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            reinterpret_cast<float4*>(dst)[i] = make_float4(x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]);
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (i < nvalid) dst[i] = x[i];
        }
      } else {
        __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(g.out) + vox * g.out_c + n0;
        if (nvalid >= 16 && (g.out_c & 7) == 0) {
          uint4 u0, u1;
          u0.x = pack_bf16x2(x[0], x[1]); u0.y = pack_bf16x2(x[2], x[3]); u0.z = pack_bf16x2(x[4], x[5]); u0.w = pack_bf16x2(x[6], x[7]);
          u1.x = pack_bf16x2(x[8], x[9]); u1.y = pack_bf16x2(x[10], x[11]); u1.z = pack_bf16x2(x[12], x[13]); u1.w = pack_bf16x2(x[14], x[15]);
          reinterpret_cast<uint4*>(dst)[0] = u0;
          reinterpret_cast<uint4*>(dst)[1] = u1;
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (i < nvalid) dst[i] = __float2bfloat16(x[i]);
        }
      }
    } else if (g.store_mode == 1) {
      // 'b (c p1 p2) t h w -> b c t (h p1) (w p2)': conv channel n = 4c + 2 p1 + p2; 16 n = 4 output channels x 4 pixels
      const int to = tt + g.out_t_offset;
      if (to < 0 || to >= g.out_t_total) continue;
      const int c0 = n0 >> 2;
      __nv_bfloat16* base = reinterpret_cast<__nv_bfloat16*>(g.out);
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const int p1 = p >> 1, p2 = p & 1;
        const size_t vox = ((static_cast<size_t>(bb) * g.out_t_total + to) * g.out_h + (2 * hh + p1)) * g.out_w + (2 * ww + p2);
        uint2 u;
        u.x = pack_bf16x2(x[0 + p], x[4 + p]);
        u.y = pack_bf16x2(x[8 + p], x[12 + p]);
        *reinterpret_cast<uint2*>(base + vox * g.out_c + c0) = u;
      }
    } else {
      // 'b (c p) t h w -> b c (t p) h w': conv channel n = 2c + p; frame 2t + p (+offset; negative = dropped frame)
      const int c0 = n0 >> 1;
      __nv_bfloat16* base = reinterpret_cast<__nv_bfloat16*>(g.out);
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        const int to = 2 * tt + p + g.out_t_offset;
        if (to < 0 || to >= g.out_t_total) continue;
        const size_t vox = ((static_cast<size_t>(bb) * g.out_t_total + to) * g.out_h + hh) * g.out_w + ww;
        uint4 u;
        u.x = pack_bf16x2(x[0 + p], x[2 + p]);
        u.y = pack_bf16x2(x[4 + p], x[6 + p]);
        u.z = pack_bf16x2(x[8 + p], x[10 + p]);
        u.w = pack_bf16x2(x[12 + p], x[14 + p]);
        *reinterpret_cast<uint4*>(base + vox * g.out_c + c0) = u;
      }
    }
  }
}

template <int BN>
__global__ void __launch_bounds__(CONV_THREADS, 1)
conv3d_tc_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w, const ConvArgs g) {
  using Cfg = ConvCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  __shared__ __align__(8) uint64_t full_bar[STAGES];
  __shared__ __align__(8) uint64_t empty_bar[STAGES];
  __shared__ __align__(8) uint64_t tmem_full_bar[2];
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&tm_w);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], 128);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(&tmem_base_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  const int cchunks = g.cin / CBK;
  const int num_kb = g.taps * cchunks;
  // tile index -> (n tile fastest, then spatial tile, frame, batch): CTAs running together share the same input patch
  const int sp_tiles = g.tiles_h * g.tiles_w;
  const long long total_tiles = static_cast<long long>(g.b) * g.t * sp_tiles * g.n_tiles;

  if (warp == 0 && elect_one()) {   // elect.sync: the compiler keeps the role's code on the uniform datapath
    // ===== TMA producer =====
    int stage = 0;
    uint32_t phase = 0;
    const int ph = g.kh >> 1, pw = g.kw >> 1;
    for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int nt = static_cast<int>(tile % g.n_tiles);
      long long r = tile / g.n_tiles;
      const int sp = static_cast<int>(r % sp_tiles);
      r /= sp_tiles;
      const int tt = static_cast<int>(r % g.t);
      const int bb = static_cast<int>(r / g.t);
      const int h0 = (sp / g.tiles_w) * g.th, w0 = (sp % g.tiles_w) * g.tw;
      for (int kb = 0; kb < num_kb; ++kb) {
        // K order (dt, dh, channel chunk, dw): the same accumulation order as the kw-reuse kernel, so which kernel a call
        // is dispatched to (it depends on the number of tiles, i.e. on the temporal chunking) never changes the bits
        const int grp = kb / g.kw;
        const int dw = kb - grp * g.kw;
        const int dtdh = grp / cchunks;
        const int cc = grp - dtdh * cchunks;
        const int dt = dtdh / g.kh, dh = dtdh - dt * g.kh;
        const int wk = (dtdh * g.kw + dw) * cchunks + cc;     // K block of tap (dt, dh, dw), chunk cc in the weight matrix
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
        mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
        tma_load_5d(sa, &tm_x, &full_bar[stage], cc * CBK, w0 * g.sw + dw - pw, h0 * g.sh + dh - ph, tt * g.st + dt, bb);
        tma_load_2d(sa + Cfg::A_BYTES, &tm_w, &full_bar[stage], wk * CBK, nt * BN);
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1 && elect_one()) {
    // ===== MMA issuer =====
    constexpr uint32_t idesc = make_idesc_bf16(CBM, BN, 0, 0);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * BN;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
        const uint64_t da = make_smem_desc_kmajor_sw128(sa);
        const uint64_t db = make_smem_desc_kmajor_sw128(sa + Cfg::A_BYTES);
#pragma unroll
        for (int kk = 0; kk < CBK / 16; ++kk) umma_ss(tmem_d, da + 2 * kk, db + 2 * kk, idesc, (kb | kk) != 0 ? 1u : 0u);
        umma_commit(&empty_bar[stage]);
        if (kb == num_kb - 1) umma_commit(&tmem_full_bar[acc]);
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: thread == output voxel =====
    const int q = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    const int rrow = q * 32 + lane;
    const int lh = rrow / g.tw, lw = rrow - lh * g.tw;
    for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int nt = static_cast<int>(tile % g.n_tiles);
      long long r = tile / g.n_tiles;
      const int sp = static_cast<int>(r % sp_tiles);
      r /= sp_tiles;
      const int tt = static_cast<int>(r % g.t);
      const int bb = static_cast<int>(r / g.t);
      const int hh = (sp / g.tiles_w) * g.th + lh, ww = (sp % g.tiles_w) * g.tw + lw;
      const bool valid = hh < g.h && ww < g.w;
      const int n_base = nt * BN;

      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN;
      conv_epilogue_tile<BN>(g, taddr, bb, tt, hh, ww, valid, n_base);
      tc_fence_before();
      mbar_arrive(&tmem_empty_bar[acc]);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// 2-CTA variant: a CTA pair computes two spatially adjacent 128-voxel patches x BN channels with ONE 256-row MMA; each CTA
// loads its own input box and half of the weight tile (protocol as gemm2_bf16_tc_kernel in pf_gemm.cu).
// ---------------------------------------------------------------------------------------------------------------
template <int BN>
struct Conv2Cfg {
  static constexpr int A_BYTES = CBM * CBK * 2;
  static constexpr int B_BYTES = (BN / 2) * CBK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN >= 256) ? 6 : 8;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024;
  static constexpr int TMEM_COLS = (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
};

template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(CONV_THREADS, 1)
conv3d2_tc_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w, const ConvArgs g) {
  using Cfg = Conv2Cfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  __shared__ __align__(8) uint64_t full_bar[STAGES];
  __shared__ __align__(8) uint64_t empty_bar[STAGES];
  __shared__ __align__(8) uint64_t tmem_full_bar[2];
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&tm_w);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 2);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], 256);
    }
    fence_barrier_init();
  }
  cluster_sync_all();
  if (warp == 2) {
    tmem_alloc2(&tmem_base_slot, Cfg::TMEM_COLS);
    tmem_relinquish2();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  const int cchunks = g.cin / CBK;
  const int num_kb = g.taps * cchunks;
  const int sp_tiles = g.tiles_h * g.tiles_w;
  const int sp_pairs = (sp_tiles + 1) / 2;
  const long long total_tiles = static_cast<long long>(g.b) * g.t * sp_pairs * g.n_tiles;
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;

  // decode: n tile fastest, then patch pair, frame, batch; this CTA owns patch 2*pair + rank (may fall off the end)
  auto decode = [&](long long tile, int& nt, int& tt, int& bb, int& h0, int& w0) {
    nt = static_cast<int>(tile % g.n_tiles);
    long long r = tile / g.n_tiles;
    const int sp = static_cast<int>(r % sp_pairs) * 2 + static_cast<int>(rank);
    r /= sp_pairs;
    tt = static_cast<int>(r % g.t);
    bb = static_cast<int>(r / g.t);
    if (sp < sp_tiles) {
      h0 = (sp / g.tiles_w) * g.th;
      w0 = (sp % g.tiles_w) * g.tw;
    } else {
      h0 = g.h + g.th;   // entirely outside: TMA zero-fills, the epilogue stores nothing
      w0 = 0;
    }
  };

  if (warp == 0 && elect_one()) {
    int stage = 0;
    uint32_t phase = 0;
    const int ph = g.kh >> 1, pw = g.kw >> 1;
    for (long long tile = cluster_id; tile < total_tiles; tile += num_clusters) {
      int nt, tt, bb, h0, w0;
      decode(tile, nt, tt, bb, h0, w0);
      for (int kb = 0; kb < num_kb; ++kb) {
        // K order (dt, dh, channel chunk, dw): the same accumulation order as the kw-reuse kernel, so which kernel a call
        // is dispatched to (it depends on the number of tiles, i.e. on the temporal chunking) never changes the bits
        const int grp = kb / g.kw;
        const int dw = kb - grp * g.kw;
        const int dtdh = grp / cchunks;
        const int cc = grp - dtdh * cchunks;
        const int dt = dtdh / g.kh, dh = dtdh - dt * g.kh;
        const int wk = (dtdh * g.kw + dw) * cchunks + cc;     // K block of tap (dt, dh, dw), chunk cc in the weight matrix
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
        if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
        else mbar_arrive_remote(&full_bar[stage], 0);
        tma_load_5d_2cta(sa, &tm_x, &full_bar[stage], cc * CBK, w0 * g.sw + dw - pw, h0 * g.sh + dh - ph, tt * g.st + dt, bb);
        tma_load_2d_2cta(sa + Cfg::A_BYTES, &tm_w, &full_bar[stage], wk * CBK, nt * BN + static_cast<int>(rank) * (BN / 2));
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1 && leader && elect_one()) {
    constexpr uint32_t idesc = make_idesc_bf16(2 * CBM, BN, 0, 0);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (long long tile = cluster_id; tile < total_tiles; tile += num_clusters) {
      mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * BN;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
        const uint64_t da = make_smem_desc_kmajor_sw128(sa);
        const uint64_t db = make_smem_desc_kmajor_sw128(sa + Cfg::A_BYTES);
#pragma unroll
        for (int kk = 0; kk < CBK / 16; ++kk) umma_ss_2cta(tmem_d, da + 2 * kk, db + 2 * kk, idesc, (kb | kk) != 0 ? 1u : 0u);
        umma_commit_2cta(&empty_bar[stage]);
        if (kb == num_kb - 1) umma_commit_2cta(&tmem_full_bar[acc]);
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    const int rrow = q * 32 + lane;
    const int lh = rrow / g.tw, lw = rrow - lh * g.tw;
    for (long long tile = cluster_id; tile < total_tiles; tile += num_clusters) {
      int nt, tt, bb, h0, w0;
      decode(tile, nt, tt, bb, h0, w0);
      const int hh = h0 + lh, ww = w0 + lw;
      const bool valid = hh < g.h && ww < g.w;
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN;
      conv_epilogue_tile<BN>(g, taddr, bb, tt, hh, ww, valid, nt * BN);
      tc_fence_before();
      if (leader) mbar_arrive(&tmem_empty_bar[acc]);
      else mbar_arrive_remote(&tmem_empty_bar[acc], 0);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, Cfg::TMEM_COLS);
  }
}

template <int BN>
static int launch_conv2(const CUtensorMap& tm_x, const CUtensorMap& tm_w, const ConvArgs& g, cudaStream_t stream) {
  using Cfg = Conv2Cfg<BN>;
  auto kern = conv3d2_tc_kernel<BN>;
  if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(kern), Cfg::SMEM_BYTES, "conv2")) return rc;
  const int sp_pairs = (g.tiles_h * g.tiles_w + 1) / 2;
  const long long total = static_cast<long long>(g.b) * g.t * sp_pairs * g.n_tiles;
  int sms = num_sms();
  if (sms <= 0) sms = 148;
  long long clusters = sms / 2;
  if (total < clusters) clusters = total;
  kern<<<static_cast<int>(2 * clusters), CONV_THREADS, Cfg::SMEM_BYTES, stream>>>(tm_x, tm_w, g);
  return check_launch("pf_causal_conv3d(2cta)");
}

template <int BN>
struct Conv2wCfg {
  // kw-tap reuse: ONE haloed input row of 128 + 2 voxels (x 64 channels) per (kt, kh, channel chunk) feeds the three kw
  // taps -- the A descriptor of tap kw starts kw rows (kw * 128 B) into the tile -- so the input patch crosses L2 -> smem
  // 9x instead of 27x (ncu on the 128->128 full-resolution conv: tensor pipe 46 %, operand traffic bound).
  static constexpr int A_ROWS = CBM + 2;
  static constexpr int A_TX = A_ROWS * CBK * 2;                   // bytes the TMA delivers
  static constexpr int A_BYTES = (A_TX + 1023) / 1024 * 1024;     // padded: the weight tiles stay 1024-aligned
  static constexpr int B_BYTES = (BN / 2) * CBK * 2;              // one tap, this CTA's half of the filters
  static constexpr int STAGE_BYTES = A_BYTES + 3 * B_BYTES;
  static constexpr int STAGES = (226 * 1024) / STAGE_BYTES;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024;
  static constexpr int TMEM_COLS = (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
};

template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(CONV_THREADS, 1)
conv3d2w_tc_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w, const ConvArgs g) {
  using Cfg = Conv2wCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  __shared__ __align__(8) uint64_t full_bar[STAGES];
  __shared__ __align__(8) uint64_t empty_bar[STAGES];
  __shared__ __align__(8) uint64_t tmem_full_bar[2];
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&tm_w);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 2);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], 256);
    }
    fence_barrier_init();
  }
  cluster_sync_all();
  if (warp == 2) {
    tmem_alloc2(&tmem_base_slot, Cfg::TMEM_COLS);
    tmem_relinquish2();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  const int cchunks = g.cin / CBK;
  const int num_grp = g.kt * g.kh * cchunks;   // (dt, dh, channel chunk) groups, three kw taps each
  const int sp_tiles = g.tiles_h * g.tiles_w;
  const int sp_pairs = (sp_tiles + 1) / 2;
  const long long total_tiles = static_cast<long long>(g.b) * g.t * sp_pairs * g.n_tiles;
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;

  // decode: n tile fastest, then patch pair, frame, batch; this CTA owns patch 2*pair + rank (may fall off the end)
  auto decode = [&](long long tile, int& nt, int& tt, int& bb, int& h0, int& w0) {
    nt = static_cast<int>(tile % g.n_tiles);
    long long r = tile / g.n_tiles;
    const int sp = static_cast<int>(r % sp_pairs) * 2 + static_cast<int>(rank);
    r /= sp_pairs;
    tt = static_cast<int>(r % g.t);
    bb = static_cast<int>(r / g.t);
    if (sp < sp_tiles) {
      h0 = (sp / g.tiles_w) * g.th;
      w0 = (sp % g.tiles_w) * g.tw;
    } else {
      h0 = g.h + g.th;   // entirely outside: TMA zero-fills, the epilogue stores nothing
      w0 = 0;
    }
  };

  if (warp == 0 && elect_one()) {
    int stage = 0;
    uint32_t phase = 0;
    const int ph = g.kh >> 1, pw = g.kw >> 1;
    for (long long tile = cluster_id; tile < total_tiles; tile += num_clusters) {
      int nt, tt, bb, h0, w0;
      decode(tile, nt, tt, bb, h0, w0);
      for (int grp = 0; grp < num_grp; ++grp) {
        const int dtdh = grp / cchunks;
        const int cc = grp - dtdh * cchunks;
        const int dt = dtdh / g.kh, dh = dtdh - dt * g.kh;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
        if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * (Cfg::A_TX + 3 * Cfg::B_BYTES));
        else mbar_arrive_remote(&full_bar[stage], 0);
        tma_load_5d_2cta(sa, &tm_x, &full_bar[stage], cc * CBK, w0 - pw, h0 + dh - ph, tt + dt, bb);   // 130 voxels
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const int kb = (dtdh * 3 + kw) * cchunks + cc;        // K block of tap (dt, dh, kw), channel chunk cc
          tma_load_2d_2cta(sa + Cfg::A_BYTES + kw * Cfg::B_BYTES, &tm_w, &full_bar[stage], kb * CBK,
                           nt * BN + static_cast<int>(rank) * (BN / 2));
        }
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1 && leader && elect_one()) {
    constexpr uint32_t idesc = make_idesc_bf16(2 * CBM, BN, 0, 0);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (long long tile = cluster_id; tile < total_tiles; tile += num_clusters) {
      mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * BN;
      for (int grp = 0; grp < num_grp; ++grp) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          // rows [kw, kw + 128) of the haloed tile: start address advanced by kw 128-byte rows inside the 1024-byte swizzle
          // atom.  The swizzle is applied on absolute smem address bits, so the advanced descriptor reads exactly what
          // the TMA wrote; the base-offset field [49,52) must stay 0 (pinned on hardware: tools/gpu_check.py probe_rowoff)
          const uint64_t da = make_smem_desc_kmajor_sw128(sa + kw * 128) | (static_cast<uint64_t>(g.kw_baseoff ? kw : 0) << 49);
          const uint64_t db = make_smem_desc_kmajor_sw128(sa + Cfg::A_BYTES + kw * Cfg::B_BYTES);
#pragma unroll
          for (int kk = 0; kk < CBK / 16; ++kk)
            umma_ss_2cta(tmem_d, da + 2 * kk, db + 2 * kk, idesc, (grp | kw | kk) != 0 ? 1u : 0u);
        }
        umma_commit_2cta(&empty_bar[stage]);
        if (grp == num_grp - 1) umma_commit_2cta(&tmem_full_bar[acc]);
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    const int rrow = q * 32 + lane;
    const int lh = rrow / g.tw, lw = rrow - lh * g.tw;
    for (long long tile = cluster_id; tile < total_tiles; tile += num_clusters) {
      int nt, tt, bb, h0, w0;
      decode(tile, nt, tt, bb, h0, w0);
      const int hh = h0 + lh, ww = w0 + lw;
      const bool valid = hh < g.h && ww < g.w;
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN;
      conv_epilogue_tile<BN>(g, taddr, bb, tt, hh, ww, valid, nt * BN);
      tc_fence_before();
      if (leader) mbar_arrive(&tmem_empty_bar[acc]);
      else mbar_arrive_remote(&tmem_empty_bar[acc], 0);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, Cfg::TMEM_COLS);
  }
}

template <int BN>
static int launch_conv2w(const CUtensorMap& tm_x, const CUtensorMap& tm_w, const ConvArgs& g, cudaStream_t stream) {
  using Cfg = Conv2wCfg<BN>;
  auto kern = conv3d2w_tc_kernel<BN>;
  if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(kern), Cfg::SMEM_BYTES, "conv2")) return rc;
  const int sp_pairs = (g.tiles_h * g.tiles_w + 1) / 2;
  const long long total = static_cast<long long>(g.b) * g.t * sp_pairs * g.n_tiles;
  int sms = num_sms();
  if (sms <= 0) sms = 148;
  long long clusters = sms / 2;
  if (total < clusters) clusters = total;
  kern<<<static_cast<int>(2 * clusters), CONV_THREADS, Cfg::SMEM_BYTES, stream>>>(tm_x, tm_w, g);
  return check_launch("pf_causal_conv3d(2cta, kw reuse)");
}

template <int BN>
static int launch_conv(const CUtensorMap& tm_x, const CUtensorMap& tm_w, const ConvArgs& g, cudaStream_t stream) {
  using Cfg = ConvCfg<BN>;
  auto kern = conv3d_tc_kernel<BN>;
  if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(kern), Cfg::SMEM_BYTES, "conv")) return rc;
  const long long total = static_cast<long long>(g.b) * g.t * g.tiles_h * g.tiles_w * g.n_tiles;
  int grid = num_sms();
  if (grid <= 0) grid = 148;
  if (total < grid) grid = static_cast<int>(total);
  kern<<<grid, CONV_THREADS, Cfg::SMEM_BYTES, stream>>>(tm_x, tm_w, g);
  return check_launch("pf_causal_conv3d");
}

int warmup_conv() {
  int rc = 0;
#define PF_WARM(KERN, CFG) if (!rc) rc = ensure_dyn_smem(reinterpret_cast<const void*>(KERN), CFG::SMEM_BYTES, #KERN)
  PF_WARM((conv3d_tc_kernel<256>), ConvCfg<256>);
  PF_WARM((conv3d_tc_kernel<128>), ConvCfg<128>);
  PF_WARM((conv3d_tc_kernel<64>), ConvCfg<64>);
  PF_WARM((conv3d2_tc_kernel<256>), Conv2Cfg<256>);
  PF_WARM((conv3d2_tc_kernel<128>), Conv2Cfg<128>);
  PF_WARM((conv3d2w_tc_kernel<256>), Conv2wCfg<256>);
  PF_WARM((conv3d2w_tc_kernel<128>), Conv2wCfg<128>);
#undef PF_WARM
  return rc;
}

}  // namespace pf

extern "C" int pf_causal_conv3d(const pf_conv3d_desc* d, void* stream_) {
  using namespace pf;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PF_REQUIRE(d && d->x && d->wgt && d->out, "pf_causal_conv3d: null pointer");
  PF_REQUIRE(d->cin % 64 == 0 && d->cin > 0, "pf_causal_conv3d: cin=%d must be a multiple of 64 (pad the channels)", d->cin);
  PF_REQUIRE(d->cout % 64 == 0 && d->cout > 0, "pf_causal_conv3d: cout=%d must be a multiple of 64 (pad the filters)", d->cout);
  PF_REQUIRE((d->kt == 1 || d->kt == 3) && (d->kh == 1 || d->kh == 3) && d->kw == d->kh, "pf_causal_conv3d: kernel must be 1x1x1 or 3x3x3 (kt in {1,3})");
  PF_REQUIRE(d->b > 0 && d->t > 0 && d->h > 0 && d->w > 0, "pf_causal_conv3d: bad shape");
  PF_REQUIRE(d->store_mode >= 0 && d->store_mode <= 2, "pf_causal_conv3d: bad store_mode");
  PF_REQUIRE(d->store_channels > 0 && d->store_channels <= d->cout, "pf_causal_conv3d: bad store_channels");
  if (d->store_mode == 1) PF_REQUIRE(d->store_channels == d->cout && d->out_c * 4 == d->cout && !d->out_f32 && !d->residual, "pf_causal_conv3d: spatial depth-to-space needs out_c = cout/4, bf16, no residual");
  if (d->store_mode == 2) PF_REQUIRE(d->store_channels == d->cout && d->out_c * 2 == d->cout && !d->out_f32 && !d->residual, "pf_causal_conv3d: temporal depth-to-space needs out_c = cout/2, bf16, no residual");
  if (d->store_mode == 0) PF_REQUIRE(d->out_c >= d->store_channels, "pf_causal_conv3d: out_c < store_channels");
  PF_REQUIRE(d->out_f32 >= 0 && d->out_f32 <= 2 && (d->out_f32 != 2 || !d->residual), "pf_causal_conv3d: out_f32 must be 0 (bf16), 1 (fp32) or 2 (uint8 image, no residual)");
  const int st = d->stride_t > 1 ? d->stride_t : 1, sh = d->stride_h > 1 ? d->stride_h : 1, sw = d->stride_w > 1 ? d->stride_w : 1;
  PF_REQUIRE(st <= 2 && sh <= 2 && sw <= 2 && sh == sw, "pf_causal_conv3d: strides must be 1 or 2 with stride_h == stride_w");
  if (st > 1 || sh > 1) PF_REQUIRE(d->store_mode == 0 && d->kt == 3 && d->kh == 3, "pf_causal_conv3d: strided convs are plain-store 3x3x3 (CausalDownsample2x R:322, CausalTemporalDownsample2x R:486)");

  ConvArgs g{};
  g.b = d->b; g.t = d->t; g.h = d->h; g.w = d->w; g.cin = d->cin; g.cout = d->cout;
  g.kt = d->kt; g.kh = d->kh; g.kw = d->kw; g.taps = d->kt * d->kh * d->kw;
  g.st = st; g.sh = sh; g.sw = sw;
  int tw = 128;
  while (tw > 8 && tw / 2 >= d->w) tw >>= 1;   // smallest power of two >= w, clamped to [8, 128]
  g.tw = tw; g.th = 128 / tw;
  g.tiles_w = (d->w + g.tw - 1) / g.tw;
  g.tiles_h = (d->h + g.th - 1) / g.th;
  const int bn = (d->cout % 256 == 0) ? 256 : (d->cout % 128 == 0) ? 128 : 64;
  g.n_tiles = d->cout / bn;
  g.bias = d->bias;
  g.store_mode = d->store_mode;
  g.out = d->out; g.out_f32 = d->out_f32;
  g.out_t_total = d->out_t_total; g.out_t_offset = d->out_t_offset;
  g.out_h = d->store_mode == 1 ? 2 * d->h : d->h;
  g.out_w = d->store_mode == 1 ? 2 * d->w : d->w;
  g.out_c = d->out_c;
  g.store_channels = d->store_channels;
  g.residual = static_cast<const __nv_bfloat16*>(d->residual);
  g.res_t_total = d->res_t_total; g.res_t_offset = d->res_t_offset;

  // 2-CTA tiles when there is enough work for 74 CTA pairs; kernel_variant pins a kernel (tests)
  PF_REQUIRE(d->kernel_variant >= 0 && d->kernel_variant <= 3, "pf_causal_conv3d: bad kernel_variant %d", d->kernel_variant);
  bool two_cta = bn >= 128 && static_cast<long long>(d->b) * d->t * g.tiles_h * g.tiles_w * g.n_tiles >= 296;
  if (d->kernel_variant == 1) two_cta = false;
  if (d->kernel_variant >= 2) {
    PF_REQUIRE(bn >= 128, "pf_causal_conv3d: kernel_variant %d (2-CTA) needs cout %% 128 == 0", d->kernel_variant);
    two_cta = true;
  }
  // input geometry: (t-1)*st + kt frames (the kt-1 causal frames physically first), h*sh x w*sw voxels (symmetric pad 1 is
  // the TMA's out-of-bounds zero fill).  A strided conv loads every sh-th / sw-th voxel of a (th*sh) x (tw*sw) box.
  // kw-tap reuse (conv3d2w): full 128-voxel rows, 3x3x3, unit stride.
  // Measured on B200: 128->128 on 2x768x1280 1.65 -> 1.01 ms (1057 -> 1724 TFLOP/s), 256->256 on 2x384x640 1765 -> 1888.
  const bool kwr_ok = g.th == 1 && g.tw == 128 && d->kt == 3 && d->kh == 3 && st == 1 && sh == 1;
  if (d->kernel_variant == 3) PF_REQUIRE(kwr_ok, "pf_causal_conv3d: kernel_variant 3 (kw reuse) needs w > 64, a 3x3x3 kernel and unit stride");
  const bool kwr = two_cta && kwr_ok && d->kernel_variant != 2;
  g.kw_baseoff = 0;   // the UMMA descriptor's base-offset field must stay 0 (pinned by the probe, tools/gpu_check.py probe_rowoff)
  const int tin = (d->t - 1) * st + d->kt;
  const int hin = d->h * sh, win = d->w * sw;
  CUtensorMap tm_x, tm_w;
  {
    const uint64_t dims[5] = {static_cast<uint64_t>(d->cin), static_cast<uint64_t>(win), static_cast<uint64_t>(hin),
                              static_cast<uint64_t>(tin), static_cast<uint64_t>(d->b)};
    const uint64_t s0 = static_cast<uint64_t>(d->cin) * 2;
    const uint64_t strides[4] = {s0, s0 * win, s0 * win * hin, s0 * win * hin * tin};
    const uint32_t box[5] = {CBK, static_cast<uint32_t>(kwr ? g.tw + 2 : g.tw * sw), static_cast<uint32_t>(g.th * sh), 1, 1};
    const uint32_t estr[5] = {1, static_cast<uint32_t>(sw), static_cast<uint32_t>(sh), 1, 1};
    int rc = encode_tensor_map(&tm_x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, d->x, dims, strides, box,
                               CU_TENSOR_MAP_SWIZZLE_128B, estr);
    if (rc) return rc;
  }
  {
    const uint64_t kdim = static_cast<uint64_t>(g.taps) * d->cin;
    const uint64_t dims[2] = {kdim, static_cast<uint64_t>(d->cout)};
    const uint64_t strides[1] = {kdim * 2};
    const uint32_t box[2] = {CBK, static_cast<uint32_t>(two_cta ? bn / 2 : bn)};
    int rc = encode_tensor_map(&tm_w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d->wgt, dims, strides, box,
                               CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  if (kwr) {
    if (bn == 256) return launch_conv2w<256>(tm_x, tm_w, g, stream);
    return launch_conv2w<128>(tm_x, tm_w, g, stream);
  }
  if (two_cta) {
    if (bn == 256) return launch_conv2<256>(tm_x, tm_w, g, stream);
    return launch_conv2<128>(tm_x, tm_w, g, stream);
  }
  switch (bn) {
    case 256: return launch_conv<256>(tm_x, tm_w, g, stream);
    case 128: return launch_conv<128>(tm_x, tm_w, g, stream);
    default: return launch_conv<64>(tm_x, tm_w, g, stream);
  }
}
