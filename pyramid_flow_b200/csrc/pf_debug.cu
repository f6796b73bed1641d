// pf_debug.cu — single-CTA tcgen05 probe: pins UMMA descriptor encodings on hardware (tests/tools only).
//
// D[128, N] = A[128, K] . B  with B either K-major ([N, K] row-major) or MN-major ([K, N] row-major), and A either
// from shared memory (TMA, SWIZZLE_128B) or from tensor memory (packed bf16 pairs written with tcgen05.st).
// All descriptor fields that are uncertain are host-provided, so several hypotheses cost one GPU call.
#include "../../include/pf_b200.h"
#include "pf_common.cuh"

namespace pf {

struct ProbeArgs {
  const __nv_bfloat16* a;
  float* d;
  int n, k;
  int b_row_blocks, b_col_blocks, b_box_rows;
  int b_mn_major;
  uint32_t b_lbo, b_sbo, b_k_step_bytes, b_kblock_bytes;
  int a_from_tmem;
  int a_rows, a_row_offset, a_base_offset;   // A tile of a_rows (>= 128) rows in smem, MMA reads rows [off, off+128)
};

__global__ void __launch_bounds__(128, 1)
umma_probe_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                  const ProbeArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t load_bar;
  __shared__ __align__(8) uint64_t mma_bar;
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int a_box_bytes = (p.a_rows * 128 + 1023) & ~1023;   // one 64-wide k block: a_rows x 128 B, padded to the swizzle atom
  const int a_bytes = a_box_bytes * (p.k / 64);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + a_bytes;
  const int box_bytes = p.b_box_rows * 128;

  if (tid == 0) {
    mbar_init(&load_bar, 1);
    mbar_init(&mma_bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const uint32_t tmem_d = tmem_base;        // columns [0, N)
  const uint32_t tmem_a = tmem_base + 256;  // columns [256, 256 + K/2)

  if (tid == 0) {
    uint32_t bytes = p.b_row_blocks * p.b_col_blocks * box_bytes;
    if (!p.a_from_tmem) bytes += p.a_rows * 128 * (p.k / 64);
    mbar_arrive_expect_tx(&load_bar, bytes);
    if (!p.a_from_tmem) {
      for (int kb = 0; kb < p.k / 64; ++kb) tma_load_2d(smem_a + kb * a_box_bytes, &tm_a, &load_bar, kb * 64, 0);
    }
    int idx = 0;
    for (int cb = 0; cb < p.b_col_blocks; ++cb)
      for (int rb = 0; rb < p.b_row_blocks; ++rb, ++idx)
        tma_load_2d(smem_b + idx * box_bytes, &tm_b, &load_bar, cb * 64, rb * p.b_box_rows);
  }
  if (p.a_from_tmem) {
    // lane == row; 32-bit column c holds (a[row, 2c], a[row, 2c+1])
    const uint32_t* arow = reinterpret_cast<const uint32_t*>(p.a + static_cast<size_t>(tid) * p.k);
    for (int c0 = 0; c0 < p.k / 2; c0 += 16) {
      uint32_t v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = arow[c0 + i];
      tmem_st16(tmem_a + (static_cast<uint32_t>(warp * 32) << 16) + c0, v);
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (tid == 0) {
    mbar_wait(&load_bar, 0);
    tc_fence_after();
    const uint32_t idesc = make_idesc_bf16(128, p.n, 0, p.b_mn_major);
    const uint32_t sa = smem_u32(smem_a);
    const uint32_t sb = smem_u32(smem_b);
    for (int j = 0; j < p.k / 16; ++j) {
      const uint32_t b_off = (j / 4) * p.b_kblock_bytes + (j % 4) * p.b_k_step_bytes;
      const uint64_t db = make_smem_desc(sb + b_off, p.b_lbo, p.b_sbo);
      if (p.a_from_tmem) {
        umma_ts(tmem_d, tmem_a + j * 8, db, idesc, j != 0);
      } else {
        // row offset: the start address moves by whole 128-byte rows inside the 1024-byte swizzle atom; the descriptor's
        // base-offset field [49,52) carries (start >> 7) & 7 (hypothesis under test when a_row_offset != 0)
        const uint64_t da = make_smem_desc_kmajor_sw128(sa + (j / 4) * a_box_bytes + p.a_row_offset * 128 + (j % 4) * 32) |
                            (static_cast<uint64_t>(p.a_base_offset & 7) << 49);
        umma_ss(tmem_d, da, db, idesc, j != 0);
      }
    }
    umma_commit(&mma_bar);
  }
  mbar_wait(&mma_bar, 0);
  tc_fence_after();
  for (int c = 0; c < p.n; c += 16) {
    uint32_t v[16];
    tmem_ld16(tmem_d + (static_cast<uint32_t>(warp * 32) << 16) + c, v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i) p.d[static_cast<size_t>(tid) * p.n + c + i] = __uint_as_float(v[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace pf

extern "C" int pf_debug_umma(const pf_umma_probe* p, void* stream_) {
  using namespace pf;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PF_REQUIRE(p && p->a && p->b && p->d, "pf_debug_umma: null pointer");
  PF_REQUIRE(p->k % 64 == 0 && p->k >= 64 && p->k <= 256, "pf_debug_umma: k must be 64..256, multiple of 64");
  PF_REQUIRE(p->n % 16 == 0 && p->n >= 16 && p->n <= 256, "pf_debug_umma: n must be 16..256, multiple of 16");
  PF_REQUIRE(p->b_cols % 64 == 0 && p->b_rows % p->b_box_rows == 0 && p->b_box_rows <= 256, "pf_debug_umma: bad b box");
  ProbeArgs a{};
  a.a = static_cast<const __nv_bfloat16*>(p->a);
  a.d = p->d;
  a.n = p->n;
  a.k = p->k;
  a.b_row_blocks = p->b_rows / p->b_box_rows;
  a.b_col_blocks = p->b_cols / 64;
  a.b_box_rows = p->b_box_rows;
  a.b_mn_major = p->b_mn_major;
  a.b_lbo = p->b_lbo;
  a.b_sbo = p->b_sbo;
  a.b_k_step_bytes = p->b_k_step_bytes;
  a.b_kblock_bytes = p->b_kblock_bytes;
  a.a_from_tmem = p->a_from_tmem;
  a.a_rows = p->a_rows > 0 ? p->a_rows : 128;
  a.a_row_offset = p->a_row_offset;
  a.a_base_offset = p->a_base_offset;
  PF_REQUIRE(a.a_rows >= 128 && a.a_rows <= 256 && a.a_row_offset >= 0 && a.a_row_offset + 128 <= a.a_rows,
             "pf_debug_umma: bad a_rows / a_row_offset");
  PF_REQUIRE(!(a.a_from_tmem && a.a_rows != 128), "pf_debug_umma: row offsets need A in shared memory");
  const int smem = ((a.a_rows * 128 + 1023) & ~1023) * (p->k / 64) + p->b_rows * p->b_cols * 2 + 2048;
  PF_REQUIRE(smem <= 220 * 1024, "pf_debug_umma: operands do not fit in shared memory");

  CUtensorMap tm_a, tm_b;
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(p->k), static_cast<uint64_t>(a.a_rows)};
    const uint64_t strides[1] = {static_cast<uint64_t>(p->k) * 2};
    const uint32_t box[2] = {64, static_cast<uint32_t>(a.a_rows)};
    int rc = encode_tensor_map(&tm_a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, p->a, dims, strides, box,
                               CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(p->b_cols), static_cast<uint64_t>(p->b_rows)};
    const uint64_t strides[1] = {static_cast<uint64_t>(p->b_cols) * 2};
    const uint32_t box[2] = {64, static_cast<uint32_t>(p->b_box_rows)};
    int rc = encode_tensor_map(&tm_b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, p->b, dims, strides, box,
                               CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  cudaError_t e = cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  if (e != cudaSuccess) {
    set_error("pf_debug_umma: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    return -2;
  }
  umma_probe_kernel<<<1, 128, smem, stream>>>(tm_a, tm_b, a);
  return check_launch("pf_debug_umma");
}
