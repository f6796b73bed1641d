"""Tensor-level wrappers over the C-ABI (one function per entry point in include/pf_b200.h).

Each wrapper validates dtypes/contiguity, passes raw pointers + the current stream, and raises on a non-zero status.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import (AttnDesc, GemmDesc, UmmaProbe, PF_EPI_GATE_RESID, PF_EPI_GELU_BF16, PF_EPI_QKV_GELU,
                   PF_EPI_QKV_ROPE, PF_EPI_STORE_BF16, PF_EPI_STORE_F32)


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def gemm(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], epilogue: int, *,
         batches: int = 1, rows_per_batch: Optional[int] = None, row_begin: int = 0, row_count: Optional[int] = None,
         out: Optional[torch.Tensor] = None, ldo: Optional[int] = None, out_batch_rows: Optional[int] = None,
         out_row_begin: Optional[int] = None, out_col_begin: int = 0,
         gate: Optional[torch.Tensor] = None, gate_batch_stride: int = 0,
         q_out=None, k_out=None, v_out=None, rope=None, q_norm_w=None, k_norm_w=None, norm_eps: float = 1e-6,
         heads: int = 0, head_dim: int = 0, seq_len: int = 0, n_split: int = 0, kernel_variant: int = 0,
         peer: Optional[dict] = None) -> None:
    """epilogue(A . W^T + bias); see pf_gemm_bf16 in include/pf_b200.h for the addressing rules."""
    assert a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and a.is_cuda and w.is_cuda
    assert a.stride(-1) == 1 and w.is_contiguous()
    n, k = w.shape
    lda = a.stride(-2)
    if rows_per_batch is None:
        rows_per_batch = a.numel() // (a.shape[-1] * batches) if a.is_contiguous() else a.shape[-2]
    if row_count is None:
        row_count = rows_per_batch - row_begin
    d = GemmDesc()
    d.a, d.lda = a.data_ptr(), lda
    d.batches, d.rows_per_batch, d.row_begin, d.row_count = batches, rows_per_batch, row_begin, row_count
    d.w, d.n, d.k = w.data_ptr(), n, k
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.is_contiguous()
    d.bias = _ptr(bias)
    d.epilogue = epilogue
    d.out = _ptr(out)
    if out is not None:
        d.ldo = ldo if ldo is not None else out.stride(-2)
    d.out_batch_rows = out_batch_rows if out_batch_rows is not None else rows_per_batch
    d.out_row_begin = out_row_begin if out_row_begin is not None else row_begin
    d.out_col_begin = out_col_begin
    if gate is not None:
        assert gate.dtype == torch.float32
    d.gate, d.gate_batch_stride = _ptr(gate), gate_batch_stride
    d.q_out, d.k_out, d.v_out = _ptr(q_out), _ptr(k_out), _ptr(v_out)
    for t in (rope, q_norm_w, k_norm_w):
        if t is not None:
            assert t.dtype == torch.float32 and t.is_contiguous()
    d.rope, d.q_norm_w, d.k_norm_w = _ptr(rope), _ptr(q_norm_w), _ptr(k_norm_w)
    d.norm_eps = norm_eps
    d.heads, d.head_dim, d.seq_len, d.n_split = heads, head_dim, seq_len, n_split
    d.kernel_variant = kernel_variant
    if peer is not None:      # sequence parallel: QKV heads stored straight into the owning rank's buffer (pf_b200.h)
        for i, pp in enumerate(peer["peer_ptrs"]):
            d.peer_qkv[i] = pp
        d.peer_count, d.peer_heads = len(peer["peer_ptrs"]), peer["peer_heads"]
        d.peer_seq, d.peer_row0 = peer["peer_seq"], peer["peer_row0"]
    _lib.check(_lib.load().pf_gemm_bf16(C.byref(d), _lib.stream_ptr()), "pf_gemm_bf16")


def linear_bf16(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, gelu: bool = False) -> torch.Tensor:
    """Convenience: y_bf16[M, N] = (gelu)(x[M, K] . w[N, K]^T + bias)."""
    m = x.shape[0]
    out = torch.empty(m, w.shape[0], dtype=torch.bfloat16, device=x.device)
    gemm(x, w, bias, PF_EPI_GELU_BF16 if gelu else PF_EPI_STORE_BF16, rows_per_batch=m, out=out)
    return out


def ln_modulate(x: torch.Tensor, y: torch.Tensor, shift: torch.Tensor, scale: torch.Tensor, mod_batch_stride: int, *,
                batches: int, rows_per_batch: int, row_begin: int, row_count: int, eps: float = 1e-6) -> None:
    assert x.dtype == torch.float32 and y.dtype == torch.bfloat16 and x.is_contiguous() and y.is_contiguous()
    dim = x.shape[-1]
    _lib.check(_lib.load().pf_ln_modulate(x.data_ptr(), y.data_ptr(), batches, rows_per_batch, row_begin, row_count,
                                          dim, shift.data_ptr(), scale.data_ptr(), mod_batch_stride, eps,
                                          _lib.stream_ptr()), "pf_ln_modulate")


def small_linear(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], y: torch.Tensor, *, act_in: int = 0,
                 act_out: int = 0, accumulate: bool = False, round_in_bf16: bool = False) -> None:
    assert x.dtype == torch.float32 and y.dtype == torch.float32 and w.dtype == torch.bfloat16
    assert x.is_contiguous() and w.is_contiguous() and y.is_contiguous()
    m, k = x.shape
    n = w.shape[0]
    assert w.shape[1] == k and y.shape == (m, n)
    _lib.check(_lib.load().pf_small_linear(x.data_ptr(), m, k, w.data_ptr(), _ptr(bias), n, y.data_ptr(), act_in,
                                           act_out, int(accumulate), int(round_in_bf16), _lib.stream_ptr()),
               "pf_small_linear")


def timestep_embedding(t: torch.Tensor, dim: int, round_bf16: bool = True) -> torch.Tensor:
    assert t.dtype == torch.float32 and t.is_contiguous()
    out = torch.empty(t.shape[0], dim, dtype=torch.float32, device=t.device)
    _lib.check(_lib.load().pf_timestep_embedding(t.data_ptr(), t.shape[0], dim, out.data_ptr(), int(round_bf16),
                                                 _lib.stream_ptr()), "pf_timestep_embedding")
    return out


def patchify(latent: torch.Tensor, tokens: torch.Tensor, rows_per_batch: int, tok_begin: int) -> None:
    assert latent.is_contiguous() and latent.dtype in (torch.float32, torch.bfloat16) and tokens.dtype == torch.bfloat16
    b, c, t, h, w = latent.shape
    _lib.check(_lib.load().pf_patchify(latent.data_ptr(), int(latent.dtype == torch.float32), b, c, t, h, w,
                                       tokens.data_ptr(), rows_per_batch, tok_begin, _lib.stream_ptr()), "pf_patchify")


def unpatchify(x: torch.Tensor, rows_per_batch: int, row_begin: int, out: torch.Tensor) -> None:
    assert x.dtype == torch.float32 and x.is_contiguous() and out.is_contiguous()
    b, c, t, h, w = out.shape
    _lib.check(_lib.load().pf_unpatchify(x.data_ptr(), rows_per_batch, row_begin, b, c, t, h, w, out.data_ptr(),
                                         int(out.dtype == torch.float32), _lib.stream_ptr()), "pf_unpatchify")


def cfg_euler_step(v2: torch.Tensor, guidance: float, dsigma: float, x: torch.Tensor, x_out: torch.Tensor) -> None:
    assert v2.dtype == torch.float32 and x.dtype == torch.float32 and x_out.dtype == torch.float32
    n = x.numel()
    assert v2.numel() == 2 * n
    _lib.check(_lib.load().pf_cfg_euler_step(v2.data_ptr(), guidance, dsigma, x.data_ptr(), x_out.data_ptr(), n,
                                             _lib.stream_ptr()), "pf_cfg_euler_step")


def stage_hop(x: torch.Tensor, z: torch.Tensor, alpha: float, beta: float, gamma: float) -> torch.Tensor:
    """x [b, c, t, h, w] (bf16/fp32) -> alpha * nearest_x2(x) + beta * block_noise(z), z iid normal fp32 [b, c, t, 2h, 2w]
    (pf_stage_hop; reference P:729-743 + P:697-703)."""
    import ctypes as C
    assert x.is_cuda and x.is_contiguous() and z.is_contiguous() and z.dtype == torch.float32
    assert x.dtype in (torch.float32, torch.bfloat16)
    b, c, t, h, w = x.shape
    assert tuple(z.shape) == (b, c, t, 2 * h, 2 * w)
    cov = torch.eye(4, dtype=torch.float64) * (1 + gamma) - torch.ones(4, 4, dtype=torch.float64) * gamma
    chol = torch.linalg.cholesky(cov).to(torch.float32).flatten().tolist()
    out = torch.empty(b, c, t, 2 * h, 2 * w, device=x.device, dtype=x.dtype)
    arr = (C.c_float * 16)(*chol)
    _lib.check(_lib.load().pf_stage_hop(x.data_ptr(), int(x.dtype == torch.float32), z.data_ptr(), out.data_ptr(), b * c * t, h, w,
                                        alpha, beta, arr, _lib.stream_ptr()), "pf_stage_hop")
    return out


def attn_build_schedule(seg: torch.Tensor, time: torch.Tensor):
    """seg/time: int32 CPU tensors [batch, seq] -> (schedule int32 CPU [batch, q_tiles, stride], allowed_pairs [batch])."""
    seg = seg.to(torch.int32).contiguous().cpu()
    time = time.to(torch.int32).contiguous().cpu()
    batch, seq = seg.shape
    lib = _lib.load()
    stride = lib.pf_attn_build_schedule(None, None, batch, seq, None, None)
    if stride < 0:
        _lib.check(stride, "pf_attn_build_schedule")
    qt = (seq + 127) // 128
    sched = torch.zeros(batch, qt, stride, dtype=torch.int32)
    pairs = torch.zeros(batch, dtype=torch.int64)
    rc = lib.pf_attn_build_schedule(seg.data_ptr(), time.data_ptr(), batch, seq, sched.data_ptr(), pairs.data_ptr())
    if rc < 0:
        _lib.check(rc, "pf_attn_build_schedule")
    return sched, pairs


class PairSchedule:
    """Schedule of the two-q-tile attention kernel: `sched` int32 [batch, n_pairs, stride] (pf_attn_build_pair_schedule),
    `mask_index` int32 [batch, n_pairs, 2 * stride] and `mask_bits` int32 [blocks, 128, 4] (pf_attn_build_pair_masks).
    `group3` (optional) = the same three tensors for groups of three q tiles (pf_attn_build_group_schedule / _masks), the
    schedule of the three-q-tile kernel."""

    def __init__(self, sched: torch.Tensor, mask_index: torch.Tensor, mask_bits: torch.Tensor,
                 group3: Optional["PairSchedule"] = None):
        self.sched, self.mask_index, self.mask_bits, self.group3 = sched, mask_index, mask_bits, group3

    def to(self, device) -> "PairSchedule":
        bits = self.mask_bits.to(device)
        g3 = self.group3
        if g3 is not None:     # the group schedule indexes the SAME block pool when it was built from this pair schedule
            g3 = PairSchedule(g3.sched.to(device), g3.mask_index.to(device), bits if g3.mask_bits is self.mask_bits else g3.mask_bits.to(device))
        return PairSchedule(self.sched.to(device), self.mask_index.to(device), bits, g3)

    def __getitem__(self, idx) -> "PairSchedule":          # batch slice (block indices are global: the bit pool is shared)
        assert isinstance(idx, slice)
        return PairSchedule(self.sched[idx], self.mask_index[idx], self.mask_bits,
                            None if self.group3 is None else self.group3[idx])


def attn_build_pair_schedule(sched: torch.Tensor, seq: int, seg: torch.Tensor, time: torch.Tensor) -> PairSchedule:
    """Tile schedule (int32 CPU [batch, q_tiles, stride]) + the seg/time ids -> PairSchedule (CPU tensors) of the two-q-tile
    kernel: merged kv lists of adjacent q tiles and the precomputed 128-bit row masks of their partial tiles."""
    sched = sched.to(torch.int32).contiguous().cpu()
    seg = seg.to(torch.int32).contiguous().cpu()
    time = time.to(torch.int32).contiguous().cpu()
    batch, qt, stride = sched.shape
    n_pairs = (qt + 1) // 2
    lib = _lib.load()
    ps = torch.zeros(batch, n_pairs, stride, dtype=torch.int32)
    _lib.check(lib.pf_attn_build_pair_schedule(sched.data_ptr(), batch, seq, stride, ps.data_ptr()), "pf_attn_build_pair_schedule")
    midx = torch.full((batch, n_pairs, 2 * stride), -1, dtype=torch.int32)
    n = lib.pf_attn_build_pair_masks(seg.data_ptr(), time.data_ptr(), ps.data_ptr(), batch, seq, stride, midx.data_ptr(), None, 0)
    if n < 0:
        _lib.check(int(n), "pf_attn_build_pair_masks")
    bits = torch.zeros(max(1, int(n)), 128, 4, dtype=torch.int32)
    n2 = lib.pf_attn_build_pair_masks(seg.data_ptr(), time.data_ptr(), ps.data_ptr(), batch, seq, stride, midx.data_ptr(),
                                      bits.data_ptr(), int(n))
    assert n2 == n
    pso = PairSchedule(ps, midx, bits)
    pso.group3 = attn_build_group_schedule(sched, seq, seg, time, 3, share=pso)
    return pso


def attn_build_group_schedule(sched: torch.Tensor, seq: int, seg: torch.Tensor, time: torch.Tensor, group: int = 3,
                              share: Optional[PairSchedule] = None) -> PairSchedule:
    """The pair schedule generalised to groups of `group` q tiles (the three-q-tile kernel): sched int32 [batch, n_groups,
    stride] with entries (kv_tile << 8) | 2 flag bits per tile, mask_index [batch, n_groups, group * stride], mask_bits.
    `share` = the pair schedule of the same tile schedule: its block pool is reused (a block depends on (q tile, kv tile) only),
    no bits are built and `mask_bits` IS `share.mask_bits`."""
    sched = sched.to(torch.int32).contiguous().cpu()
    seg = seg.to(torch.int32).contiguous().cpu()
    time = time.to(torch.int32).contiguous().cpu()
    batch, qt, stride = sched.shape
    n_groups = (qt + group - 1) // group
    lib = _lib.load()
    gs = torch.zeros(batch, n_groups, stride, dtype=torch.int32)
    _lib.check(lib.pf_attn_build_group_schedule(sched.data_ptr(), batch, seq, stride, group, gs.data_ptr()), "pf_attn_build_group_schedule")
    midx = torch.full((batch, n_groups, group * stride), -1, dtype=torch.int32)
    if share is not None:
        assert share.sched.shape[-1] == stride and share.sched.is_contiguous() and share.mask_index.is_contiguous()
        n = lib.pf_attn_build_group_masks(seg.data_ptr(), time.data_ptr(), gs.data_ptr(), batch, seq, stride, group, midx.data_ptr(),
                                          None, 0, share.sched.data_ptr(), share.mask_index.data_ptr())
        if n < 0:
            _lib.check(int(n), "pf_attn_build_group_masks")
        assert n <= share.mask_bits.shape[0]
        return PairSchedule(gs, midx, share.mask_bits)
    n = lib.pf_attn_build_group_masks(seg.data_ptr(), time.data_ptr(), gs.data_ptr(), batch, seq, stride, group, midx.data_ptr(), None, 0,
                                      None, None)
    if n < 0:
        _lib.check(int(n), "pf_attn_build_group_masks")
    bits = torch.zeros(max(1, int(n)), 128, 4, dtype=torch.int32)
    n2 = lib.pf_attn_build_group_masks(seg.data_ptr(), time.data_ptr(), gs.data_ptr(), batch, seq, stride, group, midx.data_ptr(),
                                       bits.data_ptr(), int(n), None, None)
    assert n2 == n
    return PairSchedule(gs, midx, bits)


def attn_fwd(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, out: torch.Tensor, seg: torch.Tensor,
             time: torch.Tensor, sched: torch.Tensor, scale: float, variant: int = 0, q_row_begin: int = 0,
             pair_sched: Optional[PairSchedule] = None, ldo: Optional[int] = None, peer: Optional[dict] = None) -> None:
    """q,k,v bf16 [B,H,S,64]; out bf16 [B,S,*] (row stride = out.stride(1)); seg/time/sched int32 on device.
    Only q rows >= q_row_begin (multiple of 128) are computed; other rows of `out` are left untouched."""
    assert q.dtype == torch.bfloat16 and q.is_contiguous() and k.is_contiguous() and v.is_contiguous()
    b, h, s, hd = q.shape
    d = AttnDesc()
    d.q, d.k, d.v = q.data_ptr(), k.data_ptr(), v.data_ptr()
    if out is not None:
        d.out = out.data_ptr()
        d.ldo = out.stride(-2)
    if ldo is not None:
        d.ldo = ldo
    if peer is not None:      # sequence parallel: output rows stored straight into the owning rank's buffer (pf_b200.h)
        for i, pp in enumerate(peer["peer_ptrs"]):
            d.peer_out[i] = pp
        d.peer_count, d.peer_chunk_rows, d.peer_col_begin = len(peer["peer_ptrs"]), peer["peer_chunk_rows"], peer["peer_col_begin"]
    d.batch, d.heads, d.seq, d.head_dim = b, h, s, hd
    d.scale = scale
    d.seg, d.time, d.tile_sched = seg.data_ptr(), time.data_ptr(), sched.data_ptr()
    d.sched_stride = sched.shape[-1]
    d.variant = variant
    d.q_row_begin = q_row_begin
    if pair_sched is not None:
        assert pair_sched.sched.dtype == torch.int32 and pair_sched.sched.shape[-1] == sched.shape[-1]
        assert pair_sched.sched.is_contiguous() and pair_sched.mask_index.is_contiguous() and pair_sched.mask_bits.is_contiguous()
        d.pair_sched = pair_sched.sched.data_ptr()
        d.pair_mask_index = pair_sched.mask_index.data_ptr()
        d.pair_mask_bits = pair_sched.mask_bits.data_ptr()
        g3 = pair_sched.group3
        if g3 is not None:
            assert g3.sched.is_contiguous() and g3.mask_index.is_contiguous() and g3.mask_bits.is_contiguous()
            assert g3.sched.shape[-1] == sched.shape[-1]
            d.group_sched, d.group_mask_index, d.group_mask_bits = g3.sched.data_ptr(), g3.mask_index.data_ptr(), g3.mask_bits.data_ptr()
    _lib.check(_lib.load().pf_attn_fwd_masked(C.byref(d), _lib.stream_ptr()), "pf_attn_fwd_masked")


def debug_umma(a: torch.Tensor, b: torch.Tensor, n: int, k: int, *, b_box_rows: int, b_mn_major: int, b_lbo: int,
               b_sbo: int, b_k_step_bytes: int, b_kblock_bytes: int, a_from_tmem: int, a_row_offset: int = 0,
               a_base_offset: int = 0) -> torch.Tensor:
    d = torch.zeros(128, n, dtype=torch.float32, device=a.device)
    p = UmmaProbe()
    p.a, p.b, p.d = a.data_ptr(), b.data_ptr(), d.data_ptr()
    p.n, p.k = n, k
    p.b_rows, p.b_cols = b.shape
    p.b_box_rows, p.b_mn_major = b_box_rows, b_mn_major
    p.b_lbo, p.b_sbo, p.b_k_step_bytes, p.b_kblock_bytes = b_lbo, b_sbo, b_k_step_bytes, b_kblock_bytes
    p.a_rows, p.a_row_offset, p.a_base_offset = a.shape[0], a_row_offset, a_base_offset
    p.a_from_tmem = a_from_tmem
    _lib.check(_lib.load().pf_debug_umma(C.byref(p), _lib.stream_ptr()), "pf_debug_umma")
    return d
