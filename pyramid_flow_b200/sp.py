"""Sequence / CFG parallel layout of the DiT step over N GPUs (one process per GPU, `torch.distributed`, NCCL on NVLink).

Follows the reference's sequence-parallel layout (trainer_misc/sp_utils.py:21-47 groups; Ulysses-style head<->sequence
all-to-all at the attention boundary, flux_modules/modeling_flux_block.py:266-325, 519-565 via trainer_misc/communicate.py)
with two changes the reference cannot make:
  * the CFG pair is split first (uncond / cond on separate halves of the world: no traffic until the velocity combine),
    so every rank runs batch 1 and the reference's `B % sp == 0` transposition trick (F:471-485) is not needed;
  * 30 heads do not divide by 4 or 8: heads are zero-padded to the next multiple of the SP degree (32 at sp=4, a 6.7 %
    attention overhead) instead of restricting miniFLUX to sp=2.
The joint sequence [text ; clips] is cut into `sp` contiguous chunks (all S of the 768p schedule are multiples of 8).

world = cfg_ways (2 if world >= 2 else 1) x sp;   rank -> (cfg_rank = rank // sp, sp_rank = rank % sp).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import torch
import torch.distributed as dist


@dataclass
class ParallelLayout:
    world: int
    rank: int
    cfg_ways: int
    sp: int
    cfg_rank: int
    sp_rank: int
    sp_group: Optional[object] = None    # ranks sharing a CFG branch
    cfg_group: Optional[object] = None   # the two ranks holding the same token chunk of the two branches

    @property
    def enabled(self) -> bool:
        return self.world > 1


def make_layout(world: Optional[int] = None, rank: Optional[int] = None, create_groups: bool = True) -> ParallelLayout:
    if world is None:
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        rank = dist.get_rank() if world > 1 else 0
    if world == 1:
        return ParallelLayout(1, 0, 1, 1, 0, 0)
    assert world % 2 == 0, "world size must be even (CFG pair is split first)"
    cfg_ways, sp = 2, world // 2
    lay = ParallelLayout(world, rank, cfg_ways, sp, rank // sp, rank % sp)
    if create_groups:
        # every rank must create every group (torch.distributed contract)
        for c in range(cfg_ways):
            g = dist.new_group(list(range(c * sp, (c + 1) * sp)))
            if c == lay.cfg_rank:
                lay.sp_group = g
        for s in range(sp):
            g = dist.new_group([s, sp + s])
            if s == lay.sp_rank:
                lay.cfg_group = g
    return lay


def padded_heads(heads: int, sp: int) -> int:
    return (heads + sp - 1) // sp * sp


def chunk_bounds(seq: int, sp: int, sp_rank: int) -> Tuple[int, int]:
    assert seq % sp == 0, f"sequence length {seq} must be divisible by the SP degree {sp}"
    n = seq // sp
    return sp_rank * n, (sp_rank + 1) * n


def heads_to_sequence(x: torch.Tensor, lay: ParallelLayout) -> torch.Tensor:
    """Attention-boundary exchange #1 (reference B:285,295): x [Hp, S_local, hd] holds ALL (padded) heads of this rank's
    token chunk; returns [Hp/sp, S, hd] = this rank's head group over the WHOLE sequence.  One all_to_all_single."""
    hp, s_l, hd = x.shape
    hg = hp // lay.sp
    recv = torch.empty(lay.sp, hg, s_l, hd, dtype=x.dtype, device=x.device)
    dist.all_to_all_single(recv, x.view(lay.sp, hg, s_l, hd), group=lay.sp_group)
    # recv[src] = head group `sp_rank` of rank src's chunk -> concatenate the chunks along the sequence
    return recv.permute(1, 0, 2, 3).reshape(hg, lay.sp * s_l, hd)


def heads_to_sequence_qkv_begin(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, lay: ParallelLayout):
    """Start the three exchanges of one attention back to back (async on NCCL's stream); independent work — the single
    block's proj_mlp GEMM — can be launched on the compute stream before `heads_to_sequence_qkv_end` waits."""
    hp, s_l, hd = q.shape
    hg = hp // lay.sp
    recvs, works = [], []
    for x in (q, k, v):
        r = torch.empty(lay.sp, hg, s_l, hd, dtype=x.dtype, device=x.device)
        works.append(dist.all_to_all_single(r, x.view(lay.sp, hg, s_l, hd), group=lay.sp_group, async_op=True))
        recvs.append(r)
    return works, recvs, (hg, lay.sp * s_l, hd)


def heads_to_sequence_qkv_end(handle):
    works, recvs, shape = handle
    for w in works:
        w.wait()
    return tuple(r.permute(1, 0, 2, 3).reshape(shape) for r in recvs)


def heads_to_sequence_qkv(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, lay: ParallelLayout):
    return heads_to_sequence_qkv_end(heads_to_sequence_qkv_begin(q, k, v, lay))


def sequence_to_heads(o: torch.Tensor, lay: ParallelLayout) -> torch.Tensor:
    """Attention-boundary exchange #2 (reference B:314,321): o [S, Hg*hd] (this rank's head group, whole sequence, token
    major); returns [S_local, Hp*hd] = all heads for this rank's token chunk."""
    s, w = o.shape
    s_l = s // lay.sp
    recv = torch.empty(lay.sp, s_l, w, dtype=o.dtype, device=o.device)
    dist.all_to_all_single(recv, o.view(lay.sp, s_l, w), group=lay.sp_group)
    # recv[src] = head group src for my chunk -> heads concatenated along the feature axis
    return recv.permute(1, 0, 2).reshape(s_l, lay.sp * w)
