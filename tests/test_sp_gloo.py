"""Multi-rank host logic on CPU (gloo, 4 processes = CFG 2 x SP 2): the Ulysses exchanges of pyramid_flow_b200/sp.py
reproduce the single-process tensors, and the layout arithmetic (chunks, padded heads) is consistent."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pyramid_flow_b200 import sp as SP


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lay = SP.make_layout()
        assert (lay.cfg_ways, lay.sp) == (2, world // 2) and lay.rank == rank
        heads, hd, seq = 6, 8, 40
        hp = SP.padded_heads(heads, lay.sp)
        g = torch.Generator().manual_seed(100 + lay.cfg_rank)      # same full tensors on the ranks of one CFG branch
        q_full = torch.randn(hp, seq, hd, generator=g)
        c0, c1 = SP.chunk_bounds(seq, lay.sp, lay.sp_rank)
        mine = SP.heads_to_sequence(q_full[:, c0:c1].contiguous(), lay)
        hg = hp // lay.sp
        assert torch.equal(mine, q_full[lay.sp_rank * hg:(lay.sp_rank + 1) * hg])       # my head group, whole sequence
        k_full, v_full = q_full * 2, q_full + 1
        a, b_, c_ = SP.heads_to_sequence_qkv(*(t[:, c0:c1].contiguous() for t in (q_full, k_full, v_full)), lay)
        sl_h = slice(lay.sp_rank * hg, (lay.sp_rank + 1) * hg)
        assert torch.equal(a, q_full[sl_h]) and torch.equal(b_, k_full[sl_h]) and torch.equal(c_, v_full[sl_h])
        # inverse exchange on the token-major attention output [S, Hg*hd]
        o_full = torch.randn(seq, hp * hd, generator=g)                                  # all heads, token major
        o_mine = o_full[:, lay.sp_rank * hg * hd:(lay.sp_rank + 1) * hg * hd].contiguous()
        back = SP.sequence_to_heads(o_mine, lay)
        assert torch.equal(back, o_full[c0:c1])
        # CFG combine path: all_gather over the cfg group orders [uncond ; cond]
        v = torch.full((1, 3), float(lay.cfg_rank))
        full = torch.empty(2, 3)
        dist.all_gather_into_tensor(full, v, group=lay.cfg_group)
        assert full[:, 0].tolist() == [0.0, 1.0]
        ret[rank] = 1
    finally:
        dist.destroy_process_group()


def test_ulysses_exchange_world4():
    world = 4
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert len(ret) == world


def test_layout_arithmetic():
    assert SP.padded_heads(30, 1) == 30 and SP.padded_heads(30, 2) == 30 and SP.padded_heads(30, 4) == 32
    for sp in (1, 2, 4):
        b = [SP.chunk_bounds(15488, sp, r) for r in range(sp)]
        assert b[0][0] == 0 and b[-1][1] == 15488 and all(b[i][1] == b[i + 1][0] for i in range(sp - 1))
    lay = SP.make_layout(8, 5, create_groups=False)
    assert (lay.cfg_ways, lay.sp, lay.cfg_rank, lay.sp_rank) == (2, 4, 1, 1)
