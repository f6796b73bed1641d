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


# ---- VAE context-parallel ring (vae.py): the host-side schedule + halo plumbing with a stand-in for the kernels -------------
class _FakeConv:
    kt = 3
    cache = None


def _fake_vae():
    """B200CausalVAE with the CUDA chunk decoder replaced by a causal 3-tap FIR over time (uses the real `_halo`) followed
    by the x8 temporal up-sampling with the first-frame rule, so only the schedule / halo / gather logic is exercised."""
    from pyramid_flow_b200.vae import B200CausalVAE, VaeConfigB200
    vae = B200CausalVAE.__new__(B200CausalVAE)
    torch.nn.Module.__init__(vae)
    vae.cfg = VaeConfigB200(spatial_up_sample=(False, False, False, False))
    vae.register_buffer("_anchor", torch.zeros(1))
    vae._cp, vae._cp_ctx, vae.cp_frames_per_round = None, None, 4
    vae.convs = {"fir": _FakeConv()}

    def decode_chunk(z, first):
        x = z[0].permute(1, 2, 3, 0).contiguous().float()[..., :3]          # [T, h, w, 3]
        buf = torch.zeros(x.shape[0] + 2, *x.shape[1:])
        buf[2:] = x
        vae._halo(vae.convs["fir"], buf, first)
        y = buf[:-2] + 2.0 * buf[1:-1] + 3.0 * buf[2:]                         # causal: frame t sees t-2, t-1, t
        y = y.repeat_interleave(8, dim=0)
        return y[7:] if first else y

    vae._decode_chunk = decode_chunk
    return vae


def _cp_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        for n_frames, c in [(5, 4), (12, 4), (31, 4), (23, 2), (9, 3)]:
            g = torch.Generator().manual_seed(n_frames)
            z = torch.randn(1, 16, n_frames, 3, 4, generator=g)
            vae = _fake_vae()
            ref = vae._decode_sample(z, 2)                                  # single process, chunked with the feature cache
            assert ref.shape[0] == 8 * (n_frames - 1) + 1
            vae.cp_frames_per_round = c
            vae.set_context_parallel(None)
            out = vae._decode_sample(z, 2)
            assert vae._cp is not None and torch.equal(out, ref), (n_frames, c, (out - ref).abs().max())
        ret[rank] = 1
    finally:
        dist.destroy_process_group()


def test_vae_context_parallel_ring_world2_and_3():
    for world in (2, 3):
        ctx = mp.get_context("spawn")
        ret = ctx.Manager().dict()
        port = _free_port()
        procs = [ctx.Process(target=_cp_worker, args=(r, world, port, ret)) for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(timeout=120)
        assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
        assert sorted(ret.keys()) == list(range(world))
