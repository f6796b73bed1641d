"""Hardware bring-up checks, run on the GPU box: `python tools/gpu_check.py [group ...]`.

Each group runs in its own subprocess under a timeout so a hung kernel cannot take the others down.
Results go to stdout and gpurun_out/gpu_check_<group>.log.  These are bring-up diagnostics; the parity tests
proper live in tests/ (-m gpu).
"""
from __future__ import annotations

import os
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

GROUPS = ["probe", "probe_rowoff", "conv_kwreuse", "probe_ts", "gemm", "epilogue", "elementwise", "attn", "gemm_perf", "gemm_epi_perf", "gemm_qkv_perf", "attn_perf", "attn_trace", "attn_cta_trace", "vae_perf"]


def _rel_err(a, b):
    import torch
    a = a.float()
    b = b.float()
    return ((a - b).abs().max() / (b.abs().max() + 1e-20)).item()


# ----------------------------------------------------------------------------------------------------------------
def group_probe():
    import torch
    from pyramid_flow_b200 import ops
    torch.manual_seed(0)
    dev = "cuda"

    def run(name, a, b_glob, ref, **kw):
        try:
            d = ops.debug_umma(a, b_glob, **kw)
            torch.cuda.synchronize()
            err = _rel_err(d, ref)
            print(f"[probe] {name}: rel_err={err:.3e} {'OK' if err < 2e-2 else 'MISMATCH'}", flush=True)
            return err
        except Exception as e:  # noqa: BLE001
            print(f"[probe] {name}: EXC {e}", flush=True)
            return None

    for k in (64, 128):
        for n in (64, 128, 256):
            a = torch.randn(128, k, device=dev).bfloat16()
            bt = torch.randn(n, k, device=dev).bfloat16()  # [N, K] K-major
            ref = a.float() @ bt.float().t()
            run(f"kmajor n={n} k={k}", a, bt, ref, n=n, k=k, b_box_rows=n, b_mn_major=0, b_lbo=16, b_sbo=1024,
                b_k_step_bytes=32, b_kblock_bytes=n * 128, a_from_tmem=0)
    # MN-major B: global [K, N] row-major (N contiguous) — V as stored [kv, hd]; boxes [k rows x 64 cols], one per
    # 64-wide n block, sequential in smem => LBO (n-atom stride) = k*128 bytes, SBO (8-row k group stride) = 1024.
    for k in (64, 128):
        for n in (64, 128):
            a = torch.randn(128, k, device=dev).bfloat16()
            b = torch.randn(k, n, device=dev).bfloat16()  # [K, N]
            ref = a.float() @ b.float()
            run(f"mnmajor n={n} k={k}", a, b, ref, n=n, k=k, b_box_rows=k, b_mn_major=1, b_lbo=k * 128, b_sbo=1024,
                b_k_step_bytes=2048, b_kblock_bytes=8192, a_from_tmem=0)


def group_probe_rowoff():
    """A descriptor advanced by whole rows inside the 128B-swizzle atom (the conv's kw-tap reuse of one haloed patch):
    which value of the base-offset field makes D = A[off:off+128] . B^T ?"""
    import torch
    from pyramid_flow_b200 import ops
    torch.manual_seed(0)
    dev = "cuda"
    n = 128
    for k in (64, 128):
        a = torch.randn(136, k, device=dev).bfloat16()
        bt = torch.randn(n, k, device=dev).bfloat16()
        for off in (0, 1, 2, 5, 8):
            ref = a[off:off + 128].float() @ bt.float().t()
            for bo in sorted({0, off & 7}):
                try:
                    d = ops.debug_umma(a, bt, n=n, k=k, b_box_rows=n, b_mn_major=0, b_lbo=16, b_sbo=1024, b_k_step_bytes=32,
                                       b_kblock_bytes=n * 128, a_from_tmem=0, a_row_offset=off, a_base_offset=bo)
                    torch.cuda.synchronize()
                    err = _rel_err(d, ref)
                    print(f"[probe_rowoff] k={k} row offset {off}, base_offset field {bo}: rel_err={err:.3e} {'OK' if err < 2e-2 else 'MISMATCH'}", flush=True)
                except Exception as e:  # noqa: BLE001
                    print(f"[probe_rowoff] k={k} off={off} bo={bo}: EXC {e}", flush=True)


def group_conv_kwreuse():
    """kw-tap reuse conv kernel (kernel_variant 3) vs the default 2-CTA kernel and F.conv3d: correctness + time."""
    import torch
    import torch.nn.functional as F
    from pyramid_flow_b200.vae import B200CausalVAE, _Conv
    torch.manual_seed(0)
    dev = torch.device("cuda:0")
    holder = B200CausalVAE.__new__(B200CausalVAE)
    for (ci, co, t, h, w, check) in [(64, 128, 2, 96, 160, True), (128, 256, 2, 40, 300, True), (128, 128, 2, 768, 1280, False),
                                     (256, 256, 2, 384, 640, False)]:
        wt = (torch.randn(co, ci, 3, 3, 3) * (ci * 27) ** -0.5).bfloat16().float()
        bias = torch.randn(co) * 0.1
        cv = _Conv({"c.conv.weight": wt, "c.conv.bias": bias}, "c", dev)
        x = torch.randn(t, h, w, ci, device=dev).bfloat16()
        xin = torch.zeros(t + 2, h, w, ci, device=dev, dtype=torch.bfloat16)
        xin[2:] = x
        ref = None
        if check:
            xr = F.pad(x.permute(3, 0, 1, 2)[None].float(), (1, 1, 1, 1, 2, 0))
            ref = F.conv3d(xr, wt.to(dev), bias.to(dev))[0].permute(1, 2, 3, 0)
        fl = 2.0 * 27 * ci * co * t * h * w / 1e9
        for kv in (2, 3):     # 2 = 2-CTA one box per tap, 3 = 2-CTA kw-tap reuse (pf_conv3d_desc.kernel_variant)
            out = torch.zeros(t, h, w, co, device=dev, dtype=torch.bfloat16)
            try:
                ms = _time_cuda(lambda: B200CausalVAE._conv(holder, cv, xin, t, h, w, out=out, kernel_variant=kv), iters=5, warm=2)
                if ref is not None:
                    err = (out.float() - ref).abs().max().item()
                else:   # big shapes: sampled output voxels against an fp32 reference of those voxels only
                    err = _conv_sampled_err(x, wt.to(dev), bias.to(dev), out)
                print(f"[conv_kwreuse] {ci}->{co} on {t}x{h}x{w}: kernel_variant={kv}: {ms:.3f} ms = {fl/ms:.0f} TF/s, max_abs_err {err:.3e}", flush=True)
            except Exception as e:  # noqa: BLE001
                print(f"[conv_kwreuse] {ci}->{co} kernel_variant={kv}: EXC {e}", flush=True)


def _conv_sampled_err(x, wt, bias, out, n=4096):
    """max |out - conv(x)| over n random output voxels (x [T,H,W,Cin] channels-last, causal 3x3x3, zero spatial pad)."""
    import torch
    t, h, w, ci = x.shape
    g = torch.Generator(device=x.device).manual_seed(0)
    ts = torch.randint(0, t, (n,), device=x.device, generator=g)
    hs = torch.randint(0, h, (n,), device=x.device, generator=g)
    ws_ = torch.randint(0, w, (n,), device=x.device, generator=g)
    hs[: n // 8] = torch.where(torch.arange(n // 8, device=x.device) % 2 == 0, 0, h - 1)     # borders
    ws_[n // 8: n // 4] = torch.where(torch.arange(n // 8, device=x.device) % 2 == 0, 0, w - 1)
    xp = torch.nn.functional.pad(x.float(), (0, 0, 1, 1, 1, 1, 2, 0))                       # [T+2, H+2, W+2, C]
    acc = bias.float()[None].repeat(n, 1)
    for dt in range(3):
        for dh in range(3):
            for dw in range(3):
                patch = xp[ts + dt, hs + dh, ws_ + dw]                                       # [n, Cin]
                acc += patch @ wt[:, :, dt, dh, dw].float().t()
    return (out[ts, hs, ws_].float() - acc).abs().max().item()


def group_probe_ts():
    import torch
    from pyramid_flow_b200 import ops
    torch.manual_seed(0)
    dev = "cuda"

    def run(name, a, b_glob, ref, **kw):
        d = ops.debug_umma(a, b_glob, **kw)
        torch.cuda.synchronize()
        err = _rel_err(d, ref)
        print(f"[probe] {name}: rel_err={err:.3e} {'OK' if err < 2e-2 else 'MISMATCH'}", flush=True)

    # A from TMEM (packed bf16 pairs), B K-major
    for k in (64, 128):
        for n in (64, 128):
            a = torch.randn(128, k, device=dev).bfloat16()
            bt = torch.randn(n, k, device=dev).bfloat16()
            ref = a.float() @ bt.float().t()
            run(f"a_tmem n={n} k={k}", a, bt, ref, n=n, k=k, b_box_rows=n, b_mn_major=0, b_lbo=16, b_sbo=1024,
                b_k_step_bytes=32, b_kblock_bytes=n * 128, a_from_tmem=1)
    # A from TMEM + B MN-major (the P.V configuration)
    k, n = 128, 64
    a = torch.randn(128, k, device=dev).bfloat16()
    b = torch.randn(k, n, device=dev).bfloat16()
    ref = a.float() @ b.float()
    run("a_tmem + mnmajor n=64 k=128", a, b, ref, n=n, k=k, b_box_rows=k, b_mn_major=1, b_lbo=k * 128, b_sbo=1024,
        b_k_step_bytes=2048, b_kblock_bytes=8192, a_from_tmem=1)


def group_gemm():
    import torch
    from pyramid_flow_b200 import ops
    torch.manual_seed(0)
    dev = "cuda"
    for (m, n, k) in [(128, 64, 64), (128, 256, 64), (256, 256, 128), (300, 1920, 1920), (1000, 192, 512),
                      (4096, 7680, 1920), (777, 128, 4096), (2048, 1920, 9600)]:
        x = (torch.randn(m, k, device=dev) * 0.5).bfloat16()
        w = (torch.randn(n, k, device=dev) * 0.05).bfloat16()
        bias = torch.randn(n, device=dev)
        try:
            y = ops.linear_bf16(x, w, bias)
            torch.cuda.synchronize()
            ref = x.float() @ w.float().t() + bias
            err = _rel_err(y, ref)
            print(f"[gemm] m={m} n={n} k={k}: rel_err={err:.3e} {'OK' if err < 1e-2 else 'MISMATCH'}", flush=True)
            if err >= 1e-2:
                diff = (y.float() - ref).abs()
                bad = (diff > 0.05 * ref.abs().max()).nonzero()
                print(f"        bad elements: {bad.shape[0]} first {bad[:8].tolist()}", flush=True)
        except Exception as e:  # noqa: BLE001
            print(f"[gemm] m={m} n={n} k={k}: EXC {e}", flush=True)


def group_epilogue():
    import torch
    import torch.nn.functional as F
    from pyramid_flow_b200 import ops
    from pyramid_flow_b200._lib import (PF_EPI_GATE_RESID, PF_EPI_GELU_BF16, PF_EPI_QKV_GELU, PF_EPI_QKV_ROPE,
                                        PF_EPI_STORE_F32)
    torch.manual_seed(1)
    dev = "cuda"
    B, S, D, H = 2, 300, 384, 6
    T0 = 40  # "text" rows
    x = (torch.randn(B, S, D, device=dev) * 0.5).bfloat16()
    # ---- GELU, batched row range
    w = (torch.randn(4 * D, D, device=dev) * 0.05).bfloat16()
    bias = torch.randn(4 * D, device=dev) * 0.1
    out = torch.zeros(B, S, 4 * D, device=dev, dtype=torch.bfloat16)
    ops.gemm(x, w, bias, PF_EPI_GELU_BF16, batches=B, rows_per_batch=S, row_begin=T0, row_count=S - T0, out=out)
    torch.cuda.synchronize()
    ref = F.gelu(x[:, T0:].float() @ w.float().t() + bias, approximate="tanh")
    print(f"[epi] gelu range: rel_err={_rel_err(out[:, T0:], ref):.3e}; untouched rows zero: {bool((out[:, :T0] == 0).all())}", flush=True)
    # ---- STORE_F32
    w2 = (torch.randn(D, D, device=dev) * 0.05).bfloat16()
    b2 = torch.randn(D, device=dev) * 0.1
    o32 = torch.zeros(B, S, D, device=dev)
    ops.gemm(x, w2, b2, PF_EPI_STORE_F32, batches=B, rows_per_batch=S, row_begin=0, row_count=T0, out=o32)
    torch.cuda.synchronize()
    ref = x[:, :T0].float() @ w2.float().t() + b2
    print(f"[epi] store_f32: rel_err={_rel_err(o32[:, :T0], ref):.3e}; rest zero: {bool((o32[:, T0:] == 0).all())}", flush=True)
    # ---- GATE_RESID
    resid = torch.randn(B, S, D, device=dev)
    resid0 = resid.clone()
    gate = torch.randn(B, 3 * D, device=dev)
    ops.gemm(x, w2, b2, PF_EPI_GATE_RESID, batches=B, rows_per_batch=S, row_begin=T0, row_count=S - T0, out=resid,
             gate=gate[:, D:], gate_batch_stride=3 * D)
    torch.cuda.synchronize()
    ref = resid0[:, T0:] + gate[:, None, D:2 * D] * (x[:, T0:].float() @ w2.float().t() + b2)
    print(f"[epi] gate_resid: rel_err={_rel_err(resid[:, T0:], ref):.3e}; text rows untouched: {bool((resid[:, :T0] == resid0[:, :T0]).all())}", flush=True)
    # ---- QKV_ROPE (+ QKV_GELU)
    hd = 64
    wq = (torch.randn(3 * D, D, device=dev) * 0.05).bfloat16()
    bq = torch.randn(3 * D, device=dev) * 0.1
    qn = 1 + 0.1 * torch.randn(hd, device=dev)
    kn = 1 + 0.1 * torch.randn(hd, device=dev)
    ang = torch.randn(S, hd // 2, device=dev)
    rope = torch.stack([ang.cos(), ang.sin()], dim=-1).contiguous()  # [S, 32, 2]

    def ref_qkv(xr, pos0):
        y = xr.float() @ wq.float().t() + bq
        q, k, v = y.chunk(3, dim=-1)
        n = xr.shape[1]

        def nr(t, wn):
            t = t.view(B, n, H, hd)
            t = t * torch.rsqrt(t.pow(2).mean(-1, keepdim=True) + 1e-6) * wn
            c = rope[pos0:pos0 + n, :, 0][None, :, None, :]
            s = rope[pos0:pos0 + n, :, 1][None, :, None, :]
            t2 = t.view(B, n, H, hd // 2, 2)
            o = torch.stack([c * t2[..., 0] - s * t2[..., 1], s * t2[..., 0] + c * t2[..., 1]], dim=-1)
            return o.view(B, n, H, hd).transpose(1, 2)
        return nr(q, qn), nr(k, kn), v.view(B, n, H, hd).transpose(1, 2)

    qo = torch.zeros(B, H, S, hd, device=dev, dtype=torch.bfloat16)
    ko = torch.zeros_like(qo)
    vo = torch.zeros_like(qo)
    ops.gemm(x, wq, bq, PF_EPI_QKV_ROPE, batches=B, rows_per_batch=S, row_begin=T0, row_count=S - T0,
             q_out=qo, k_out=ko, v_out=vo, rope=rope, q_norm_w=qn, k_norm_w=kn, heads=H, head_dim=hd, seq_len=S)
    torch.cuda.synchronize()
    rq, rk, rv = ref_qkv(x[:, T0:], T0)
    print(f"[epi] qkv_rope: q={_rel_err(qo[:, :, T0:], rq):.3e} k={_rel_err(ko[:, :, T0:], rk):.3e} v={_rel_err(vo[:, :, T0:], rv):.3e}; text rows zero: {bool((qo[:, :, :T0] == 0).all())}", flush=True)
    # fused q|k|v|mlp
    wm = (torch.randn(4 * D, D, device=dev) * 0.05).bfloat16()
    bm = torch.randn(4 * D, device=dev) * 0.1
    wcat = torch.cat([wq, wm], 0).contiguous()
    bcat = torch.cat([bq, bm], 0).contiguous()
    cat = torch.zeros(B, S, 5 * D, device=dev, dtype=torch.bfloat16)
    qo.zero_(); ko.zero_(); vo.zero_()
    ops.gemm(x, wcat, bcat, PF_EPI_QKV_GELU, batches=B, rows_per_batch=S, row_begin=0, row_count=S, out=cat,
             out_col_begin=D, q_out=qo, k_out=ko, v_out=vo, rope=rope, q_norm_w=qn, k_norm_w=kn, heads=H, head_dim=hd,
             seq_len=S, n_split=3 * D)
    torch.cuda.synchronize()
    rq, rk, rv = ref_qkv(x, 0)
    rm = F.gelu(x.float() @ wm.float().t() + bm, approximate="tanh")
    print(f"[epi] qkv_gelu: q={_rel_err(qo, rq):.3e} k={_rel_err(ko, rk):.3e} v={_rel_err(vo, rv):.3e} mlp={_rel_err(cat[..., D:], rm):.3e}; attn cols zero: {bool((cat[..., :D] == 0).all())}", flush=True)


def group_elementwise():
    import torch
    import torch.nn.functional as F
    from pyramid_flow_b200 import ops
    torch.manual_seed(2)
    dev = "cuda"
    B, S, D = 2, 333, 1920
    x = torch.randn(B, S, D, device=dev) * 2 + 0.3
    mod = torch.randn(B, 6 * D, device=dev) * 0.3
    y = torch.zeros(B, S, D, device=dev, dtype=torch.bfloat16)
    ops.ln_modulate(x, y, mod[:, 0:], mod[:, D:], 6 * D, batches=B, rows_per_batch=S, row_begin=77, row_count=S - 77)
    torch.cuda.synchronize()
    ref = F.layer_norm(x[:, 77:], (D,), eps=1e-6) * (1 + mod[:, None, D:2 * D]) + mod[:, None, :D]
    print(f"[elt] ln_modulate: rel_err={_rel_err(y[:, 77:], ref):.3e}; skipped rows zero: {bool((y[:, :77] == 0).all())}", flush=True)
    # small linear
    xm = torch.randn(2, 1920, device=dev)
    w = (torch.randn(5000, 1920, device=dev) * 0.05).bfloat16()
    b = torch.randn(5000, device=dev)
    yo = torch.zeros(2, 5000, device=dev)
    ops.small_linear(xm, w, b, yo, act_in=1)
    torch.cuda.synchronize()
    ref = F.silu(xm) @ w.float().t() + b
    print(f"[elt] small_linear silu-in: rel_err={_rel_err(yo, ref):.3e}", flush=True)
    ops.small_linear(xm, w, b, yo, act_out=1, accumulate=True)
    torch.cuda.synchronize()
    ref2 = ref + F.silu(xm @ w.float().t() + b)
    print(f"[elt] small_linear accumulate: rel_err={_rel_err(yo, ref2):.3e}", flush=True)
    # timestep embedding
    t = torch.tensor([972.0, 3.5], device=dev)
    e = ops.timestep_embedding(t, 256, round_bf16=False)
    half = 128
    freq = torch.exp(-torch.log(torch.tensor(10000.0)) * torch.arange(half, device=dev).float() / half)
    arg = t[:, None] * freq[None]
    ref = torch.cat([arg.cos(), arg.sin()], -1)
    print(f"[elt] timestep_embedding: max_abs={float((e - ref).abs().max()):.3e}", flush=True)
    # patchify / unpatchify
    lat = torch.randn(2, 16, 2, 8, 12, device=dev).bfloat16()
    L = 2 * 4 * 6
    tok = torch.zeros(2, 10 + L, 64, device=dev, dtype=torch.bfloat16)
    ops.patchify(lat, tok, 10 + L, 10)
    from einops import rearrange
    ref = rearrange(rearrange(lat, "b c t h w -> b t h w c"), "b t (h p1) (w p2) c -> b (t h w) (p1 p2 c)", p1=2, p2=2)
    print(f"[elt] patchify exact: {bool((tok[:, 10:] == ref).all())}", flush=True)
    out = torch.zeros(2, 16, 2, 8, 12, device=dev)
    ops.unpatchify(tok.float().contiguous(), 10 + L, 10, out)
    print(f"[elt] unpatchify roundtrip exact: {bool((out == lat.float()).all())}", flush=True)
    v2 = torch.randn(2, 1000, device=dev)
    xs = torch.randn(1000, device=dev)
    xo = torch.zeros(1000, device=dev)
    ops.cfg_euler_step(v2, 5.0, -0.05, xs, xo)
    ref = xs + (-0.05) * (v2[0] + 5.0 * (v2[1] - v2[0]))
    print(f"[elt] cfg_euler: max_abs={float((xo - ref).abs().max()):.3e}", flush=True)


def _time_cuda(fn, iters=10, warm=3):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s = torch.cuda.Event(enable_timing=True)
    e = torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def group_gemm_perf():
    import torch
    from pyramid_flow_b200 import ops
    dev = "cuda"
    for (m, n, k) in [(30976, 1920, 1920), (30976, 5760, 1920), (30976, 7680, 1920), (30976, 1920, 7680),
                      (30976, 1920, 9600), (30976, 2048, 8192), (30976, 1536, 6144), (7744, 1920, 9600)]:
        x = (torch.randn(m, k, device=dev) * 0.5).bfloat16()
        w = (torch.randn(n, k, device=dev) * 0.05).bfloat16()
        out = torch.empty(m, n, device=dev, dtype=torch.bfloat16)
        res = {}
        for mode, kvar in (("0", 1), ("1", 2), ("auto", 0)):
            res[mode] = _time_cuda(lambda: ops.gemm(x, w, None, 0, rows_per_batch=m, out=out, kernel_variant=kvar))
        ms_ref = _time_cuda(lambda: torch.matmul(x, w.t(), out=out))
        fl = 2.0 * m * n * k / 1e9
        print(f"[gemm_perf] m={m} n={n} k={k}: 1-CTA {res['0']:.3f} ms = {fl/res['0']:.0f} TF/s | 2-CTA {res['1']:.3f} ms = {fl/res['1']:.0f} | auto {res['auto']:.3f} ms = {fl/res['auto']:.0f} | cuBLAS {ms_ref:.3f} ms = {fl/ms_ref:.0f}", flush=True)


def group_attn():
    import torch
    from pyramid_flow_b200 import ops
    sys.path.insert(0, str(ROOT))
    torch.manual_seed(3)
    dev = "cuda"

    def case(name, B, H, S, seg, tim, variant=0):
        q = torch.randn(B, H, S, 64, device=dev).bfloat16()
        k = torch.randn(B, H, S, 64, device=dev).bfloat16()
        v = torch.randn(B, H, S, 64, device=dev).bfloat16()
        out = torch.zeros(B, S, H * 64, device=dev, dtype=torch.bfloat16)
        sched, pairs = ops.attn_build_schedule(seg, tim)
        try:
            ops.attn_fwd(q, k, v, out, seg.to(dev).int(), tim.to(dev).int(), sched.to(dev), 0.125, variant)
            torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001
            print(f"[attn] {name}: EXC {e}", flush=True)
            return
        sg = seg.to(dev)
        tm = tim.to(dev)
        mask = (sg[:, :, None] == sg[:, None, :]) & (tm[:, :, None] >= tm[:, None, :])
        ref = torch.nn.functional.scaled_dot_product_attention(q.float(), k.float(), v.float(), attn_mask=mask[:, None])
        ref = ref.transpose(1, 2).reshape(B, S, H * 64)
        err = (out.float() - ref).abs().max().item()
        print(f"[attn] {name} (variant {variant}): max_abs_err={err:.3e} pairs={pairs.tolist()} {'OK' if err < 3e-2 else 'MISMATCH'}", flush=True)

    for variant in (0, 1):
        S = 256
        case("dense S=256", 1, 2, S, torch.ones(1, S, dtype=torch.int32), torch.zeros(1, S, dtype=torch.int32), variant)
        S = 384
        tim = torch.cat([torch.zeros(128), torch.ones(128), 2 * torch.ones(128)]).int()[None]
        case("tile-aligned causal S=384", 1, 2, S, torch.ones(1, S, dtype=torch.int32), tim, variant)
        S = 128 + 60 * 5
        tim = torch.cat([torch.zeros(128 + 60)] + [torch.full((60,), i + 1.0) for i in range(4)]).int()[None].repeat(2, 1)
        seg = torch.ones(2, S, dtype=torch.int32)
        seg[0, 37:128] = 0
        case("ragged text + 5 frames of 60, S=428", 2, 3, S, seg, tim, variant)
        S = 128 + 240 * 9
        tim = torch.cat([torch.zeros(128 + 240)] + [torch.full((240,), i + 1.0) for i in range(8)]).int()[None].repeat(2, 1)
        seg = torch.ones(2, S, dtype=torch.int32)
        seg[0, 100:128] = 0
        case("S=2288", 2, 4, S, seg, tim, variant)


def group_attn_perf():
    import torch
    from pyramid_flow_b200 import ops
    dev = "cuda"
    B, H = 2, 30
    # 768p last unit, stage 2: text 128 | 28 frames @240 | 1 @960 | 2 @3840  (S = 15488)
    lens = [128 + 240] + [240] * 27 + [960, 3840, 3840]
    tim = torch.cat([torch.full((n,), float(i)) for i, n in enumerate(lens)]).int()[None].repeat(B, 1)
    S = tim.shape[1]
    seg = torch.ones(B, S, dtype=torch.int32)
    q = torch.randn(B, H, S, 64, device=dev).bfloat16()
    k = torch.randn(B, H, S, 64, device=dev).bfloat16()
    v = torch.randn(B, H, S, 64, device=dev).bfloat16()
    out = torch.zeros(B, S, H * 64, device=dev, dtype=torch.bfloat16)
    sched, pairs = ops.attn_build_schedule(seg, tim)
    sd, td, scd = seg.to(dev), tim.to(dev), sched.to(dev)
    flops = 4.0 * 64 * H * float(pairs.sum())
    ps = ops.attn_build_pair_schedule(sched, S, seg, tim).to(dev)
    ref = None
    for variant in (3, 0x10, 0x20, 0):
        out.zero_()
        ms = _time_cuda(lambda: ops.attn_fwd(q, k, v, out, sd, td, scd, 0.125, variant, pair_sched=ps), iters=8, warm=2)
        if ref is None:
            ref = out.clone()
        dev_ = (out.float() - ref.float()).abs().max().item()
        print(f"[attn_perf] variant {variant:#x} S={S} B={B} H={H}: {ms:.3f} ms, {flops/ms/1e9:.0f} TFLOP/s (masked-pair flops), "
              f"max |out - one-tile kernel| {dev_:.2e}", flush=True)


def group_gemm_epi_perf():
    """Epilogue cost: the same K=1920 GEMM with STORE / GELU+bias / GATE_RESID / QKV_ROPE epilogues."""
    import torch
    from pyramid_flow_b200 import ops
    dev = "cuda"
    m, k = 30976, 1920
    x = (torch.randn(m, k, device=dev) * 0.5).bfloat16()
    for n, epi, name in [(7680, 0, "store"), (7680, 1, "gelu+bias"), (1920, 0, "store"), (1920, 3, "gate_resid")]:
        w = (torch.randn(n, k, device=dev) * 0.05).bfloat16()
        bias = torch.randn(n, device=dev) * 0.1
        fl = 2.0 * m * n * k / 1e9
        if epi == 3:
            h = torch.zeros(2, m // 2, n, device=dev, dtype=torch.float32)
            gate = torch.randn(2, n, device=dev) * 0.1
            fn = lambda: ops.gemm(x, w, bias, 3, rows_per_batch=m // 2, out=h, gate=gate, gate_batch_stride=n, batches=2)
        else:
            out = torch.empty(m, n, device=dev, dtype=torch.bfloat16)
            fn = lambda: ops.gemm(x, w, bias if epi == 1 else None, epi, rows_per_batch=m, out=out)
        try:
            ms = _time_cuda(fn)
            print(f"[gemm_epi_perf] m={m} n={n} k={k} {name}: {ms:.3f} ms = {fl/ms:.0f} TF/s", flush=True)
            if epi == 3:      # the staged (transposed) read-modify-write, pf_set_option(PF_OPT_GEMM_STAGED_RESID)
                from pyramid_flow_b200 import _lib
                old = _lib.get_option(_lib.PF_OPT_GEMM_STAGED_RESID)
                _lib.set_option(_lib.PF_OPT_GEMM_STAGED_RESID, 1 - old)
                ms2 = _time_cuda(fn)
                _lib.set_option(_lib.PF_OPT_GEMM_STAGED_RESID, old)
                print(f"[gemm_epi_perf] m={m} n={n} k={k} {name} with staged={1 - old}: {ms2:.3f} ms = {fl/ms2:.0f} TF/s", flush=True)
        except Exception as e:  # noqa: BLE001
            print(f"[gemm_epi_perf] {name}: EXC {e}", flush=True)
    # the sequence-parallel chunk shape (8 GPUs): wave-quantisation-aware tile width on/off, K = 1920 / 7680 / 9600
    from pyramid_flow_b200 import _lib
    for kk in (1920, 7680, 9600):
        mm, nn = 3872, 1920
        xs = (torch.randn(mm, kk, device=dev) * 0.5).bfloat16()
        ws = (torch.randn(nn, kk, device=dev) * 0.05).bfloat16()
        hh = torch.zeros(1, mm, nn, device=dev, dtype=torch.float32)
        gg = torch.randn(1, nn, device=dev) * 0.1
        f2 = lambda: ops.gemm(xs, ws, None, 3, rows_per_batch=mm, out=hh, gate=gg, gate_batch_stride=nn, batches=1)
        res = {}
        for wave in (0, 1):
            _lib.set_option(_lib.PF_OPT_GEMM_WAVE_TILING, wave)
            res[wave] = _time_cuda(f2, iters=20)
        _lib.set_option(_lib.PF_OPT_GEMM_WAVE_TILING, 0)
        fl2 = 2.0 * mm * nn * kk / 1e9
        print(f"[gemm_epi_perf] m={mm} n={nn} k={kk} gate_resid: default tiling {res[0]:.3f} ms = {fl2/res[0]:.0f} TF/s | wave-aware {res[1]:.3f} ms = {fl2/res[1]:.0f} TF/s", flush=True)


def group_gemm_qkv_perf():
    """QKV GEMM with the RMSNorm + RoPE head-major epilogue vs the plain store at the same shape (N=5760, K=1920)."""
    import torch
    from pyramid_flow_b200 import ops
    dev = "cuda"
    B, S, H, D = 2, 15488, 30, 1920
    m = B * S
    x = (torch.randn(B, S, D, device=dev) * 0.5).bfloat16()
    w = (torch.randn(3 * D, D, device=dev) * 0.05).bfloat16()
    bias = torch.randn(3 * D, device=dev) * 0.1
    q, k, v = (torch.empty(B, H, S, 64, device=dev, dtype=torch.bfloat16) for _ in range(3))
    rope = torch.randn(S, 32, 2, device=dev)
    nq, nk = torch.ones(64, device=dev), torch.ones(64, device=dev)
    out = torch.empty(m, 3 * D, device=dev, dtype=torch.bfloat16)
    fl = 2.0 * m * 3 * D * D / 1e9
    ms0 = _time_cuda(lambda: ops.gemm(x, w, bias, 0, batches=B, rows_per_batch=S, out=out.view(B, S, 3 * D)))
    ms1 = _time_cuda(lambda: ops.gemm(x, w, bias, 4, batches=B, rows_per_batch=S, q_out=q, k_out=k, v_out=v, rope=rope,
                                      q_norm_w=nq, k_norm_w=nk, heads=H, head_dim=64, seq_len=S))
    print(f"[gemm_qkv_perf] m={m} n={3*D} k={D}: store+bias {ms0:.3f} ms = {fl/ms0:.0f} TF/s | qkv_rope {ms1:.3f} ms = {fl/ms1:.0f} TF/s", flush=True)


def group_attn_trace():
    """clock64 timeline of one CTA of the attention kernel at the bench shape (variant 2)."""
    import torch
    from pyramid_flow_b200 import ops, _lib
    dev = "cuda"
    B, H = 2, 30
    lens = [128 + 240] + [240] * 27 + [960, 3840, 3840]
    tim = torch.cat([torch.full((n,), float(i)) for i, n in enumerate(lens)]).int()[None].repeat(B, 1)
    S = tim.shape[1]
    seg = torch.ones(B, S, dtype=torch.int32)
    q = torch.randn(B, H, S, 64, device=dev).bfloat16()
    k = torch.randn(B, H, S, 64, device=dev).bfloat16()
    v = torch.randn(B, H, S, 64, device=dev).bfloat16()
    out = torch.zeros(B, S, H * 64, device=dev, dtype=torch.bfloat16)
    sched, pairs = ops.attn_build_schedule(seg, tim)
    sd, td, scd = seg.to(dev), tim.to(dev), sched.to(dev)
    N, SL = 48, 8
    buf = torch.zeros(3 * N * SL, dtype=torch.int64, device=dev)
    _lib.check(_lib.load().pf_debug_attn_trace(buf.data_ptr()), "trace")
    variant = int(os.environ.get("PF_TRACE_VARIANT", "2"), 0)     # 2 = one-tile trace kernel; 0x10 = two-q-tile kernel
    ps = ops.attn_build_pair_schedule(sched, S, seg, tim).to(dev)
    if variant & 0x10:
        n_cta = ((((S + 127) // 128) + 1) // 2) * H * B
    for _ in range(2):
        ops.attn_fwd(q, k, v, out, sd, td, scd, 0.125, variant, pair_sched=ps)
    torch.cuda.synchronize()
    _lib.check(_lib.load().pf_debug_attn_trace(None), "trace")
    t = buf.cpu().view(3, N, SL)
    t0 = int(t[0, 0, 0])
    print("[attn_trace] softmax slots: 0 S_full seen, 1 ld done, 5 max done, 2 reference known (p_full(j-1) + partner max), "
          "3 exps done, 4 P stored (after pv_done(j-1)), 6 arrived p_full | MMA slots: 0 s_free seen, 1 QK(j+1) issued, "
          "2 p_full seen, 3 v_full, 4 PV(j) issued")
    order = [0, 1, 5, 2, 3, 4, 6]
    for j in range(8, 16):
        sm = " ".join(f"{int(t[0, j, i]) - t0:7d}" for i in order)
        mm = " ".join(f"{int(t[2, j, i]) - t0:7d}" for i in range(5))
        print(f"[attn_trace] j={j:2d}: softmax0 {sm} | mma {mm}")
    per = (int(t[0, 40, 0]) - int(t[0, 8, 0])) / 32.0
    d = t[:, 8:40].double()
    ph = lambda r, x, y: float((d[r, :, y] - d[r, :, x]).mean())
    print(f"[attn_trace] cycles per kv tile (this CTA): {per:.0f}")
    print(f"[attn_trace] softmax half0 mean phases: S_full->ld {ph(0,0,1):.0f} | ld->max {ph(0,1,5):.0f} | max->ref_known "
          f"{ph(0,5,2):.0f} | exps {ph(0,2,3):.0f} | pv_done wait + P store {ph(0,3,4):.0f} | fence+arrive {ph(0,4,6):.0f} | "
          f"arrive->next S_full {float((d[0,1:,0]-d[0,:-1,6]).mean()):.0f}")
    print(f"[attn_trace] MMA: arrive(half0)->p_full seen {float((d[2,:,2]-d[0,:,6]).mean()):.0f}, p_full->PV issued "
          f"{ph(2,2,4):.0f}, ld_done(half0)->s_free seen {float((d[2,:,0]-d[0,:,1]).mean()):.0f}, "
          f"QK issue {ph(2,0,1):.0f}, QK(j+1) issued->S_full(j+1) seen {float((d[0,1:,0]-d[2,:-1,1]).mean()):.0f}")


def group_attn_cta_trace():
    """Per-CTA cost model of the attention kernel at the bench shape: cycles(CTA) ~ F + c * kv_tiles (least squares over all
    7260 CTAs of one launch of the trace variant), plus how busy each SM's two CTA slots were."""
    import torch
    from pyramid_flow_b200 import ops, _lib
    dev = "cuda"
    B, H = 2, 30
    lens = [128 + 240] + [240] * 27 + [960, 3840, 3840]
    tim = torch.cat([torch.full((n,), float(i)) for i, n in enumerate(lens)]).int()[None].repeat(B, 1)
    S = tim.shape[1]
    seg = torch.ones(B, S, dtype=torch.int32)
    q = torch.randn(B, H, S, 64, device=dev).bfloat16()
    k = torch.randn(B, H, S, 64, device=dev).bfloat16()
    v = torch.randn(B, H, S, 64, device=dev).bfloat16()
    out = torch.zeros(B, S, H * 64, device=dev, dtype=torch.bfloat16)
    sched, pairs = ops.attn_build_schedule(seg, tim)
    sd, td, scd = seg.to(dev), tim.to(dev), sched.to(dev)
    n_cta = ((S + 127) // 128) * H * B
    buf = torch.zeros(n_cta * 8, dtype=torch.int64, device=dev)
    lib = _lib.load()
    _lib.check(lib.pf_debug_attn_cta_trace(buf.data_ptr(), n_cta), "cta trace")
    variant = int(os.environ.get("PF_TRACE_VARIANT", "2"), 0)     # 2 = one-tile trace kernel; 0x10 = two-q-tile kernel
    ps = ops.attn_build_pair_schedule(sched, S, seg, tim).to(dev)
    if variant & 0x10:
        n_cta = ((((S + 127) // 128) + 1) // 2) * H * B
    for _ in range(2):
        ops.attn_fwd(q, k, v, out, sd, td, scd, 0.125, variant, pair_sched=ps)
    torch.cuda.synchronize()
    _lib.check(lib.pf_debug_attn_cta_trace(None, 0), "cta trace")
    r = buf.cpu()[:n_cta * 8].view(n_cta, 8).double()
    ph = lambda a_, b_: float((r[:, b_] - r[:, a_]).mean())
    print(f"[attn_cta_trace] mean phases per CTA (clk): entry -> alloc+sync done {ph(0, 4):.0f} | -> first S tile seen {ph(4, 5):.0f} | "
          f"softmax loop {ph(5, 6):.0f} ({float(((r[:, 6] - r[:, 5]) / r[:, 2]).mean()):.0f} per kv tile) | loop end -> exit {ph(6, 1):.0f} | "
          f"last PV issued -> exit {ph(7, 1):.0f}")
    for lo, hi in [(1, 8), (8, 32), (32, 64), (64, 200)]:
        m_ = (r[:, 2] >= lo) & (r[:, 2] < hi)
        if m_.any():
            rr = r[m_]
            print(f"[attn_cta_trace]   kv tiles in [{lo}, {hi}): entry->sync {float((rr[:, 4] - rr[:, 0]).mean()):.0f}, sync->first S {float((rr[:, 5] - rr[:, 4]).mean()):.0f}, "
                  f"loop per tile {float(((rr[:, 6] - rr[:, 5]) / rr[:, 2]).mean()):.0f}, loop end->exit {float((rr[:, 1] - rr[:, 6]).mean()):.0f}")
    dur, nkv, sm = r[:, 1] - r[:, 0], r[:, 2], r[:, 3].long()
    A = torch.stack([torch.ones_like(nkv), nkv], 1)
    sol = torch.linalg.lstsq(A, dur[:, None]).solution.flatten()
    print(f"[attn_cta_trace] {n_cta} CTAs, kv tiles per CTA mean {nkv.mean():.1f}: cycles(CTA) = {sol[0]:.0f} + {sol[1]:.0f} * kv_tiles "
          f"(fixed part = {100 * sol[0] * n_cta / dur.sum():.1f} % of all CTA cycles)")
    for lo, hi in [(1, 8), (8, 32), (32, 64), (64, 200)]:
        m = (nkv >= lo) & (nkv < hi)
        if m.any():
            print(f"[attn_cta_trace] CTAs with {lo:3d} <= kv tiles < {hi:3d}: {int(m.sum()):5d}, cycles per kv tile {float((dur[m] / nkv[m]).mean()):.0f}")
    t_begin, t_end = r[:, 0].min(), r[:, 1].max()
    busy = torch.zeros(int(sm.max()) + 1, dtype=torch.float64).index_add_(0, sm, dur)
    # clock64 is per SM: spans are only comparable within one SM
    span = torch.zeros_like(busy)
    for s_id in range(busy.numel()):
        m = sm == s_id
        if m.any():
            span[s_id] = r[m, 1].max() - r[m, 0].min()
    print(f"[attn_cta_trace] per SM: CTA-slot occupancy (sum of CTA cycles / (2 * span)) mean {float((busy / (2 * span)).mean()):.3f} "
          f"min {float((busy / (2 * span)).min()):.3f}; span mean {float(span.mean()):.0f} cycles, max {float(span.max()):.0f}")


def group_attn_timeline():
    """clock64 timeline of CTA (0, 0, 0) of the two-q-tile attention kernel (pf_attn2.cu, timeline instantiation) at the bench
    shape: softmax thread 0 of q tile A / B and the two MMA issuers, first 64 kv tiles; PF_TL_PHASE = PF_OPT_ATTN_TILE_PHASE."""
    import torch
    from pyramid_flow_b200 import ops, _lib
    dev = "cuda"
    B, H = 2, 30
    lens = [128 + 240] + [240] * 27 + [960, 3840, 3840]
    tim = torch.cat([torch.full((n,), float(i)) for i, n in enumerate(lens)]).int()[None].repeat(B, 1)
    S = tim.shape[1]
    seg = torch.ones(B, S, dtype=torch.int32)
    q = torch.randn(B, H, S, 64, device=dev).bfloat16()
    k = torch.randn(B, H, S, 64, device=dev).bfloat16()
    v = torch.randn(B, H, S, 64, device=dev).bfloat16()
    out = torch.zeros(B, S, H * 64, device=dev, dtype=torch.bfloat16)
    sched, pairs = ops.attn_build_schedule(seg, tim)
    sd, td, scd = seg.to(dev), tim.to(dev), sched.to(dev)
    ps = ops.attn_build_pair_schedule(sched, S, seg, tim).to(dev)
    N, SL = 64, 12
    names = ["top", "S_full", "ld+s_free", "max+rescale test", "exp first half", "pv wait+st", "exp second half", "st+wait+p_full"]
    ns = len(names)
    for phase in [int(x) for x in os.environ.get("PF_TL_PHASE", "-1").split()]:
        if phase >= 0:
            _lib.set_option(_lib.PF_OPT_ATTN_TILE_PHASE, phase)
        print(f"[attn_timeline] PF_OPT_ATTN_TILE_PHASE = {_lib.get_option(_lib.PF_OPT_ATTN_TILE_PHASE)}")
        buf = torch.zeros(4 * N * SL, dtype=torch.int64, device=dev)
        _lib.check(_lib.load().pf_debug_attn_trace(buf.data_ptr()), "trace")
        for _ in range(2):
            ops.attn_fwd(q, k, v, out, sd, td, scd, 0.125, 0x10, pair_sched=ps)
        torch.cuda.synchronize()
        _lib.check(_lib.load().pf_debug_attn_trace(None), "trace")
        t = buf.cpu().view(4, N, SL).double()
        t0 = float(t[0, 16, 0])
        print("[attn_timeline] softmax slots " + " ".join(f"{i}:{n}" for i, n in enumerate(names)) +
              " | MMA slots 0:s_free seen 1:QK(j+1) issued 2:p_full seen 3:v_full 4:PV issued")
        for j in range(16, 19):
            for r, nm in ((0, "sm A"), (1, "sm B"), (2, "mmaA"), (3, "mmaB")):
                n = ns if r < 2 else 5
                print(f"[attn_timeline] j={j:2d} {nm}: " + " ".join(f"{int(t[r, j, i] - t0):7d}" for i in range(n)))
        lo, hi = 8, 56
        per = float(t[0, hi, 0] - t[0, lo, 0]) / (hi - lo)
        for r, nm in ((0, "A"), (1, "B")):
            d = t[r, lo:hi]
            ph = " | ".join(f"{names[i + 1]} {float((d[:, i + 1] - d[:, i]).mean()):.0f}" for i in range(ns - 1))
            nxt = float((t[r, lo + 1:hi + 1, 0] - t[r, lo:hi, ns - 1]).mean())
            print(f"[attn_timeline] tile {nm}: cycles per kv tile {per:.0f}; mean phase durations: {ph} | loop back {nxt:.0f}")
        for r, nm in ((2, "A"), (3, "B")):
            d = t[r, lo:hi]
            sm = t[r - 2, lo:hi]
            print(f"[attn_timeline] MMA {nm}: s_free arrive->seen {float((d[:, 0] - sm[:, 2]).mean()):.0f} | QK issue {float((d[:, 1] - d[:, 0]).mean()):.0f} | "
                  f"p_full arrive->seen {float((d[:, 2] - sm[:, 7]).mean()):.0f} | v_full wait {float((d[:, 3] - d[:, 2]).mean()):.0f} | PV issue {float((d[:, 4] - d[:, 3]).mean()):.0f}")
    _lib.set_option(_lib.PF_OPT_ATTN_TILE_PHASE, 800)


def group_attn_phase_sweep():
    """Attention launch time at the bench shape as a function of PF_OPT_ATTN_TILE_PHASE (clocks q tile B starts late)."""
    import torch
    from pyramid_flow_b200 import ops, _lib
    dev = "cuda"
    B, H = 2, 30
    lens = [128 + 240] + [240] * 27 + [960, 3840, 3840]
    tim = torch.cat([torch.full((n,), float(i)) for i, n in enumerate(lens)]).int()[None].repeat(B, 1)
    S = tim.shape[1]
    seg = torch.ones(B, S, dtype=torch.int32)
    q = torch.randn(B, H, S, 64, device=dev).bfloat16()
    k = torch.randn(B, H, S, 64, device=dev).bfloat16()
    v = torch.randn(B, H, S, 64, device=dev).bfloat16()
    out = torch.zeros(B, S, H * 64, device=dev, dtype=torch.bfloat16)
    sched, pairs = ops.attn_build_schedule(seg, tim)
    sd, td, scd = seg.to(dev), tim.to(dev), sched.to(dev)
    flops = 4.0 * 64 * H * float(pairs.sum())
    ps = ops.attn_build_pair_schedule(sched, S, seg, tim).to(dev)
    ref = None
    variants = [int(x, 0) for x in os.environ.get("PF_SWEEP_VARIANTS", "0x10").split()]
    for delay in [int(x) for x in os.environ.get("PF_SWEEP_DELAYS", "0 300 600 800 1000 1200 1500 1800 2400 3000").split()]:
        _lib.set_option(_lib.PF_OPT_ATTN_TILE_PHASE, delay)
        for variant in variants:
            out.zero_()
            ms = _time_cuda(lambda: ops.attn_fwd(q, k, v, out, sd, td, scd, 0.125, variant, pair_sched=ps), iters=8, warm=2)
            if ref is None:
                ref = out.clone()
            same = bool(torch.equal(out, ref)) if variant == variants[0] else None
            print(f"[attn_phase_sweep] phase {delay:5d} clk, variant {variant:#x}: {ms:.3f} ms, {flops/ms/1e9:.0f} TFLOP/s"
                  + ("" if same is None else f", bits == phase-0 output: {same}"), flush=True)
    _lib.set_option(_lib.PF_OPT_ATTN_TILE_PHASE, 800)


def group_vae_perf():
    import torch
    from pyramid_flow_b200.vae import B200CausalVAE, VaeConfigB200
    from pyramid_flow_b200 import _lib
    dev = torch.device("cuda:0")
    from bench import random_vae_state_dict
    cfg = VaeConfigB200()
    sd = random_vae_state_dict(dev)
    vae = B200CausalVAE(cfg, sd, device=dev)
    for (T, h, w, win) in [(3, 48, 80, 2), (5, 96, 160, 2)]:
        z = torch.randn(1, 16, T, h, w, device=dev).bfloat16()
        torch.cuda.synchronize()
        n0 = _lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        out = vae.decode(z, temporal_chunk=True, window_size=win).sample   # warm-up (allocations)
        torch.cuda.synchronize()
        e0.record()
        out = vae.decode(z, temporal_chunk=True, window_size=win).sample
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        frames = out.shape[2]
        macs = 1.10e7 * frames * out.shape[3] * out.shape[4]
        print(f"[vae_perf] latent {T}x{h}x{w} -> {tuple(out.shape)}: {ms:.1f} ms, {frames / (ms * 1e-3):.1f} frames/s, "
              f"{2 * macs / (ms * 1e-3) / 1e12:.0f} TFLOP/s (1.10e7 MAC/px-frame), peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB, launches {(_lib.launch_count() - n0) // 2}", flush=True)


# ----------------------------------------------------------------------------------------------------------------
def main():
    args = sys.argv[1:]
    if args and args[0] == "--child":
        from pyramid_flow_b200 import _lib
        _lib.require_device()
        globals()["group_" + args[1]]()
        return
    groups = args or ["probe", "gemm", "epilogue", "elementwise"]
    outdir = ROOT / "gpurun_out"
    outdir.mkdir(exist_ok=True)
    for g in groups:
        t0 = time.time()
        log = outdir / f"gpu_check_{g}.log"
        with open(log, "w") as f:
            try:
                p = subprocess.run([sys.executable, __file__, "--child", g], stdout=subprocess.PIPE,
                                   stderr=subprocess.STDOUT, text=True, timeout=int(os.environ.get("PF_CHECK_TIMEOUT", "240")))
                out, rc = p.stdout, p.returncode
            except subprocess.TimeoutExpired as e:
                out = (e.stdout or b"").decode() if isinstance(e.stdout, bytes) else (e.stdout or "")
                out += "\n[TIMEOUT]\n"
                rc = -999
            f.write(out)
        print(f"===== {g}: rc={rc} ({time.time()-t0:.1f}s)\n{out}", flush=True)


if __name__ == "__main__":
    main()
