"""Run ONE launch of a hot kernel at the bench shape (for `ncu --set full`): python tools/prof_one.py attn|gemm|ln"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from pyramid_flow_b200 import ops

which = sys.argv[1] if len(sys.argv) > 1 else "attn"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = "cuda"
if which == "attn":
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
    for _ in range(reps):
        ops.attn_fwd(q, k, v, out, sd, td, scd, 0.125)
elif which == "gemm":
    m, n, k = 30976, 7680, 1920
    x = (torch.randn(m, k, device=dev) * 0.5).bfloat16()
    w = (torch.randn(n, k, device=dev) * 0.05).bfloat16()
    o = torch.empty(m, n, device=dev, dtype=torch.bfloat16)
    for _ in range(reps):
        ops.gemm(x, w, None, 1, rows_per_batch=m, out=o)
torch.cuda.synchronize()
