"""Why is the graph-replayed step slower than the host-launched one in bench.py?  Alternate blocks of 10 steps of each and print
ms/step with the SM clock / power sampled during each block.   python tools/graph_vs_eager.py"""
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch

from bench import random_flux_state_dict, step_clip_shapes
from pyramid_flow_b200.dit import B200FluxTransformer

dev = torch.device("cuda:0")
cfg, sd = random_flux_state_dict(dict(num_layers=8, num_single_layers=16), dev, seed=0)
model = B200FluxTransformer(cfg, sd, device=dev)
del sd
g = torch.Generator().manual_seed(100)
clips = [torch.randn(s, generator=g).bfloat16().to(dev) for s in step_clip_shapes(2)]
enc = (torch.randn(2, 128, 4096, generator=g) * 0.2).bfloat16().to(dev)
mask = torch.ones(2, 128, dtype=torch.int64, device=dev)
pooled = torch.randn(2, 768, generator=g).bfloat16().to(dev)
t = torch.tensor([3.0, 3.0]).bfloat16().to(dev)


def step():
    return model(sample=[clips], timestep_ratio=t, encoder_hidden_states=enc, encoder_attention_mask=mask, pooled_projections=pooled)[0]


lines = []
proc = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "50"],
                        stdout=subprocess.PIPE, text=True)
threading.Thread(target=lambda: [lines.append((time.time(), ln.strip())) for ln in proc.stdout], daemon=True).start()


def block(name, graph, n=10):
    model.use_cuda_graph = graph
    step()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    s.record()
    for _ in range(n):
        step()
    e.record()
    torch.cuda.synchronize()
    t1 = time.time()
    smp = [ln.split(",") for ts, ln in lines if t0 <= ts <= t1]
    clk = sorted(float(x[0]) for x in smp) if smp else [0]
    pw = sorted(float(x[1]) for x in smp) if smp else [0]
    print(f"[graph_vs_eager] {name:26s}: {s.elapsed_time(e) / n:8.2f} ms/step | SM clock median {clk[len(clk) // 2]:.0f} MHz, power median {pw[len(pw) // 2]:.0f} W ({len(smp)} samples)", flush=True)


for _ in range(3):
    step()
for rnd in range(2):
    block("graph replay", True)
    block("host-launched", False)
model.attn_events = []
block("host-launched + attn events", False)
model.attn_events = None
block("graph replay", True)
proc.terminate()
