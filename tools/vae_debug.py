import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from oracle import vae_oracle as VO
from pyramid_flow_b200.vae import B200CausalVAE, VaeConfigB200

kw = dict(block_out_channels=(64, 128, 128, 128), layers_per_block=(2, 2, 2, 2))
cfg = VO.VaeDecoderConfig(**kw)
params = VO.synthetic_vae_params(cfg, seed=0)
g = torch.Generator().manual_seed(2)
z = torch.randn(1, 16, 3, 6, 10, generator=g).bfloat16().float()
dev = torch.device("cuda:0")
vae = B200CausalVAE(VaeConfigB200(**kw), params, device=dev)
full = vae.decode(z.to(dev), temporal_chunk=False).sample.float().cpu()
full2 = vae.decode(z.to(dev), temporal_chunk=False).sample.float().cpu()
print("run-to-run", (full - full2).abs().max().item())
for w in (1, 2):
    c = vae.decode(z.to(dev), temporal_chunk=True, window_size=w).sample.float().cpu()
    d = (c - full).abs().amax(dim=(0, 1, 3, 4))
    print("window", w, "per-frame max diff", [f"{x:.3f}" for x in d.tolist()])
with torch.no_grad():
    ref = VO.decode(params, cfg, z)
print("full vs oracle per-frame", [f"{x:.3f}" for x in (full - ref).abs().amax(dim=(0, 1, 3, 4)).tolist()])
c1 = vae.decode(z.to(dev), temporal_chunk=True, window_size=1).sample.float().cpu()
print("chunk1 vs oracle per-frame", [f"{x:.3f}" for x in (c1 - ref).abs().amax(dim=(0, 1, 3, 4)).tolist()])
