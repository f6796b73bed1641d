"""ORACLE (test infrastructure, not product code): plain-PyTorch restatement of the reference SD3-MMDiT DiT step.

Restates `PyramidDiffusionMMDiT.forward` (pyramid_dit/mmdit_modules/modeling_pyramid_mmdit.py:420-497) for the configuration
the pipeline builds (P:82-89 + checkpoint config): pos_embed_type='sincos', temp_pos_embed_type='rope',
add_temp_pos_embed=True, use_t5_mask=True, use_flash_attn=False, use_temporal_causal=True, interp_condition_pos=True.
  M = .../modeling_pyramid_mmdit.py   MB = .../modeling_mmdit_block.py   ME = .../modeling_embedding.py
  MN = .../modeling_normalization.py
Differences from miniFLUX (oracle/flux_oracle.py): conv2d(k=2,s=2) patch embed + cropped/interpolated 2-D sincos table
(ME:269-358), temporal-only RoPE with one 64-wide axis (M:116, M:235-262), 24 double blocks only, the last block is
`context_pre_only` (MB:585-622, 659-660), q/k RMSNorm eps 1e-5 (JointAttention default, MB:409).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.  Pinned by
tests/golden/mmdit_small.pt (oracle/pin/make_golden.py mmdit).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Sequence

import numpy as np
import torch
import torch.nn.functional as F

from . import flux_oracle as FO

Params = Dict[str, torch.Tensor]


@dataclass
class MMDiTConfig:
    num_layers: int = 24
    num_attention_heads: int = 24
    attention_head_dim: int = 64
    in_channels: int = 16
    patch_size: int = 2
    joint_attention_dim: int = 4096
    pooled_projection_dim: int = 2048
    pos_embed_max_size: int = 192
    sample_size: int = 128

    @property
    def inner_dim(self) -> int:
        return self.num_attention_heads * self.attention_head_dim


def sincos_2d_table(embed_dim: int, grid: int, base_size: int) -> torch.Tensor:
    """get_2d_sincos_pos_embed (ME:23-74) with interpolation_scale=1: [grid*grid, D] float32 (sin|cos of h, then of w)."""
    gh = np.arange(grid, dtype=np.float32) / (grid / base_size)
    gw = np.arange(grid, dtype=np.float32) / (grid / base_size)
    g = np.stack(np.meshgrid(gw, gh), axis=0).reshape([2, 1, grid, grid])

    def one(d, pos):
        omega = 1.0 / 10000 ** (np.arange(d // 2, dtype=np.float64) / (d / 2.0))
        out = np.einsum("m,d->md", pos.reshape(-1), omega)
        return np.concatenate([np.sin(out), np.cos(out)], axis=1)

    emb = np.concatenate([one(embed_dim // 2, g[0]), one(embed_dim // 2, g[1])], axis=1)
    return torch.from_numpy(emb).float()


def cropped_pos_embed(table: torch.Tensor, max_size: int, h: int, w: int, ori_h: int, ori_w: int) -> torch.Tensor:
    """PatchEmbed3D.cropped_pos_embed with interp_condition_pos=True (ME:283-297): centre crop of the finest clip's grid,
    bilinear down-sampling for coarser clips.  h, w, ori_* are TOKEN grid sizes.  Returns [h*w, D]."""
    top, left = (max_size - ori_h) // 2, (max_size - ori_w) // 2
    e = table.reshape(1, max_size, max_size, -1)[:, top:top + ori_h, left:left + ori_w, :]
    if ori_h != h or ori_w != w:
        e = F.interpolate(e.permute(0, 3, 1, 2), size=(h, w), mode="bilinear").permute(0, 2, 3, 1)
    return e.reshape(h * w, -1)


def patch_embed(p: Params, cfg: MMDiTConfig, clips: List[torch.Tensor]) -> torch.Tensor:
    """PatchEmbed3D.forward (ME:360-389) -> [B, L, D]: conv2d(k=s=patch) per frame + spatial sincos (no temporal sincos:
    temp_pos_embed_type is 'rope')."""
    ps = cfg.patch_size
    oh, ow = clips[-1].shape[-2] // ps, clips[-1].shape[-1] // ps
    outs = []
    for c in clips:
        b, ch, t, hh, ww = c.shape
        x = F.conv2d(c.permute(0, 2, 1, 3, 4).reshape(b * t, ch, hh, ww), p["pos_embed.proj.weight"], p["pos_embed.proj.bias"],
                     stride=ps)
        x = x.flatten(2).transpose(1, 2)                                      # (b t) n c
        pe = cropped_pos_embed(p["pos_embed.pos_embed"][0], cfg.pos_embed_max_size, hh // ps, ww // ps, oh, ow)
        x = (x + pe[None].to(x.device)).to(x.dtype)
        outs.append(x.reshape(b, t * x.shape[1], -1))
    return torch.cat(outs, dim=1)


def time_ids(clip_shapes: Sequence[Sequence[int]], text_len: int, patch: int = 2) -> torch.Tensor:
    """[S] running frame index per token, text = 0 (M:235-262, M:301-303)."""
    out = [torch.zeros(text_len)]
    t0 = 0
    for s in clip_shapes:
        t, h, w = s[-3], s[-2] // patch, s[-1] // patch
        out.append(torch.arange(t0, t0 + t, dtype=torch.float32)[:, None].repeat(1, h * w).reshape(-1))
        t0 += t
    return torch.cat(out)


def joint_block(p: Params, pre: str, x, ctx, temb, cs, mask, heads, last: bool):
    """JointTransformerBlock.forward (MB:624-671) + JointAttention.forward (MB:470-562)."""
    b, lv, d = x.shape
    hd = d // heads
    sh_a, sc_a, g_a, sh_m, sc_m, g_m = FO.linear(p, pre + ".norm1.linear", F.silu(temb)).chunk(6, dim=1)
    xn = FO.layer_norm(x) * (1 + sc_a[:, None]) + sh_a[:, None]
    if last:   # AdaLayerNormContinuous: (scale, shift)  (MN forward, chunk order scale first)
        csc, csh = FO.linear(p, pre + ".norm1_context.linear", F.silu(temb)).chunk(2, dim=1)
        cn = FO.layer_norm(ctx) * (1 + csc[:, None]) + csh[:, None]
    else:
        csh_a, csc_a, cg_a, csh_m, csc_m, cg_m = FO.linear(p, pre + ".norm1_context.linear", F.silu(temb)).chunk(6, dim=1)
        cn = FO.layer_norm(ctx) * (1 + csc_a[:, None]) + csh_a[:, None]

    def hv(t):
        return t.view(b, -1, heads, hd)

    eps = 1e-5
    q = FO.rms_norm(hv(FO.linear(p, pre + ".attn.to_q", xn)), p[pre + ".attn.norm_q.weight"], eps)
    k = FO.rms_norm(hv(FO.linear(p, pre + ".attn.to_k", xn)), p[pre + ".attn.norm_k.weight"], eps)
    v = hv(FO.linear(p, pre + ".attn.to_v", xn))
    cq = FO.rms_norm(hv(FO.linear(p, pre + ".attn.add_q_proj", cn)), p[pre + ".attn.norm_add_q.weight"], eps)
    ck = FO.rms_norm(hv(FO.linear(p, pre + ".attn.add_k_proj", cn)), p[pre + ".attn.norm_add_k.weight"], eps)
    cv = hv(FO.linear(p, pre + ".attn.add_v_proj", cn))
    t = ctx.shape[1]
    o = FO.joint_attention(torch.cat([cq, q], 1), torch.cat([ck, k], 1), torch.cat([cv, v], 1), cs, mask, heads)
    x = x + g_a[:, None] * FO.linear(p, pre + ".attn.to_out.0", o[:, t:])
    xn2 = FO.layer_norm(x) * (1 + sc_m[:, None]) + sh_m[:, None]
    x = x + g_m[:, None] * FO.linear(p, pre + ".ff.net.2", F.gelu(FO.linear(p, pre + ".ff.net.0.proj", xn2), approximate="tanh"))
    if last:
        return None, x
    ctx = ctx + cg_a[:, None] * FO.linear(p, pre + ".attn.to_add_out", o[:, :t])
    cn2 = FO.layer_norm(ctx) * (1 + csc_m[:, None]) + csh_m[:, None]
    ctx = ctx + cg_m[:, None] * FO.linear(p, pre + ".ff_context.net.2",
                                          F.gelu(FO.linear(p, pre + ".ff_context.net.0.proj", cn2), approximate="tanh"))
    return ctx, x


def mmdit_forward(p: Params, cfg: MMDiTConfig, clips: List[torch.Tensor], timestep, encoder_hidden_states,
                  encoder_attention_mask, pooled_projections):
    """PyramidDiffusionMMDiT.forward (M:420-497) for `sample=[clips]`; returns [B, 16, t, h, w] of the last clip."""
    dev = encoder_hidden_states.device
    heads = cfg.num_attention_heads
    temb = FO.time_text_embed(p, timestep, pooled_projections)                   # M:429 (same module structure, ME:171-184)
    ctx = FO.linear(p, "context_embedder", encoder_hidden_states)                # M:430
    t_len = ctx.shape[1]
    x = patch_embed(p, cfg, clips)                                               # M:318
    tid = time_ids([c.shape for c in clips], t_len, cfg.patch_size)
    cs = FO.rope_table(tid[:, None], (cfg.attention_head_dim,)).to(dev)          # M:116, M:301-305
    seg = FO.token_segments(encoder_attention_mask, x.shape[1])
    mask = FO.attention_mask(seg, tid).to(dev)                                   # M:350-378
    for i in range(cfg.num_layers):
        ctx, x = joint_block(p, f"transformer_blocks.{i}", x, ctx, temb, cs, mask, heads, last=(i == cfg.num_layers - 1))
    scale, shift = FO.linear(p, "norm_out.linear", F.silu(temb).to(x.dtype)).chunk(2, dim=1)
    h = FO.layer_norm(x) * (1 + scale[:, None]) + shift[:, None]
    h = FO.linear(p, "proj_out", h)
    _, _, t, hh, ww = clips[-1].shape
    ps = cfg.patch_size
    n_last = t * (hh // ps) * (ww // ps)
    return FO.unpatchify(h[:, -n_last:], t, hh // ps, ww // ps, ps)              # M:407-415


def mmdit_param_shapes(cfg: MMDiTConfig) -> Dict[str, tuple]:
    d, hd = cfg.inner_dim, cfg.attention_head_dim
    s: Dict[str, tuple] = {}

    def lin(name, o, i):
        s[name + ".weight"] = (o, i)
        s[name + ".bias"] = (o,)

    s["pos_embed.pos_embed"] = (1, cfg.pos_embed_max_size ** 2, d)
    s["pos_embed.proj.weight"] = (d, cfg.in_channels, cfg.patch_size, cfg.patch_size)
    s["pos_embed.proj.bias"] = (d,)
    lin("time_text_embed.timestep_embedder.linear_1", d, 256)
    lin("time_text_embed.timestep_embedder.linear_2", d, d)
    lin("time_text_embed.text_embedder.linear_1", d, cfg.pooled_projection_dim)
    lin("time_text_embed.text_embedder.linear_2", d, d)
    lin("context_embedder", d, cfg.joint_attention_dim)
    for i in range(cfg.num_layers):
        pre = f"transformer_blocks.{i}"
        last = i == cfg.num_layers - 1
        lin(pre + ".norm1.linear", 6 * d, d)
        lin(pre + ".norm1_context.linear", (2 if last else 6) * d, d)
        for n in ("to_q", "to_k", "to_v", "add_k_proj", "add_v_proj", "add_q_proj"):
            lin(f"{pre}.attn.{n}", d, d)
        lin(pre + ".attn.to_out.0", d, d)
        for n in ("norm_q", "norm_k", "norm_add_q", "norm_add_k"):
            s[f"{pre}.attn.{n}.weight"] = (hd,)
        lin(pre + ".ff.net.0.proj", 4 * d, d)
        lin(pre + ".ff.net.2", d, 4 * d)
        if not last:
            lin(pre + ".attn.to_add_out", d, d)
            lin(pre + ".ff_context.net.0.proj", 4 * d, d)
            lin(pre + ".ff_context.net.2", d, 4 * d)
    lin("norm_out.linear", 2 * d, d)
    lin("proj_out", cfg.patch_size ** 2 * cfg.in_channels, d)
    return s


def synthetic_mmdit_params(cfg: MMDiTConfig, seed: int = 0, device: str = "cpu") -> Params:
    g = torch.Generator().manual_seed(seed)
    out: Params = {}
    for name, shp in mmdit_param_shapes(cfg).items():
        if name == "pos_embed.pos_embed":
            out[name] = sincos_2d_table(cfg.inner_dim, cfg.pos_embed_max_size, cfg.sample_size // cfg.patch_size)[None].to(device)
        elif len(shp) >= 2:
            fan_in = 1
            for k in shp[1:]:
                fan_in *= k
            std = fan_in ** -0.5
            if ".norm" in name and name.endswith("linear.weight"):
                std *= 0.5
            out[name] = (torch.randn(shp, generator=g) * std).bfloat16().float().to(device)
        elif name.endswith("weight"):
            out[name] = (1.0 + 0.1 * torch.randn(shp, generator=g)).to(device)
        else:
            out[name] = (0.02 * torch.randn(shp, generator=g)).to(device)
    return out
