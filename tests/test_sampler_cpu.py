"""The sampler-loop mirror (pyramid_flow_b200/sampler.py + scheduler.py) against the UNMODIFIED reference generate() loop
(tests/golden/sampler_small.pt), with the oracle standing in for the DiT.  CPU only."""
import torch

from oracle import flux_oracle as FO
from pyramid_flow_b200.sampler import B200PyramidSampler, block_noise
from pyramid_flow_b200.scheduler import B200FlowMatchScheduler


class OracleDit:
    """Adapter exposing the reference DiT call surface on top of the oracle restatement (test-only)."""

    def __init__(self, cfg, params):
        self.cfg, self.params = cfg, params
        self.config = type("C", (), {"in_channels": cfg.in_channels})()
        self.shapes = []

    def __call__(self, sample, timestep_ratio, encoder_hidden_states, encoder_attention_mask, pooled_projections):
        clips = sample[0]
        self.shapes.append([tuple(c.shape) for c in clips])
        return [FO.flux_forward(self.params, self.cfg, clips, timestep_ratio, encoder_hidden_states,
                                encoder_attention_mask, pooled_projections)]


def test_sampler_loop_matches_reference_generate(golden_dir):
    g = torch.load(golden_dir / "sampler_small.pt", weights_only=False)
    cfg = FO.FluxConfig(**g["cfg"])
    dit = OracleDit(cfg, FO.synthetic_flux_params(cfg, seed=g["param_seed"]))
    noises = list(g["noises"])
    sampler = B200PyramidSampler(dit, B200FlowMatchScheduler(), block_noise_fn=lambda *a: noises.pop(0))
    gen = torch.Generator().manual_seed(g["latent_seed"])
    with torch.no_grad():
        lat = sampler.generate(g["enc"], g["mask"], g["pooled"], generator=gen, output_type="latent", **g["args"])
    assert lat.shape == g["latents"].shape
    assert not noises, "every injected block-noise tensor must be consumed in the same order"
    err = (lat - g["latents"]).abs().max().item()
    assert err < 5e-4, err                      # fp32 vs fp32, 26 DiT calls deep
    # the clip lists the DiT sees follow P:1159-1182 (compressed history: low-res old frames first, current last)
    assert dit.shapes[0] == [(2, 16, 1, 4, 4)]
    assert dit.shapes[-1] == [(2, 16, 1, 4, 4), (2, 16, 1, 8, 8), (2, 16, 1, 16, 16), (2, 16, 1, 16, 16)]
    assert sampler.dit_calls == 6 + 3 * 5


def test_block_noise_covariance():
    gen = torch.Generator().manual_seed(0)
    z = block_noise(4, 16, 2, 32, 32, 1 / 3, gen)
    blocks = z.reshape(4, 16, 2, 16, 2, 16, 2).permute(0, 1, 2, 3, 5, 4, 6).reshape(-1, 4)
    cov = (blocks.T @ blocks) / blocks.shape[0]
    target = torch.eye(4) * (1 + 1 / 3) - torch.ones(4, 4) / 3
    assert (cov - target).abs().max().item() < 0.03


def test_i2v_loop_matches_reference_generate_i2v(golden_dir):
    """generate_i2v mirror (P:791-1003) against the unmodified reference loop with a fake VAE encode (fixed image latent)."""
    g = torch.load(golden_dir / "sampler_i2v_small.pt", weights_only=False)
    cfg = FO.FluxConfig(**g["cfg"])
    dit = OracleDit(cfg, FO.synthetic_flux_params(cfg, seed=g["param_seed"]))
    noises = list(g["noises"])
    sampler = B200PyramidSampler(dit, B200FlowMatchScheduler(), block_noise_fn=lambda *a: noises.pop(0))
    gen = torch.Generator().manual_seed(g["latent_seed"])
    with torch.no_grad():
        lat = sampler.generate_i2v(g["image_tensor"], g["enc"], g["mask"], g["pooled"], generator=gen, output_type="latent",
                                   image_latent=g["image_latent_raw"], **g["args"])
    assert lat.shape == g["latents"].shape and not noises
    err = (lat - g["latents"]).abs().max().item()
    assert err < 5e-4, err
    # unit 0 is the normalised image latent itself (P:911)
    ref0 = (g["image_latent_raw"] - sampler.vae_shift_factor) * sampler.vae_scale_factor
    assert torch.allclose(lat[:, :, :1], ref0, atol=1e-6)
    assert sampler.dit_calls == 3 * 5            # 3 generated units x (2 + 1 + 2) steps
