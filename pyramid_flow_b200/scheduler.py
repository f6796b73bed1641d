"""Host-side mirror of `PyramidFlowMatchEulerDiscreteScheduler` (diffusion_schedulers/scheduling_flow_matching.py:27-297).

The pipeline keeps using the reference scheduler object unchanged; this mirror exists so that the sampler loop can be
exercised (tests, bench e2e) on a box that does not have the reference tree.  Same constructor arguments, attributes
(`start_sigmas`, `end_sigmas`, `ori_start_sigmas`, `timestep_ratios`, `timesteps`, `sigmas`, `config.gamma`) and methods
(`set_timesteps(n, stage, device)`, `step(model_output, timestep, sample)`); tables are float64 numpy like the reference's
`np.linspace` calls (S:90-149, S:179-206); the Euler update is `x_fp32 + (sigma_next - sigma) * v`, cast to v's dtype
(S:278-286).  Pinned against the reference's tables by tests/golden/scheduler.pt.
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import List, Optional, Union

import numpy as np
import torch


class SchedulerOutput:
    def __init__(self, prev_sample):
        self.prev_sample = prev_sample


class B200FlowMatchScheduler:
    order = 1

    def __init__(self, num_train_timesteps: int = 1000, shift: float = 1.0, stages: int = 3,
                 stage_range: Optional[List[float]] = None, gamma: float = 1 / 3):
        stage_range = [0, 1 / 3, 2 / 3, 1] if stage_range is None else stage_range
        self.config = SimpleNamespace(num_train_timesteps=num_train_timesteps, shift=shift, stages=stages,
                                      stage_range=stage_range, gamma=gamma)
        self.gamma = gamma
        self.timestep_ratios, self.timesteps_per_stage, self.sigmas_per_stage = {}, {}, {}
        self.start_sigmas, self.end_sigmas, self.ori_start_sigmas = {}, {}, {}
        self._build_stage_tables()
        self.sigma_min = float(self.sigmas[-1])
        self.sigma_max = float(self.sigmas[0])
        self._step_index = None

    # global (training) schedule, float32 like the reference (S:90-105)
    def _global_schedule(self):
        n = self.config.num_train_timesteps
        t = torch.from_numpy(np.linspace(1, n, n, dtype=np.float32)[::-1].copy())
        sig = t / n
        sig = self.config.shift * sig / (1 + (self.config.shift - 1) * sig)
        self.timesteps = sig * n
        self.sigmas = sig
        self._step_index = None

    def _build_stage_tables(self):
        self._global_schedule()
        c = self.config
        n = c.num_train_timesteps
        dist = []
        for s in range(c.stages):
            i0 = max(int(c.stage_range[s] * n), 0)
            i1 = min(int(c.stage_range[s + 1] * n), n)
            start = self.sigmas[i0].item()
            end = self.sigmas[i1].item() if i1 < n else 0.0
            self.ori_start_sigmas[s] = start
            if s != 0:   # re-noising correction of the stage start (S:125-130)
                ori = 1 - start
                corrected = (1 / (math.sqrt(1 + (1 / c.gamma)) * (1 - ori) + ori)) * ori
                start = 1 - corrected
            dist.append(start - end)
            self.start_sigmas[s], self.end_sigmas[s] = start, end
        tot = sum(dist)
        for s in range(c.stages):
            a = 0.0 if s == 0 else sum(dist[:s]) / tot
            b = 1.0 if s == c.stages - 1 else sum(dist[:s + 1]) / tot
            self.timestep_ratios[s] = (a, b)
        for s in range(c.stages):
            a, b = self.timestep_ratios[s]
            t_max = self.timesteps[int(a * n)]
            t_min = self.timesteps[min(int(b * n), n - 1)]
            self.timesteps_per_stage[s] = torch.from_numpy(np.linspace(t_max.item(), t_min.item(), n + 1)[:-1])
            self.sigmas_per_stage[s] = torch.from_numpy(np.linspace(1, 0, n + 1)[:-1])

    @property
    def step_index(self):
        return self._step_index

    def set_timesteps(self, num_inference_steps: int, stage_index: int, device: Union[str, torch.device, None] = None):
        self.num_inference_steps = num_inference_steps
        self._global_schedule()
        st = self.timesteps_per_stage[stage_index]
        self.timesteps = torch.from_numpy(np.linspace(st[0].item(), st[-1].item(), num_inference_steps)).to(device=device)
        sg = self.sigmas_per_stage[stage_index]
        sig = torch.from_numpy(np.linspace(sg[0].item(), sg[-1].item(), num_inference_steps)).to(device=device)
        self.sigmas = torch.cat([sig, torch.zeros(1, device=sig.device, dtype=sig.dtype)])
        self._step_index = None

    def step(self, model_output: torch.Tensor, timestep, sample: torch.Tensor, generator=None, return_dict: bool = True):
        if isinstance(timestep, int) or (isinstance(timestep, torch.Tensor) and not timestep.is_floating_point()):
            raise ValueError("Passing integer indices as timesteps to step() is not supported; pass one of scheduler.timesteps")
        if self._step_index is None:
            self._step_index = 0
        sample = sample.to(torch.float32)
        dsigma = self.sigmas[self._step_index + 1] - self.sigmas[self._step_index]
        prev = (sample + dsigma * model_output).to(model_output.dtype)
        self._step_index += 1
        return SchedulerOutput(prev) if return_dict else (prev,)

    def delta_sigma(self) -> float:
        """sigma_{i+1} - sigma_i of the NEXT step (used by the fused CFG+Euler kernel path)."""
        i = 0 if self._step_index is None else self._step_index
        return float(self.sigmas[i + 1] - self.sigmas[i])

    def advance(self) -> None:
        self._step_index = (0 if self._step_index is None else self._step_index) + 1

    def __len__(self):
        return self.config.num_train_timesteps
