"""Drop-in DDPMScheduler / DDIMScheduler for the BadDiffusion sampling and training path.

Same config surface, attributes and `step` / `add_noise` / `set_timesteps` semantics as the (patched) vendored
diffusers schedulers (D/schedulers/scheduling_ddpm.py incl. the `clip_defense` patch :137-138,:414-415, and
D/schedulers/scheduling_ddim.py).  The per-step *scalars* are computed on the host with the reference's own
0-d fp32 torch expressions (so they are bit-identical); the tensor arithmetic runs in one fused CUDA kernel
(bd_ddpm_step / bd_ddim_step) in the reference's association order.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import List, Optional, Union

import numpy as np
import torch

from . import config_utils as CU

SCHEDULER_CONFIG_NAME = "scheduler_config.json"  # D/schedulers/scheduling_utils.py:25


@dataclass
class SchedulerOutput:
    prev_sample: torch.Tensor
    pred_original_sample: Optional[torch.Tensor] = None


def _betas(num_train_timesteps, beta_start, beta_end, beta_schedule, trained_betas):
    if trained_betas is not None:
        return torch.tensor(trained_betas, dtype=torch.float32)
    if beta_schedule == "linear":
        return torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
    if beta_schedule == "scaled_linear":
        return torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
    raise NotImplementedError(f"{beta_schedule} does is not implemented for this scheduler")


def ddpm_coef_row(acp, t, num_inference_steps, num_train_timesteps, variance_type, clip_sample, clip_range,
                  clip_defense_range, prev_t=None) -> torch.Tensor:
    """{sqrt(beta_prod_t), sqrt(alpha_prod_t), c0, ct, sigma, clip, clip_defense, has_noise} exactly as
    scheduling_ddpm.py:352-411 evaluates them (0-d fp32 CPU tensors)."""
    t = int(t)
    if prev_t is None:
        prev_t = t - num_train_timesteps // num_inference_steps
    one = torch.tensor(1.0)
    ap_t = acp[t]
    ap_prev = acp[prev_t] if prev_t >= 0 else one
    bp_t = 1 - ap_t
    bp_prev = 1 - ap_prev
    cur_a = ap_t / ap_prev
    cur_b = 1 - cur_a
    c0 = (ap_prev ** 0.5 * cur_b) / bp_t
    ct = cur_a ** 0.5 * bp_prev / bp_t
    sigma = torch.tensor(0.0)
    if t > 0:
        var = torch.clamp((1 - ap_prev) / (1 - ap_t) * cur_b, min=1e-20)  # _get_variance :250-288
        if variance_type == "fixed_large":
            var = cur_b
        elif variance_type != "fixed_small":
            raise NotImplementedError(f"variance_type {variance_type} is not supported by the fused DDPM step")
        sigma = var ** 0.5
    row = torch.stack([bp_t ** 0.5, ap_t ** 0.5, c0, ct, sigma, torch.tensor(float(clip_range) if clip_sample else 0.0),
                       torch.tensor(float(clip_defense_range)), torch.tensor(1.0 if t > 0 else 0.0)])
    return row.to(torch.float32)


def ddim_coef_row(acp, t, num_inference_steps, num_train_timesteps, eta, clip_sample, clip_range, set_alpha_to_one,
                  use_clipped_model_output) -> torch.Tensor:
    """{sqrt(beta_prod_t), sqrt(alpha_prod_t), sqrt(alpha_prod_prev), (1-a_prev-std^2)^0.5, std, clip, reclip, 0}
    as in scheduling_ddim.py:315-367."""
    t = int(t)
    prev_t = t - num_train_timesteps // num_inference_steps
    final = torch.tensor(1.0) if set_alpha_to_one else acp[0]
    ap_t = acp[t]
    ap_prev = acp[prev_t] if prev_t >= 0 else final
    bp_t = 1 - ap_t
    bp_prev = 1 - ap_prev
    variance = (bp_prev / bp_t) * (1 - ap_t / ap_prev)
    std = eta * variance ** 0.5
    dirc = (1 - ap_prev - std ** 2) ** 0.5
    row = torch.stack([bp_t ** 0.5, ap_t ** 0.5, ap_prev ** 0.5, dirc, std.to(torch.float32) if torch.is_tensor(std) else torch.tensor(std),
                       torch.tensor(float(clip_range) if clip_sample else 0.0),
                       torch.tensor(1.0 if use_clipped_model_output else 0.0), torch.tensor(0.0)])
    return row.to(torch.float32)


def _draw_noise(shape, generator, device):
    """D/utils/torch_utils.py:29-70 (randn_tensor): a CPU generator draws on the CPU and the result is copied to
    the device (quirk Q11) -- this is what makes 'identical seeds' reproduce the reference's noise stream."""
    if generator is not None and generator.device.type == "cpu":
        return torch.randn(shape, generator=generator, dtype=torch.float32).to(device, non_blocking=True)
    return None  # -> in-kernel Philox


class _SchedulerBase:
    config_name = SCHEDULER_CONFIG_NAME
    order = 1

    def save_pretrained(self, save_directory, **kwargs):
        CU.save_config(self.config, type(self).__name__, save_directory, SCHEDULER_CONFIG_NAME)

    save_config = save_pretrained

    @classmethod
    def from_config(cls, config, **kwargs):
        d = dict(config)
        d.update(kwargs)
        return cls(**CU.filter_init_kwargs(cls, d))

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, subfolder=None, **kwargs):
        d = pretrained_model_name_or_path if subfolder is None else os.path.join(pretrained_model_name_or_path, subfolder)
        return cls.from_config(CU.load_config(d, SCHEDULER_CONFIG_NAME), **kwargs)

    def scale_model_input(self, sample, timestep=None):
        return sample

    def add_noise(self, original_samples, noise, timesteps):
        """scheduling_ddpm.py:422-443 through the fused batch-prep kernel with R = 0 (bit-exact)."""
        from . import ops

        if original_samples.is_cuda:
            dev = original_samples.device
            x0 = original_samples.float().contiguous()
            t = timesteps.to(dev).long().contiguous()
            self._to_device(dev)
            noisy, _ = ops.batch_prep(x0, None, None, None, t, self._alphas_dev, self._acp_dev, noise=noise.float().contiguous())
            return noisy.to(original_samples.dtype)
        acp = self.alphas_cumprod.to(dtype=original_samples.dtype)
        a = (acp[timesteps] ** 0.5).flatten()
        s = ((1 - acp[timesteps]) ** 0.5).flatten()
        while a.dim() < original_samples.dim():
            a, s = a.unsqueeze(-1), s.unsqueeze(-1)
        return a * original_samples + s * noise

    def _to_device(self, dev):
        if getattr(self, "_acp_dev", None) is None or self._acp_dev.device != dev:
            self._acp_dev = self.alphas_cumprod.to(dev)
            self._alphas_dev = self.alphas.to(dev)

    def __len__(self):
        return self.config.num_train_timesteps

    @property
    def num_train_timesteps(self):  # deprecated attribute access used at baddiffusion.py:600
        return self.config.num_train_timesteps


class DDPMScheduler(_SchedulerBase):
    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.0001, beta_end: float = 0.02,
                 beta_schedule: str = "linear", trained_betas: Optional[Union[np.ndarray, List[float]]] = None,
                 variance_type: str = "fixed_small", clip_sample: bool = True, prediction_type: str = "epsilon",
                 thresholding: bool = False, dynamic_thresholding_ratio: float = 0.995, clip_sample_range: float = 1.0,
                 sample_max_value: float = 1.0, clip_defense: bool = False, clip_defense_range: float = 1.0):
        CU.capture_init_args(self, DDPMScheduler.__init__, (), dict(
            num_train_timesteps=num_train_timesteps, beta_start=beta_start, beta_end=beta_end,
            beta_schedule=beta_schedule, trained_betas=trained_betas, variance_type=variance_type,
            clip_sample=clip_sample, prediction_type=prediction_type, thresholding=thresholding,
            dynamic_thresholding_ratio=dynamic_thresholding_ratio, clip_sample_range=clip_sample_range,
            sample_max_value=sample_max_value, clip_defense=clip_defense, clip_defense_range=clip_defense_range))
        self.betas = _betas(num_train_timesteps, beta_start, beta_end, beta_schedule, trained_betas)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.one = torch.tensor(1.0)
        self.init_noise_sigma = 1.0
        self.custom_timesteps = False
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy())
        self.variance_type = variance_type
        self._philox_calls = 0

    def set_timesteps(self, num_inference_steps: Optional[int] = None, device=None, timesteps: Optional[List[int]] = None):
        """scheduling_ddpm.py:197-248."""
        if num_inference_steps is not None and timesteps is not None:
            raise ValueError("Can only pass one of `num_inference_steps` or `custom_timesteps`.")
        if timesteps is not None:
            for i in range(1, len(timesteps)):
                if timesteps[i] >= timesteps[i - 1]:
                    raise ValueError("`custom_timesteps` must be in descending order.")
            if timesteps[0] >= self.config.num_train_timesteps:
                raise ValueError(f"`timesteps` must start before `self.config.train_timesteps`: {self.config.num_train_timesteps}.")
            ts = np.array(timesteps, dtype=np.int64)
            self.custom_timesteps = True
        else:
            if num_inference_steps > self.config.num_train_timesteps:
                raise ValueError(
                    f"`num_inference_steps`: {num_inference_steps} cannot be larger than `self.config.train_timesteps`:"
                    f" {self.config.num_train_timesteps} as the unet model trained with this scheduler can only handle"
                    f" maximal {self.config.num_train_timesteps} timesteps.")
            self.num_inference_steps = num_inference_steps
            ratio = self.config.num_train_timesteps // self.num_inference_steps
            ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64)
            self.custom_timesteps = False
        self.timesteps = torch.from_numpy(ts).to(device)

    def previous_timestep(self, timestep):
        """scheduling_ddpm.py:468-481."""
        if self.custom_timesteps:
            index = (self.timesteps == timestep).nonzero(as_tuple=True)[0][0]
            return torch.tensor(-1) if index == self.timesteps.shape[0] - 1 else self.timesteps[index + 1]
        n = self.num_inference_steps if self.num_inference_steps else self.config.num_train_timesteps
        return timestep - self.config.num_train_timesteps // n

    def _check_supported(self):
        c = self.config
        if c.prediction_type != "epsilon" or c.thresholding:
            raise NotImplementedError("the fused DDPM step implements prediction_type='epsilon' without thresholding "
                                      "(the only configuration BadDiffusion uses)")

    def coef_row(self, t) -> torch.Tensor:
        c = self.config
        n = self.num_inference_steps if self.num_inference_steps else c.num_train_timesteps
        prev_t = int(self.previous_timestep(int(t))) if self.custom_timesteps else None
        return ddpm_coef_row(self.alphas_cumprod, int(t), n, c.num_train_timesteps, self.variance_type, c.clip_sample,
                             c.clip_sample_range, c.clip_defense_range if c.clip_defense else 0.0, prev_t=prev_t)

    def coef_table(self, timesteps=None) -> torch.Tensor:
        ts = self.timesteps if timesteps is None else timesteps
        return torch.stack([self.coef_row(int(t)) for t in ts])

    def step(self, model_output, timestep, sample, generator=None, return_dict: bool = True):
        """scheduling_ddpm.py:324-420."""
        from . import ops

        self._check_supported()
        t = int(timestep)
        row = self.coef_row(t).to(sample.device)
        x = sample.float().contiguous()
        eps = model_output.float().contiguous()
        z = _draw_noise(eps.shape, generator, eps.device) if t > 0 else None
        out = torch.empty_like(x)
        seed = generator.initial_seed() if generator is not None else torch.initial_seed()
        self._philox_calls += 1
        ops.ddpm_step(x, eps, z, out, row, None, seed=seed & ((1 << 63) - 1), offset=self._philox_calls)
        out = out.to(sample.dtype)
        if not return_dict:
            return (out,)
        return SchedulerOutput(prev_sample=out)


class DDIMScheduler(_SchedulerBase):
    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.0001, beta_end: float = 0.02,
                 beta_schedule: str = "linear", trained_betas: Optional[Union[np.ndarray, List[float]]] = None,
                 clip_sample: bool = True, set_alpha_to_one: bool = True, steps_offset: int = 0,
                 prediction_type: str = "epsilon", thresholding: bool = False,
                 dynamic_thresholding_ratio: float = 0.995, clip_sample_range: float = 1.0, sample_max_value: float = 1.0):
        CU.capture_init_args(self, DDIMScheduler.__init__, (), dict(
            num_train_timesteps=num_train_timesteps, beta_start=beta_start, beta_end=beta_end,
            beta_schedule=beta_schedule, trained_betas=trained_betas, clip_sample=clip_sample,
            set_alpha_to_one=set_alpha_to_one, steps_offset=steps_offset, prediction_type=prediction_type,
            thresholding=thresholding, dynamic_thresholding_ratio=dynamic_thresholding_ratio,
            clip_sample_range=clip_sample_range, sample_max_value=sample_max_value))
        self.betas = _betas(num_train_timesteps, beta_start, beta_end, beta_schedule, trained_betas)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.init_noise_sigma = 1.0
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64))
        self._philox_calls = 0

    def set_timesteps(self, num_inference_steps: int, device=None):
        """scheduling_ddim.py:237-259."""
        if num_inference_steps > self.config.num_train_timesteps:
            raise ValueError(
                f"`num_inference_steps`: {num_inference_steps} cannot be larger than `self.config.train_timesteps`:"
                f" {self.config.num_train_timesteps} as the unet model trained with this scheduler can only handle"
                f" maximal {self.config.num_train_timesteps} timesteps.")
        self.num_inference_steps = num_inference_steps
        ratio = self.config.num_train_timesteps // self.num_inference_steps
        ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64)
        self.timesteps = torch.from_numpy(ts).to(device)
        self.timesteps += self.config.steps_offset

    def coef_row(self, t, eta=0.0, use_clipped_model_output=False) -> torch.Tensor:
        c = self.config
        return ddim_coef_row(self.alphas_cumprod, int(t), self.num_inference_steps, c.num_train_timesteps, eta,
                             c.clip_sample, c.clip_sample_range, c.set_alpha_to_one, bool(use_clipped_model_output))

    def coef_table(self, eta=0.0, use_clipped_model_output=False, timesteps=None) -> torch.Tensor:
        ts = self.timesteps if timesteps is None else timesteps
        return torch.stack([self.coef_row(int(t), eta, use_clipped_model_output) for t in ts])

    def step(self, model_output, timestep, sample, eta: float = 0.0, use_clipped_model_output: bool = False,
             generator=None, variance_noise=None, return_dict: bool = True):
        """scheduling_ddim.py:261-381."""
        from . import ops

        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' after creating the scheduler")
        c = self.config
        if c.prediction_type != "epsilon" or c.thresholding:
            raise NotImplementedError("the fused DDIM step implements prediction_type='epsilon' without thresholding")
        if eta > 0 and variance_noise is not None and generator is not None:
            raise ValueError("Cannot pass both generator and variance_noise. Please make sure that either `generator` or"
                             " `variance_noise` stays `None`.")
        row = self.coef_row(int(timestep), eta, use_clipped_model_output).to(sample.device)
        x = sample.float().contiguous()
        eps = model_output.float().contiguous()
        z = None
        if eta > 0:
            z = variance_noise.float().contiguous() if variance_noise is not None else _draw_noise(eps.shape, generator, eps.device)
        out = torch.empty_like(x)
        seed = generator.initial_seed() if generator is not None else torch.initial_seed()
        self._philox_calls += 1
        ops.ddim_step(x, eps, z, out, row, None, seed=seed & ((1 << 63) - 1), offset=self._philox_calls)
        out = out.to(sample.dtype)
        if not return_dict:
            return (out,)
        return SchedulerOutput(prev_sample=out)


# mode ids of bd_pndm_step (csrc/elementwise.cu)
PNDM_PRK0, PNDM_PRK12, PNDM_PRK3, PNDM_PLMS_FIRST, PNDM_PLMS_SECOND, PNDM_PLMS2, PNDM_PLMS3, PNDM_PLMS4 = range(8)


class PNDMScheduler(_SchedulerBase):
    """D/schedulers/scheduling_pndm.py.  The sampler behind every `--sched` other than DDPM / DDIM: model.py:598-630 hands
    DPM-Solver / UniPC / DEIS / Heun / LMSD / PNDM schedulers to the reference's patched `PNDMPipeline`, whose constructor
    rebuilds a PNDMScheduler from their config (pipeline_pndm.py:43).  The bookkeeping of `step_prk` / `step_plms`
    (counter, the `ets` history, `cur_sample`) stays on the host as slot indices; the tensor arithmetic of a step is ONE
    kernel (`bd_pndm_step`) driven by a 16-float row built here with the reference's 0-d fp32 torch expressions."""

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.0001, beta_end: float = 0.02,
                 beta_schedule: str = "linear", trained_betas: Optional[Union[np.ndarray, List[float]]] = None,
                 skip_prk_steps: bool = False, set_alpha_to_one: bool = False, prediction_type: str = "epsilon",
                 steps_offset: int = 0):
        CU.capture_init_args(self, PNDMScheduler.__init__, (), dict(
            num_train_timesteps=num_train_timesteps, beta_start=beta_start, beta_end=beta_end, beta_schedule=beta_schedule,
            trained_betas=trained_betas, skip_prk_steps=skip_prk_steps, set_alpha_to_one=set_alpha_to_one,
            prediction_type=prediction_type, steps_offset=steps_offset))
        self.betas = _betas(num_train_timesteps, beta_start, beta_end, beta_schedule, trained_betas)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.init_noise_sigma = 1.0
        self.pndm_order = 4
        self.num_inference_steps = None
        self._timesteps = np.arange(0, num_train_timesteps)[::-1].copy()
        self.prk_timesteps = self.plms_timesteps = self.timesteps = None
        self._reset()
        self._state = None   # device state of the eager `step` API

    def _reset(self):
        self.counter = 0
        self._slots = []     # history slot ids of ets, oldest first

    def set_timesteps(self, num_inference_steps: int, device=None):
        """scheduling_pndm.py:151-190."""
        self.num_inference_steps = num_inference_steps
        step_ratio = self.config.num_train_timesteps // self.num_inference_steps
        self._timesteps = (np.arange(0, num_inference_steps) * step_ratio).round()
        self._timesteps += self.config.steps_offset
        if self.config.skip_prk_steps:
            self.prk_timesteps = np.array([])
            self.plms_timesteps = np.concatenate([self._timesteps[:-1], self._timesteps[-2:-1], self._timesteps[-1:]])[::-1].copy()
        else:
            prk = np.array(self._timesteps[-self.pndm_order:]).repeat(2) + np.tile(
                np.array([0, self.config.num_train_timesteps // num_inference_steps // 2]), self.pndm_order)
            self.prk_timesteps = (prk[:-1].repeat(2)[1:-1])[::-1].copy()
            self.plms_timesteps = self._timesteps[:-3][::-1].copy()
        timesteps = np.concatenate([self.prk_timesteps, self.plms_timesteps]).astype(np.int64)
        self.timesteps = torch.from_numpy(timesteps).to(device)
        self._reset()

    def _prev_sample_coefs(self, timestep: int, prev_timestep: int):
        """_get_prev_sample :375-393 (scalars only)."""
        if self.config.prediction_type != "epsilon":
            raise NotImplementedError("the fused PNDM step implements prediction_type='epsilon' (what BadDiffusion uses)")
        a_t = self.alphas_cumprod[timestep]
        a_prev = self.alphas_cumprod[prev_timestep] if prev_timestep >= 0 else self.final_alpha_cumprod
        b_t = 1 - a_t
        b_prev = 1 - a_prev
        sample_coeff = (a_prev / a_t) ** (0.5)
        denom = a_t * b_prev ** (0.5) + (a_t * b_t * a_prev) ** (0.5)
        return sample_coeff, a_prev - a_t, denom

    def _free_slot(self) -> int:
        return next(s for s in range(4) if s not in self._slots)

    def next_row(self, timestep: int, clip: float = 0.0) -> torch.Tensor:
        """Advances the host bookkeeping by one `step` call (:191-340) and returns that step's coefficient row."""
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' after creating the scheduler")
        timestep = int(timestep)
        ratio = self.config.num_train_timesteps // self.num_inference_steps
        push = -1
        if self.counter < len(self.prk_timesteps) and not self.config.skip_prk_steps:
            diff_to_prev = 0 if self.counter % 2 else ratio // 2
            prev_timestep = timestep - diff_to_prev
            timestep = int(self.prk_timesteps[self.counter // 4 * 4])
            k = self.counter % 4
            if k == 0:
                mode, push = PNDM_PRK0, self._free_slot()
                self._slots.append(push)
            else:
                mode = PNDM_PRK12 if k < 3 else PNDM_PRK3
        else:
            if not self.config.skip_prk_steps and len(self._slots) < 3:
                raise ValueError(f"{self.__class__} can only be run AFTER scheduler has been run in 'prk' mode for at least 12 iterations")
            prev_timestep = timestep - ratio
            if self.counter != 1:
                self._slots = self._slots[-3:]
                push = self._free_slot()
                self._slots.append(push)
            else:
                prev_timestep = timestep
                timestep = timestep + ratio
            n = len(self._slots)
            if n == 1 and self.counter == 0:
                mode = PNDM_PLMS_FIRST
            elif n == 1 and self.counter == 1:
                mode = PNDM_PLMS_SECOND
            elif n == 2:
                mode = PNDM_PLMS2
            elif n == 3:
                mode = PNDM_PLMS3
            else:
                mode = PNDM_PLMS4
        cs, dd, den = self._prev_sample_coefs(timestep, prev_timestep)
        self.counter += 1
        hist = [float(self._slots[-k]) if len(self._slots) >= k else -1.0 for k in (1, 2, 3, 4)]
        row = torch.zeros(16, dtype=torch.float32)
        row[0], row[1], row[2], row[3], row[4], row[5] = float(mode), cs, dd, den, float(clip), float(push)
        row[6:10] = torch.tensor(hist)
        return row

    def coef_table(self, timesteps, clip: float = 0.0) -> torch.Tensor:
        """Rows for a whole sampling loop over `timesteps` (fresh bookkeeping, as right after set_timesteps)."""
        self._reset()
        return torch.stack([self.next_row(int(t), clip) for t in timesteps])

    def step(self, model_output, timestep, sample, return_dict: bool = True):
        """scheduling_pndm.py:191-340 (eager API: one kernel per call, state kept on the sample's device)."""
        from . import ops

        x = sample.float().contiguous()
        eps = model_output.float().contiguous()
        if self._state is None or self._state.numel() != 6 * x.numel() or self._state.device != x.device or self.counter == 0:
            self._state = torch.zeros(6 * x.numel(), device=x.device)
        row = self.next_row(int(timestep)).to(x.device)
        out = torch.empty_like(x)
        ops.pndm_step(x, eps, out, self._state, row)
        out = out.to(sample.dtype)
        if not return_dict:
            return (out,)
        return SchedulerOutput(prev_sample=out)
