"""The poisoned DDPM training step of baddiffusion.py:593-615 as one device-resident kernel sequence:

    batch-prep (poison blend + add_noise + loss target, K11) -> UNet forward -> MSE + d(eps_hat) (K12)
    -> UNet backward -> [NCCL all-reduce of the flat fp32 gradient buffer] -> global-norm -> clip + Adam -> scaler

captured in CUDA graphs (the only host work per step is the H2D copy of the batch and the graph launches).
Mixed precision follows the reference's fp16 autocast + GradScaler recipe (baddiffusion.py:116,608): fp16 operands,
fp32 accumulation and master weights, dynamic loss scale with skip-on-overflow.

Multi-GPU: one process per GPU, each rank owns B/world samples, ONE all-reduce (average) over the
flat gradient buffer per step; clip-by-global-norm and Adam then run identically on every rank.
"""
from __future__ import annotations

import math
import os
from typing import Optional

import torch

from . import ops
from .schedulers import DDPMScheduler
from .unet import UNet2DModel


def cosine_lr_lambda(step: int, warmup: int, total: int, num_cycles: float = 0.5) -> float:
    """D/optimization.py:134-138 (get_cosine_schedule_with_warmup)."""
    if step < warmup:
        return float(step) / float(max(1, warmup))
    progress = float(step - warmup) / float(max(1, total - warmup))
    return max(0.0, 0.5 * (1.0 + math.cos(math.pi * float(num_cycles) * 2.0 * progress)))


def uncovered_ranges(covered, n):
    """[lo, hi) ranges of [0, n) that no range of `covered` touches (the gradient bytes left for the last all-reduce)."""
    out, pos = [], 0
    for lo, hi in sorted(covered) + [(n, n)]:
        if lo > pos:
            out.append((pos, lo))
        pos = max(pos, hi)
    return out


class Trainer:
    def __init__(self, model: UNet2DModel, noise_sched: DDPMScheduler, batch: int, trigger: torch.Tensor,
                 target: torch.Tensor, lr: float = 2e-4, total_steps: int = 23450, warmup_steps: int = 500,
                 max_grad_norm: float = 1.0, betas=(0.9, 0.999), eps: float = 1e-8, init_scale: float = 65536.0,
                 growth_interval: int = 2000, use_graph: Optional[bool] = None, process_group=None, seed: int = 0,
                 lr_table_len: int = 0, accum_steps: int = 1, u8_input: bool = False):
        if not model.device.type == "cuda":
            raise RuntimeError("Trainer needs the model on a CUDA device")
        self.model, self.sched, self.B = model, noise_sched, batch
        dev = model.device
        self.dev = dev
        self.pg = process_group
        self.world = torch.distributed.get_world_size(process_group) if process_group is not None else 1
        self.use_graph = (os.environ.get("BD_NO_GRAPH", "0") != "1") if use_graph is None else use_graph
        cfg = model.config
        S = cfg.sample_size if isinstance(cfg.sample_size, int) else cfg.sample_size[0]
        C = cfg.in_channels
        self.shape = (batch, C, S, S)
        noise_sched._to_device(dev)
        self.alphas, self.acp = noise_sched._alphas_dev, noise_sched._acp_dev
        self.T = noise_sched.config.num_train_timesteps
        self.trigger = trigger.to(dev, torch.float32).contiguous()
        self.target = target.to(dev, torch.float32).contiguous()
        # static device buffers (graph inputs / outputs)
        self.img = torch.zeros(self.shape, device=dev)
        self.isp = torch.zeros(batch, dtype=torch.uint8, device=dev)
        # u8_input: the batch is fed as decoded uint8 NHWC pixels + per-sample h-flip coins (load_batch_u8 / step_u8);
        # the choice is baked into the captured graph like the noise mode
        self.u8_input = bool(u8_input)
        self.img_u8 = torch.zeros((batch, S, S, C), dtype=torch.uint8, device=dev) if u8_input else None
        self.flip = torch.zeros(batch, dtype=torch.uint8, device=dev) if u8_input else None
        self.t = torch.zeros(batch, dtype=torch.int64, device=dev)
        self.noise = torch.zeros(self.shape, device=dev)
        self.x_noisy = torch.empty(self.shape, device=dev)
        self.eps_target = torch.empty(self.shape, device=dev)
        self.d_eps = torch.empty(self.shape, device=dev)
        self.loss = torch.zeros(1, device=dev)
        self.mse_part = torch.empty(1024, device=dev)
        self.norm_part = torch.empty(1024, device=dev)
        # optimizer state on the flat buffers
        self.eng = model.engine(batch, True)
        self.flat = model.flat_params
        self.gflat = model.flat_grads(attach=True)
        self.m = torch.zeros_like(self.flat)
        self.v = torch.zeros_like(self.flat)
        self.state = torch.tensor([init_scale, 0.0, 0.0, 0.0, 0.0], device=dev)  # scale, tracker, found_inf, norm, skipped
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=dev)   # successful optimizer steps
        self.iter_dev = torch.zeros(1, dtype=torch.int32, device=dev)   # iterations (Philox stream counter)
        n_lr = lr_table_len if lr_table_len > 0 else total_steps + 1
        lrs = [lr * cosine_lr_lambda(i, warmup_steps, total_steps) for i in range(n_lr)]
        self.lr_table = torch.tensor(lrs, dtype=torch.float32, device=dev)
        self.max_grad_norm, self.betas, self.eps = max_grad_norm, betas, eps
        self.growth_interval = growth_interval
        self.seed = seed
        self.host_step = 0
        self._g_fb = None
        self._g_fb_acc = None     # same sequence without the gradient memset (micro-steps 2..k of an accumulation window)
        self._g_fb2 = None
        self._g_opt = None
        # gradient accumulation (baddiffusion.py:195-217 + accelerator.accumulate, :603-615): k micro-batches per optimizer
        # step; the loss of each is divided by k (accelerate does that in `backward`) -- here as one scale of the flat
        # gradient buffer in front of the optimizer -- and the all-reduce happens once, after the last micro-batch
        self.accum_steps = int(accum_steps)
        if self.accum_steps < 1:
            raise ValueError("accum_steps must be >= 1")
        self._micro = 0
        # data parallel: all-reduce the up-path gradients (the larger, contiguous half of the buffer) while the rest of
        # backward runs, the remainder at the end (BD_NO_AR_OVERLAP=1: one all-reduce after backward)
        self.overlap_allreduce = (self.world > 1 and self.accum_steps == 1
                                  and os.environ.get("BD_NO_AR_OVERLAP", "0") != "1")
        self.launches_per_step = 0
        # Host batches are copied on their own stream: the static input buffers are only read by the batch-prep kernel
        # at the start of the forward/backward sequence, so the H2D copy of step i+1 may run while the optimizer of step
        # i is still on the GPU; likewise `loss_item()` hands the loss to the host as soon as that sequence is done.
        self._copy_stream = torch.cuda.Stream(device=dev)
        self._d2h_stream = torch.cuda.Stream(device=dev)
        self._ev_seq_done = None                       # recorded after the sequence that reads the inputs / writes the loss
        self._loss_host = torch.zeros(1, dtype=torch.float32).pin_memory()

    def _copy_in(self, pairs):
        """(static device buffer, source) pairs.  Host sources go through the copy stream (ordered after the last reader of
        the buffers, not after the optimizer); device sources may have been produced on the current stream: copied there."""
        cur = torch.cuda.current_stream()
        if any(src.is_cuda for _, src in pairs) or os.environ.get("BD_NO_COPY_STREAM"):
            for dst, src in pairs:
                dst.copy_(src, non_blocking=True)
            return
        cs = self._copy_stream
        if self._ev_seq_done is not None:
            cs.wait_event(self._ev_seq_done)
        else:
            cs.wait_stream(cur)
        with torch.cuda.stream(cs):
            for dst, src in pairs:
                dst.copy_(src, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(cs)
        cur.wait_event(ev)

    def loss_item(self) -> float:
        """The loss of the last step as a Python float.  Waits for the forward/backward sequence only (a plain
        `float(trainer.loss)` also waits for the optimizer, and the next batch's H2D copy behind it)."""
        if self._ev_seq_done is None or os.environ.get("BD_NO_COPY_STREAM"):
            return float(self.loss)
        ds = self._d2h_stream
        ds.wait_event(self._ev_seq_done)
        with torch.cuda.stream(ds):
            self._loss_host.copy_(self.loss, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(ds)
        ev.synchronize()
        return float(self._loss_host)

    # ------------------------------------------------------------------ the kernel sequence
    def _fwd_bwd(self, philox_noise: bool, part: Optional[int] = None, zero: bool = True):
        """part None: the whole sequence; 0: everything up to the end of the first backward part; i > 0: backward part i
        (the data-parallel step all-reduces the finished gradient ranges in between, see UNetEngine.bwd_parts)."""
        eng = self.eng
        if part is not None and part > 0:
            eng.run_backward(part)
            return
        # fp16 shadow of the GEMM weights and the gradient memset run on the side stream under batch-prep and the timestep
        # MLP; the engine joins before its first tensor-core GEMM (f_temb) and this function before backward
        eng._fork(lambda: ops.cast_f32_to_f16(self.flat[: self.model.layout.n_gemm], eng.flat16))
        if zero:
            eng._fork(self.gflat.zero_)
        if self.u8_input:   # decoded uint8 NHWC pixels in, ToTensor / normalize / h-flip fused into batch-prep
            ops.batch_prep_u8(self.img_u8, self.flip, self.isp, self.trigger, self.target, self.t, self.alphas, self.acp,
                              noise=None if philox_noise else self.noise, seed=self.seed, offset=0,
                              x_noisy=self.x_noisy, eps_target=self.eps_target, noise_counter=self.iter_dev)
        else:
            ops.batch_prep(self.img, self.isp, self.trigger, self.target, self.t, self.alphas, self.acp,
                           noise=None if philox_noise else self.noise, seed=self.seed, offset=0,
                           x_noisy=self.x_noisy, eps_target=self.eps_target, noise_out=None, noise_counter=self.iter_dev)
        self.iter_dev.add_(1)
        eng.io["x"], eng.io["t"], eng.io["d_eps"] = self.x_noisy, self.t, self.d_eps
        eng.run_forward()
        ops.mse_fwd_bwd(eng.eps_hat, self.eps_target, self.loss, self.d_eps, self.mse_part, self.state[0:1])
        eng._join()                   # gradient writers on the main stream (timestep path) are ordered after the memset
        eng.run_backward(part)

    def _optimizer(self):
        if self.accum_steps > 1:
            self.gflat.mul_(1.0 / self.accum_steps)
        ops.grad_norm(self.gflat, self.norm_part, self.state)
        ops.adam_step(self.flat, self.gflat, self.m, self.v, self.lr_table, self.step_dev, self.state,
                      beta1=self.betas[0], beta2=self.betas[1], eps=self.eps, max_norm=self.max_grad_norm,
                      lr_len=self.lr_table.numel())
        ops.scaler_update(self.state, self.step_dev, 2.0, 0.5, self.growth_interval)

    def _capture(self, fn):
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            fn()  # warm-up outside capture (function attributes, lazy init)
        torch.cuda.current_stream().wait_stream(s)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        return g

    def _ensure_graphs(self, philox_noise: bool):
        if not self.use_graph or self._g_fb is not None:
            return
        # the warm-up run must not disturb training state: snapshot what the optimizer mutates
        snap = (self.flat.clone(), self.m.clone(), self.v.clone(), self.state.clone(), self.step_dev.clone(),
                self.iter_dev.clone())
        before = ops.launch_count()
        if self.overlap_allreduce:
            self._g_fb = self._capture(lambda: self._fwd_bwd(philox_noise, 0))
            self._g_fb2 = [self._capture(lambda i=i: self._fwd_bwd(philox_noise, i)) for i in range(1, len(self.eng.bwd_parts))]
        else:
            self._g_fb = self._capture(lambda: self._fwd_bwd(philox_noise))
        self._g_opt = self._capture(self._optimizer)
        self.launches_per_step = (ops.launch_count() - before) // 2 + 1  # + the memset of the gradient buffer
        if self.accum_steps > 1:
            self._g_fb_acc = self._capture(lambda: self._fwd_bwd(philox_noise, zero=False))
        for dst, src in zip((self.flat, self.m, self.v, self.state, self.step_dev, self.iter_dev), snap):
            dst.copy_(src)
        self._philox = philox_noise

    # ------------------------------------------------------------------ public API
    def load_batch(self, image: torch.Tensor, is_poison: torch.Tensor, noise: Optional[torch.Tensor] = None,
                   t: Optional[torch.Tensor] = None):
        """Host (pinned) or device tensors -> the static device buffers.  `t` / `noise` default to the reference's
        draws: t = randint on the device (baddiffusion.py:600); noise in-kernel (Philox) unless given."""
        pairs = [(self.img, image), (self.isp, is_poison.to(torch.uint8))]
        if t is not None:
            pairs.append((self.t, t))
        if noise is not None:
            pairs.append((self.noise, noise))
        self._copy_in(pairs)
        if t is None:
            self.t.copy_(torch.randint(0, self.T, (self.B,), device=self.dev))

    def step(self, image: torch.Tensor, is_poison: torch.Tensor, noise: Optional[torch.Tensor] = None,
             t: Optional[torch.Tensor] = None) -> torch.Tensor:
        """One optimisation step on one (local) batch; returns the device loss tensor (no host sync)."""
        if self.u8_input:
            raise RuntimeError("this Trainer was built with u8_input=True: feed it with step_u8()")
        self.load_batch(image, is_poison, noise, t)
        return self.step_resident(philox_noise=noise is None)

    def load_batch_u8(self, image_u8: torch.Tensor, flip: torch.Tensor, is_poison: torch.Tensor,
                      noise: Optional[torch.Tensor] = None, t: Optional[torch.Tensor] = None):
        """Decoded pixels (B, S, S, C) uint8 NHWC + h-flip coins (dataset.draw_flips) -> the static device buffers:
        a quarter of the fp32 batch's H2D bytes, and no CPU-side ToTensor / normalize / flip (dataset.py:120-136)."""
        if not self.u8_input:
            raise RuntimeError("build the Trainer with u8_input=True to feed uint8 batches")
        pairs = [(self.img_u8, image_u8), (self.flip, flip.to(torch.uint8)), (self.isp, is_poison.to(torch.uint8))]
        if t is not None:
            pairs.append((self.t, t))
        if noise is not None:
            pairs.append((self.noise, noise))
        self._copy_in(pairs)
        if t is None:
            self.t.copy_(torch.randint(0, self.T, (self.B,), device=self.dev))

    def step_u8(self, image_u8: torch.Tensor, flip: torch.Tensor, is_poison: torch.Tensor,
                noise: Optional[torch.Tensor] = None, t: Optional[torch.Tensor] = None) -> torch.Tensor:
        self.load_batch_u8(image_u8, flip, is_poison, noise, t)
        return self.step_resident(philox_noise=noise is None)

    def _mark_seq_done(self):
        self._ev_seq_done = torch.cuda.Event()
        self._ev_seq_done.record()

    def step_resident(self, philox_noise: bool = True) -> torch.Tensor:
        """Same, with the batch already in the static device buffers (img / isp / t / noise)."""
        if self.use_graph:
            self._ensure_graphs(philox_noise)
            assert self._philox == philox_noise, "noise mode is baked into the captured graph"
        dist = torch.distributed
        first, last = self._micro == 0, self._micro == self.accum_steps - 1
        self._micro = 0 if last else self._micro + 1
        if self.overlap_allreduce:
            # backward part i+1 runs while NCCL (its own stream) averages the gradient ranges part i finished
            parts = self.eng.bwd_parts
            works, covered = [], []
            for i in range(len(parts)):
                if self.use_graph:
                    (self._g_fb if i == 0 else self._g_fb2[i - 1]).replay()
                else:
                    self._fwd_bwd(philox_noise, i)
                if i == 0:
                    self._mark_seq_done()
                for lo, hi in parts[i][2]:
                    works.append(dist.all_reduce(self.gflat[lo:hi], op=dist.ReduceOp.AVG, group=self.pg, async_op=True))
                    covered.append((lo, hi))
            for lo, hi in uncovered_ranges(covered, self.gflat.numel()):   # whatever no part claimed
                works.append(dist.all_reduce(self.gflat[lo:hi], op=dist.ReduceOp.AVG, group=self.pg, async_op=True))
            for w in works:
                w.wait()
        else:
            if self.use_graph:
                (self._g_fb if first else self._g_fb_acc).replay()
            else:
                self._fwd_bwd(philox_noise, zero=first)
            self._mark_seq_done()
            if not last:
                return self.loss
            if self.world > 1:
                dist.all_reduce(self.gflat, op=dist.ReduceOp.AVG, group=self.pg)
        if self.use_graph:
            self._g_opt.replay()
        else:
            self._optimizer()
        self.host_step += 1
        self.model.mark_params_changed()
        return self.loss

    @property
    def loss_scale(self) -> float:
        return float(self.state[0])

    @property
    def grad_norm(self) -> float:
        return float(self.state[3])
