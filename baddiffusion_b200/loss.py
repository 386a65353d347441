"""Drop-in replacements for the reference's loss.py (the backdoored forward process and training loss).

`q_sample_diffuser` (loss.py:257-285) runs as ONE fused kernel (bd_batch_prep, explicit-R form) that is
bit-identical to the reference's ~20 elementwise torch ops; `p_losses_diffuser` (loss.py:287-307) returns a scalar
tensor that supports `.backward()` like the reference's (the MSE and its gradient are one kernel pair).
"""
from __future__ import annotations

import torch

from . import ops


def _tables(noise_sched, device):
    noise_sched._to_device(device) if hasattr(noise_sched, "_to_device") else None
    if hasattr(noise_sched, "_acp_dev") and noise_sched._acp_dev is not None and noise_sched._acp_dev.device == device:
        return noise_sched._alphas_dev, noise_sched._acp_dev
    return (noise_sched.alphas.to(device=device, dtype=torch.float32).contiguous(),
            noise_sched.alphas_cumprod.to(device=device, dtype=torch.float32).contiguous())


def q_sample_diffuser(noise_sched, x_start: torch.Tensor, R: torch.Tensor, timesteps: torch.Tensor,
                      noise: torch.Tensor = None):
    """Returns (x_noisy, target) = (add_noise(x0, eps, t) + (1 - sqrt(acp_t)) R,  rho_t R + eps)."""
    if not x_start.is_cuda:
        raise RuntimeError("baddiffusion_b200.q_sample_diffuser runs on CUDA (sm_100a) only")
    dev = x_start.device
    alphas, acp = _tables(noise_sched, dev)
    x0 = x_start.to(torch.float32).contiguous()
    Rc = R.to(device=dev, dtype=torch.float32).contiguous()
    t = timesteps.to(dev).long().contiguous()
    if noise is None:
        seed = torch.initial_seed() & ((1 << 63) - 1)
        q_sample_diffuser._calls = getattr(q_sample_diffuser, "_calls", 0) + 1
        x_noisy, target = ops.batch_prep(x0, None, None, None, t, alphas, acp, noise=None, R=Rc, seed=seed,
                                         offset=q_sample_diffuser._calls)
    else:
        x_noisy, target = ops.batch_prep(x0, None, None, None, t, alphas, acp,
                                         noise=noise.to(device=dev, dtype=torch.float32).contiguous(), R=Rc)
    return x_noisy, target


class _MSE(torch.autograd.Function):
    """F.mse_loss(target, eps_hat) (loss.py:301) with the fused forward+gradient kernel."""

    @staticmethod
    def forward(ctx, eps_hat, target):
        loss = torch.empty(1, device=eps_hat.device)
        grad = torch.empty_like(eps_hat)
        partial = torch.empty(1024, device=eps_hat.device)
        ops.mse_fwd_bwd(eps_hat.contiguous(), target.contiguous(), loss, grad, partial, None)
        ctx.save_for_backward(grad)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g, None


def p_losses_diffuser(noise_sched, model, x_start: torch.Tensor, R: torch.Tensor, timesteps: torch.Tensor,
                      noise: torch.Tensor = None, loss_type: str = "l2") -> torch.Tensor:
    if len(x_start) == 0:
        return 0
    if loss_type != "l2":
        if loss_type in ("l1", "huber"):
            raise NotImplementedError("only loss_type='l2' is on the BadDiffusion hot path (baddiffusion.py:607)")
        raise NotImplementedError()
    x_noisy, target = q_sample_diffuser(noise_sched=noise_sched, x_start=x_start, R=R, timesteps=timesteps, noise=noise)
    predicted_noise = model(x_noisy.contiguous(), timesteps.contiguous(), return_dict=False)[0]
    return _MSE.apply(predicted_noise, target)
