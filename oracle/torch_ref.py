"""CPU oracle: a plain-PyTorch fp32 restatement of the BadDiffusion hot path.

TEST INFRASTRUCTURE ONLY -- not the product.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import this module; the product path
(`baddiffusion_b200`) never does and fails loudly when its CUDA library is missing.

Every function cites the reference file:line it restates (paths relative to /root/reference,
`D/` = diffusers/src/diffusers/).  The restatement is *functional* (state_dict + config dict in,
tensors out) so that it shares no module structure with either the reference or the product.

Pinning (see tests/test_oracle_vs_golden.py): checked against
  * fixtures generated from the reference's own code in the build container
    (scripts/make_goldens.py -> tests/golden/*.npz), and
  * the reference's hard-coded known-answer values (diffusers/tests/..., restated in the tests).
"""
from __future__ import annotations

import hashlib
import math
from collections import OrderedDict
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------------
# Model configs (architecture of google/ddpm-cifar10-32 and
# google/ddpm-ema-celebahq-256 == model.py:657-679 topology)
# --------------------------------------------------------------------------------------------
UNET_DEFAULTS = dict(  # D/models/unet_2d.py:83-105
    sample_size=None, in_channels=3, out_channels=3, center_input_sample=False,
    time_embedding_type="positional", freq_shift=0, flip_sin_to_cos=True,
    down_block_types=("DownBlock2D", "AttnDownBlock2D", "AttnDownBlock2D", "AttnDownBlock2D"),
    up_block_types=("AttnUpBlock2D", "AttnUpBlock2D", "AttnUpBlock2D", "UpBlock2D"),
    block_out_channels=(224, 448, 672, 896), layers_per_block=2, mid_block_scale_factor=1,
    downsample_padding=1, act_fn="silu", attention_head_dim=8, norm_num_groups=32, norm_eps=1e-5,
    resnet_time_scale_shift="default", add_attention=True, class_embed_type=None, num_class_embeds=None,
)

CIFAR10_CONFIG = dict(
    UNET_DEFAULTS, sample_size=32, block_out_channels=(128, 256, 256, 256),
    down_block_types=("DownBlock2D", "AttnDownBlock2D", "DownBlock2D", "DownBlock2D"),
    up_block_types=("UpBlock2D", "UpBlock2D", "AttnUpBlock2D", "UpBlock2D"),
    attention_head_dim=None, norm_eps=1e-6, downsample_padding=0, flip_sin_to_cos=False, freq_shift=1,
)

CELEBAHQ_CONFIG = dict(
    UNET_DEFAULTS, sample_size=256, block_out_channels=(128, 128, 256, 256, 512, 512),
    down_block_types=("DownBlock2D",) * 4 + ("AttnDownBlock2D", "DownBlock2D"),
    up_block_types=("UpBlock2D", "AttnUpBlock2D") + ("UpBlock2D",) * 4,
    attention_head_dim=None, norm_eps=1e-6, downsample_padding=0, flip_sin_to_cos=False, freq_shift=1,
)

TINY_CONFIG = dict(  # the reference's dummy UNet, T/pipelines/ddpm/test_ddpm.py:30-41
    UNET_DEFAULTS, sample_size=32, block_out_channels=(32, 64), layers_per_block=2,
    down_block_types=("DownBlock2D", "AttnDownBlock2D"), up_block_types=("AttnUpBlock2D", "UpBlock2D"),
)


# --------------------------------------------------------------------------------------------
# Parameter inventory (state_dict keys of SURVEY.md Appendix D) and deterministic weights
# --------------------------------------------------------------------------------------------
def _resnet_shapes(prefix: str, cin: int, cout: int, temb: int, out: "OrderedDict[str, tuple]"):
    out[prefix + "norm1.weight"] = (cin,)
    out[prefix + "norm1.bias"] = (cin,)
    out[prefix + "conv1.weight"] = (cout, cin, 3, 3)
    out[prefix + "conv1.bias"] = (cout,)
    out[prefix + "time_emb_proj.weight"] = (cout, temb)
    out[prefix + "time_emb_proj.bias"] = (cout,)
    out[prefix + "norm2.weight"] = (cout,)
    out[prefix + "norm2.bias"] = (cout,)
    out[prefix + "conv2.weight"] = (cout, cout, 3, 3)
    out[prefix + "conv2.bias"] = (cout,)
    if cin != cout:  # D/models/resnet.py:541-549
        out[prefix + "conv_shortcut.weight"] = (cout, cin, 1, 1)
        out[prefix + "conv_shortcut.bias"] = (cout,)


def _attn_shapes(prefix: str, c: int, out: "OrderedDict[str, tuple]"):
    out[prefix + "group_norm.weight"] = (c,)
    out[prefix + "group_norm.bias"] = (c,)
    for n in ("query", "key", "value", "proj_attn"):
        out[prefix + n + ".weight"] = (c, c)
        out[prefix + n + ".bias"] = (c,)


def unet_topology(cfg: dict) -> dict:
    """Walks the UNet2DModel constructor (D/models/unet_2d.py:82-217) and returns the block list."""
    boc = list(cfg["block_out_channels"])
    temb = boc[0] * 4
    L = cfg["layers_per_block"]
    down = []
    out_c = boc[0]
    for i, typ in enumerate(cfg["down_block_types"]):
        in_c, out_c = out_c, boc[i]
        final = i == len(boc) - 1
        down.append(dict(type=typ, resnets=[(in_c if j == 0 else out_c, out_c) for j in range(L)],
                         attn=typ == "AttnDownBlock2D", down=not final, channels=out_c))
    rev = list(reversed(boc))
    up = []
    out_c = rev[0]
    for i, typ in enumerate(cfg["up_block_types"]):
        prev_out = out_c
        out_c = rev[i]
        in_c = rev[min(i + 1, len(boc) - 1)]
        final = i == len(boc) - 1
        res = []
        for j in range(L + 1):  # D/models/unet_2d_blocks.py:1893-1896
            skip_c = in_c if j == L else out_c
            res_in = prev_out if j == 0 else out_c
            res.append((res_in + skip_c, out_c, skip_c))
        up.append(dict(type=typ, resnets=res, attn=typ == "AttnUpBlock2D", up=not final, channels=out_c))
    return dict(temb=temb, down=down, up=up, mid=boc[-1])


def unet_param_shapes(cfg: dict) -> "OrderedDict[str, tuple]":
    topo = unet_topology(cfg)
    boc = list(cfg["block_out_channels"])
    temb = topo["temb"]
    out: "OrderedDict[str, tuple]" = OrderedDict()
    out["conv_in.weight"] = (boc[0], cfg["in_channels"], 3, 3)
    out["conv_in.bias"] = (boc[0],)
    out["time_embedding.linear_1.weight"] = (temb, boc[0])
    out["time_embedding.linear_1.bias"] = (temb,)
    out["time_embedding.linear_2.weight"] = (temb, temb)
    out["time_embedding.linear_2.bias"] = (temb,)
    for i, b in enumerate(topo["down"]):
        if b["attn"]:
            for j in range(len(b["resnets"])):
                _attn_shapes(f"down_blocks.{i}.attentions.{j}.", b["channels"], out)
        for j, (ci, co) in enumerate(b["resnets"]):
            _resnet_shapes(f"down_blocks.{i}.resnets.{j}.", ci, co, temb, out)
        if b["down"]:
            out[f"down_blocks.{i}.downsamplers.0.conv.weight"] = (b["channels"], b["channels"], 3, 3)
            out[f"down_blocks.{i}.downsamplers.0.conv.bias"] = (b["channels"],)
    for i, b in enumerate(topo["up"]):
        if b["attn"]:
            for j in range(len(b["resnets"])):
                _attn_shapes(f"up_blocks.{i}.attentions.{j}.", b["channels"], out)
        for j, (ci, co, _) in enumerate(b["resnets"]):
            _resnet_shapes(f"up_blocks.{i}.resnets.{j}.", ci, co, temb, out)
        if b["up"]:
            out[f"up_blocks.{i}.upsamplers.0.conv.weight"] = (b["channels"], b["channels"], 3, 3)
            out[f"up_blocks.{i}.upsamplers.0.conv.bias"] = (b["channels"],)
    if cfg.get("add_attention", True):
        _attn_shapes("mid_block.attentions.0.", topo["mid"], out)
    _resnet_shapes("mid_block.resnets.0.", topo["mid"], topo["mid"], temb, out)
    _resnet_shapes("mid_block.resnets.1.", topo["mid"], topo["mid"], temb, out)
    out["conv_norm_out.weight"] = (boc[0],)
    out["conv_norm_out.bias"] = (boc[0],)
    out["conv_out.weight"] = (cfg["out_channels"], boc[0], 3, 3)
    out["conv_out.bias"] = (cfg["out_channels"],)
    return out


def make_state_dict(cfg: dict, seed: int = 0, dtype=torch.float32) -> "OrderedDict[str, torch.Tensor]":
    """Synthetic checkpoint: one independent CPU generator per key (seeded from sha256(key, seed)),
    so the values do not depend on module construction order.  Scales keep activations O(1)."""
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for k, shp in unet_param_shapes(cfg).items():
        h = int.from_bytes(hashlib.sha256(f"{seed}:{k}".encode()).digest()[:8], "little") & ((1 << 63) - 1)
        g = torch.Generator().manual_seed(h)
        r = torch.randn(shp, generator=g, dtype=torch.float32)
        is_norm = ("norm" in k.split(".")[-2])
        if k.endswith(".bias"):
            v = 0.1 * r if is_norm else 0.05 * r
        elif is_norm:
            v = 1.0 + 0.1 * r
        else:
            fan_in = int(np.prod(shp[1:]))
            v = r / math.sqrt(fan_in)
        sd[k] = v.to(dtype)
    return sd


# --------------------------------------------------------------------------------------------
# UNet forward
# --------------------------------------------------------------------------------------------
def timestep_embedding(t: torch.Tensor, dim: int, flip_sin_to_cos: bool, freq_shift: float,
                       max_period: int = 10000) -> torch.Tensor:
    """D/models/embeddings.py:22-62 (scale=1)."""
    half = dim // 2
    exponent = -math.log(max_period) * torch.arange(0, half, dtype=torch.float32, device=t.device)
    exponent = exponent / (half - freq_shift)
    emb = t[:, None].float() * torch.exp(exponent)[None, :]
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
    if flip_sin_to_cos:
        emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)
    if dim % 2 == 1:
        emb = F.pad(emb, (0, 1, 0, 0))
    return emb


def resnet_block(sd, p: str, x: torch.Tensor, temb: torch.Tensor, groups: int, eps: float,
                 scale: float = 1.0) -> torch.Tensor:
    """D/models/resnet.py:551-601 (time_embedding_norm='default', no up/down, dropout 0)."""
    h = F.silu(F.group_norm(x, groups, sd[p + "norm1.weight"], sd[p + "norm1.bias"], eps))
    h = F.conv2d(h, sd[p + "conv1.weight"], sd[p + "conv1.bias"], padding=1)
    tp = F.linear(F.silu(temb), sd[p + "time_emb_proj.weight"], sd[p + "time_emb_proj.bias"])
    h = h + tp[:, :, None, None]
    h = F.silu(F.group_norm(h, groups, sd[p + "norm2.weight"], sd[p + "norm2.bias"], eps))
    h = F.conv2d(h, sd[p + "conv2.weight"], sd[p + "conv2.bias"], padding=1)
    if (p + "conv_shortcut.weight") in sd:
        x = F.conv2d(x, sd[p + "conv_shortcut.weight"], sd[p + "conv_shortcut.bias"])
    return (x + h) / scale


def attention_block(sd, p: str, x: torch.Tensor, groups: int, eps: float, head_dim: Optional[int],
                    rescale: float = 1.0) -> torch.Tensor:
    """D/models/attention.py:121-174 (non-xformers branch)."""
    b, c, hh, ww = x.shape
    heads = c // head_dim if head_dim is not None else 1
    h = F.group_norm(x, groups, sd[p + "group_norm.weight"], sd[p + "group_norm.bias"], eps)
    h = h.view(b, c, hh * ww).transpose(1, 2)
    q = F.linear(h, sd[p + "query.weight"], sd[p + "query.bias"])
    k = F.linear(h, sd[p + "key.weight"], sd[p + "key.bias"])
    v = F.linear(h, sd[p + "value.weight"], sd[p + "value.bias"])
    scale = 1 / math.sqrt(c / heads)

    def split(z):  # attention.py:77-82
        return z.reshape(b, hh * ww, heads, c // heads).permute(0, 2, 1, 3).reshape(b * heads, hh * ww, c // heads)

    q, k, v = split(q), split(k), split(v)
    scores = torch.bmm(q, k.transpose(-1, -2)) * scale
    probs = torch.softmax(scores.float(), dim=-1).type(scores.dtype)
    o = torch.bmm(probs, v)
    o = o.reshape(b, heads, hh * ww, c // heads).permute(0, 2, 1, 3).reshape(b, hh * ww, c)
    o = F.linear(o, sd[p + "proj_attn.weight"], sd[p + "proj_attn.bias"])
    o = o.transpose(-1, -2).reshape(b, c, hh, ww)
    return (o + x) / rescale


def downsample(sd, p: str, x: torch.Tensor, padding: int) -> torch.Tensor:
    """D/models/resnet.py:199-208."""
    if padding == 0:
        x = F.pad(x, (0, 1, 0, 1), mode="constant", value=0)
    return F.conv2d(x, sd[p + "conv.weight"], sd[p + "conv.bias"], stride=2, padding=padding)


def upsample(sd, p: str, x: torch.Tensor) -> torch.Tensor:
    """D/models/resnet.py:126-161."""
    x = F.interpolate(x, scale_factor=2.0, mode="nearest")
    return F.conv2d(x, sd[p + "conv.weight"], sd[p + "conv.bias"], padding=1)


def unet_forward(sd: Dict[str, torch.Tensor], cfg: dict, sample: torch.Tensor, timestep) -> torch.Tensor:
    """D/models/unet_2d.py:229-326 for the positional-embedding, unconditional UNet."""
    topo = unet_topology(cfg)
    g, eps, hd = cfg["norm_num_groups"], cfg["norm_eps"], cfg["attention_head_dim"]
    if cfg["center_input_sample"]:
        sample = 2 * sample - 1.0
    t = timestep
    if not torch.is_tensor(t):
        t = torch.tensor([t], dtype=torch.long, device=sample.device)
    elif t.dim() == 0:
        t = t[None]
    t = t * torch.ones(sample.shape[0], dtype=t.dtype, device=t.device)
    temb = timestep_embedding(t, cfg["block_out_channels"][0], cfg["flip_sin_to_cos"], cfg["freq_shift"])
    temb = temb.to(sample.dtype)
    emb = F.linear(temb, sd["time_embedding.linear_1.weight"], sd["time_embedding.linear_1.bias"])
    emb = F.linear(F.silu(emb), sd["time_embedding.linear_2.weight"], sd["time_embedding.linear_2.bias"])

    x = F.conv2d(sample, sd["conv_in.weight"], sd["conv_in.bias"], padding=1)
    skips = [x]
    for i, b in enumerate(topo["down"]):
        for j in range(len(b["resnets"])):
            x = resnet_block(sd, f"down_blocks.{i}.resnets.{j}.", x, emb, g, eps)
            if b["attn"]:
                x = attention_block(sd, f"down_blocks.{i}.attentions.{j}.", x, g, eps, hd)
            skips.append(x)
        if b["down"]:
            x = downsample(sd, f"down_blocks.{i}.downsamplers.0.", x, cfg["downsample_padding"])
            skips.append(x)
    ms = cfg["mid_block_scale_factor"]
    x = resnet_block(sd, "mid_block.resnets.0.", x, emb, g, eps, ms)
    if cfg.get("add_attention", True):
        x = attention_block(sd, "mid_block.attentions.0.", x, g, eps, hd, ms)
    x = resnet_block(sd, "mid_block.resnets.1.", x, emb, g, eps, ms)
    for i, b in enumerate(topo["up"]):
        for j in range(len(b["resnets"])):
            x = torch.cat([x, skips.pop()], dim=1)
            x = resnet_block(sd, f"up_blocks.{i}.resnets.{j}.", x, emb, g, eps)
            if b["attn"]:
                x = attention_block(sd, f"up_blocks.{i}.attentions.{j}.", x, g, eps, hd)
        if b["up"]:
            x = upsample(sd, f"up_blocks.{i}.upsamplers.0.", x)
    x = F.silu(F.group_norm(x, g, sd["conv_norm_out.weight"], sd["conv_norm_out.bias"], eps))
    return F.conv2d(x, sd["conv_out.weight"], sd["conv_out.bias"], padding=1)


# --------------------------------------------------------------------------------------------
# Schedulers
# --------------------------------------------------------------------------------------------
def beta_tables(num_train_timesteps=1000, beta_start=1e-4, beta_end=0.02):
    """D/schedulers/scheduling_ddpm.py:143,159-160 (linear schedule), fp32."""
    betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
    alphas = 1.0 - betas
    return betas, alphas, torch.cumprod(alphas, dim=0)


def timesteps_for(num_inference_steps: int, num_train_timesteps: int = 1000) -> np.ndarray:
    """scheduling_ddpm.py:238-246 / scheduling_ddim.py:252-258 (steps_offset 0)."""
    ratio = num_train_timesteps // num_inference_steps
    return (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64)


def add_noise(acp: torch.Tensor, x0: torch.Tensor, noise: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
    """scheduling_ddpm.py:422-443."""
    a = (acp[t] ** 0.5).flatten()
    s = ((1 - acp[t]) ** 0.5).flatten()
    while a.dim() < x0.dim():
        a, s = a.unsqueeze(-1), s.unsqueeze(-1)
    return a * x0 + s * noise


def ddpm_step(acp: torch.Tensor, eps_hat: torch.Tensor, t: int, x: torch.Tensor, z: Optional[torch.Tensor],
              num_inference_steps: int = 1000, num_train_timesteps: int = 1000, variance_type="fixed_small",
              clip_sample=True, clip_range=1.0, clip_defense=False, clip_defense_range=1.0) -> torch.Tensor:
    """scheduling_ddpm.py:250-288,324-420 (epsilon prediction, fixed_small / fixed_large)."""
    prev_t = t - num_train_timesteps // num_inference_steps
    one = torch.tensor(1.0)
    ap_t = acp[t]
    ap_prev = acp[prev_t] if prev_t >= 0 else one
    bp_t = 1 - ap_t
    bp_prev = 1 - ap_prev
    cur_a = ap_t / ap_prev
    cur_b = 1 - cur_a
    x0 = (x - bp_t ** 0.5 * eps_hat) / ap_t ** 0.5
    if clip_sample:
        x0 = x0.clamp(-clip_range, clip_range)
    c0 = (ap_prev ** 0.5 * cur_b) / bp_t
    ct = cur_a ** 0.5 * bp_prev / bp_t
    mu = c0 * x0 + ct * x
    var = 0
    if t > 0:
        v = torch.clamp((1 - ap_prev) / (1 - ap_t) * cur_b, min=1e-20)
        if variance_type == "fixed_large":
            v = cur_b
        elif variance_type != "fixed_small":
            raise NotImplementedError(variance_type)
        var = (v ** 0.5) * z
    out = mu + var
    if clip_defense:
        out = out.clamp(-clip_defense_range, clip_defense_range)
    return out


def ddim_step(acp: torch.Tensor, eps_hat: torch.Tensor, t: int, x: torch.Tensor, num_inference_steps: int,
              eta: float = 0.0, z: Optional[torch.Tensor] = None, num_train_timesteps: int = 1000,
              clip_sample=True, clip_range=1.0, set_alpha_to_one=True, use_clipped_model_output=False):
    """scheduling_ddim.py:261-381 (epsilon prediction)."""
    prev_t = t - num_train_timesteps // num_inference_steps
    final = torch.tensor(1.0) if set_alpha_to_one else acp[0]
    ap_t = acp[t]
    ap_prev = acp[prev_t] if prev_t >= 0 else final
    bp_t = 1 - ap_t
    x0 = (x - bp_t ** 0.5 * eps_hat) / ap_t ** 0.5
    pred_eps = eps_hat
    if clip_sample:
        x0 = x0.clamp(-clip_range, clip_range)
    variance = ((1 - ap_prev) / bp_t) * (1 - ap_t / ap_prev)
    std = eta * variance ** 0.5
    if use_clipped_model_output:
        pred_eps = (x - ap_t ** 0.5 * x0) / bp_t ** 0.5
    direction = (1 - ap_prev - std ** 2) ** 0.5 * pred_eps
    prev = ap_prev ** 0.5 * x0 + direction
    if eta > 0:
        prev = prev + std * z
    return prev


# --------------------------------------------------------------------------------------------
# BadDiffusion attack arithmetic
# --------------------------------------------------------------------------------------------
def poison_blend(image: torch.Tensor, is_poison: torch.Tensor, trigger: torch.Tensor, target: torch.Tensor,
                 vmin: float = -1.0) -> Tuple[torch.Tensor, torch.Tensor]:
    """dataset.py:275-276 (get_mask), :288-315 (clean / backdoor transforms).  Returns (R, x0) =
    (`pixel_values`, `target`) as consumed at baddiffusion.py:593-594."""
    mask = torch.where(trigger > vmin, 0, 1).to(image.dtype)
    R_p = mask[None] * image + (1 - mask[None]) * trigger[None]
    p = is_poison.reshape(-1, 1, 1, 1).bool()
    R = torch.where(p, R_p, torch.zeros_like(image))
    x0 = torch.where(p, target[None].expand_as(image), image)
    return R, x0


def q_sample(alphas: torch.Tensor, acp: torch.Tensor, x0: torch.Tensor, R: torch.Tensor, t: torch.Tensor,
             noise: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """loss.py:257-285."""
    n = len(x0)
    a = (acp[t] ** 0.5).reshape(n, 1, 1, 1)
    s = (1 - acp[t]) ** 0.5
    rho = ((1 - alphas[t] ** 0.5) * s / (1 - alphas[t])).reshape(n, 1, 1, 1)
    noisy = add_noise(acp, x0, noise, t)
    return noisy + (1 - a) * R, rho * R + noise


def p_losses(sd, cfg, alphas, acp, x0, R, t, noise) -> torch.Tensor:
    """loss.py:287-307, loss_type='l2'."""
    x_noisy, target = q_sample(alphas, acp, x0, R, t, noise)
    eps_hat = unet_forward(sd, cfg, x_noisy.contiguous(), t.contiguous())
    return F.mse_loss(target, eps_hat)


# Backdoor trigger / target tensors -----------------------------------------------------------
def normalize(x, vmin_in=0.0, vmax_in=1.0, vmin_out=-1.0, vmax_out=1.0, eps=1e-5):
    """util.py:83-111 (quirk Q7: divides by max-min+1e-5)."""
    return ((x - vmin_in) / (vmax_in - vmin_in + eps)) * (vmax_out - vmin_out) + vmin_out


def u8_batch_to_image(u8_nhwc, flips) -> torch.Tensor:
    """dataset.py:120-136 on an already-sized RGB batch: ToTensor (uint8 HWC -> fp32 CHW / 255), util.normalize to
    [-1, 1 - 2e-5] (quirk Q7), RandomHorizontalFlip with the per-sample coins `flips` (drawn by draw_flips)."""
    x = torch.as_tensor(u8_nhwc).permute(0, 3, 1, 2).to(torch.float32).div(255)
    x = normalize(x)
    fl = torch.as_tensor(flips, dtype=torch.bool).view(-1, 1, 1, 1)
    return torch.where(fl, x.flip(-1), x)


def draw_flips(n: int, generator=None, p: float = 0.5) -> torch.Tensor:
    """torchvision RandomHorizontalFlip.forward: one `torch.rand(1) < p` per image, in batch order."""
    return torch.tensor([bool(torch.rand(1, generator=generator) < p) for _ in range(n)])


def _load_bitmap(path: str, size, channel: int = 3) -> torch.Tensor:
    """dataset.py:420-434: convert -> Resize(bilinear, antialias) -> ToTensor -> normalize."""
    from PIL import Image
    from torchvision import transforms

    img = Image.open(path)
    img = img.convert("RGB") if channel == 3 else img.convert("L")
    img = transforms.Resize(size)(img)
    return normalize(transforms.ToTensor()(img))


def bg2grey(x: torch.Tensor, vmin=-1.0, vmax=1.0) -> torch.Tensor:
    """dataset.py:447-450."""
    thres = (vmax - vmin) * 0.3 + vmin
    x = x.clone()
    x[x <= thres] = thres
    return x


def box_trigger(size: int, box: int, channel: int = 3, vmin=-1.0, vmax=1.0, grey=True) -> torch.Tensor:
    """dataset.py:504-524, 560-574 (BOX_k: grey k x k box, 2 px from the bottom-right corner)."""
    trig = torch.full((channel, size, size), float(vmin))
    val = (vmin + vmax) / 2 if grey else vmax
    trig[:, -(box + 2):-2, -(box + 2):-2] = val
    return trig


def image_trigger(path: str, size: int, trigger_sz: int, x: Optional[int] = None, y: Optional[int] = None,
                  channel: int = 3, vmin=-1.0) -> torch.Tensor:
    """dataset.py:472-497 (GLASSES: centred; STOP_SIGN_k: x=y=-2)."""
    l_pad = t_pad = int((size - trigger_sz) / 2)
    r_pad = size - trigger_sz - l_pad
    b_pad = size - trigger_sz - t_pad
    residual = size - trigger_sz
    if x is not None:
        if x > 0:
            l_pad, r_pad = x, residual - x
        else:
            r_pad = -x
            l_pad = residual - r_pad
    if y is not None:
        if y > 0:
            t_pad, b_pad = y, residual - y
        else:
            b_pad = -y
            t_pad = residual - b_pad
    trig = _load_bitmap(path, trigger_sz, channel)
    trig = F.pad(trig, (l_pad, r_pad, t_pad, b_pad), value=vmin)
    trig[trig >= 0.999] = vmin
    return trig


def get_trigger(kind: str, size: int, static_dir: Optional[str] = None, channel: int = 3) -> torch.Tensor:
    """dataset.py:526-597 (the triggers reachable offline)."""
    import os

    if kind.startswith("BOX_"):
        return box_trigger(size, int(kind.split("_")[1]), channel, grey=True)
    boxes = {"SM_BOX": 14, "XSM_BOX": 11, "XXSM_BOX": 8, "XXXSM_BOX": 4, "BIG_BOX": 18}
    if kind in boxes:
        return box_trigger(size, boxes[kind], channel, grey=False)
    if kind == "GLASSES":
        return image_trigger(os.path.join(static_dir, "glasses.png"), size, int(size * 0.625), channel=channel)
    if kind.startswith("STOP_SIGN_"):
        return image_trigger(os.path.join(static_dir, "stop_sign_wo_bg.png"), size, int(kind.split("_")[2]),
                             x=-2, y=-2, channel=channel)
    if kind == "NONE":
        return torch.full((channel, size, size), -1.0)
    raise ValueError(f"Trigger type {kind} isn't found")


def get_target(kind: str, trigger: torch.Tensor, static_dir: Optional[str] = None, dx=-5, dy=-3) -> torch.Tensor:
    """dataset.py:627-655."""
    import os

    channel, size = trigger.shape[0], list(trigger.shape[-2:])
    if kind == "TRIGGER":
        return bg2grey(trigger)
    if kind == "SHIFT":
        return bg2grey(torch.roll(trigger, shifts=(0, dy, dx), dims=(0, 1, 2)))
    if kind == "CORNER":
        t = torch.full((channel, *size), -1.0)
        t[:, :10, :10] = 0.0
        return bg2grey(t)
    if kind == "HAT":
        return bg2grey(_load_bitmap(os.path.join(static_dir, "fedora-hat.png"), size, channel))
    if kind == "CAT":
        return bg2grey(_load_bitmap(os.path.join(static_dir, "cat_wo_bg.png"), size, channel))
    raise NotImplementedError(f"Target type {kind} isn't found")


# --------------------------------------------------------------------------------------------
# Sampling loops
# --------------------------------------------------------------------------------------------
def _randn(shape, generator):
    """D/utils/torch_utils.py:29-70 with a CPU generator (quirk Q11)."""
    return torch.randn(shape, generator=generator, dtype=torch.float32)


def postprocess(x: torch.Tensor) -> np.ndarray:
    """pipeline_ddpm.py:115-116."""
    return (x / 2 + 0.5).clamp(0, 1).permute(0, 2, 3, 1).numpy()


@torch.no_grad()
def ddpm_pipeline(sd, cfg, sched_cfg: dict, batch_size: int, generator=None, num_inference_steps=1000,
                  init: Optional[torch.Tensor] = None, start_from: int = 0, return_raw=False):
    """pipeline_ddpm.py:46-125."""
    _, _, acp = beta_tables(sched_cfg.get("num_train_timesteps", 1000), sched_cfg.get("beta_start", 1e-4),
                            sched_cfg.get("beta_end", 0.02))
    shape = (batch_size, cfg["in_channels"], cfg["sample_size"], cfg["sample_size"])
    image = _randn(shape, generator) if init is None else init.detach().clone()
    for t in timesteps_for(num_inference_steps)[start_from:]:
        t = int(t)
        eps_hat = unet_forward(sd, cfg, image, t)
        z = _randn(eps_hat.shape, generator) if t > 0 else None
        image = ddpm_step(acp, eps_hat, t, image, z, num_inference_steps,
                          variance_type=sched_cfg.get("variance_type", "fixed_small"),
                          clip_sample=sched_cfg.get("clip_sample", True),
                          clip_defense=sched_cfg.get("clip_defense", False),
                          clip_defense_range=sched_cfg.get("clip_defense_range", 1.0))
    return image if return_raw else postprocess(image)


@torch.no_grad()
def ddim_pipeline(sd, cfg, sched_cfg: dict, batch_size: int, generator=None, num_inference_steps=50, eta=0.0,
                  init: Optional[torch.Tensor] = None, return_raw=False):
    """pipeline_ddim.py:50-142."""
    _, _, acp = beta_tables(sched_cfg.get("num_train_timesteps", 1000), sched_cfg.get("beta_start", 1e-4),
                            sched_cfg.get("beta_end", 0.02))
    shape = (batch_size, cfg["in_channels"], cfg["sample_size"], cfg["sample_size"])
    image = _randn(shape, generator) if init is None else init.detach().clone()
    for t in timesteps_for(num_inference_steps):
        t = int(t)
        eps_hat = unet_forward(sd, cfg, image, t)
        z = _randn(eps_hat.shape, generator) if eta > 0 else None
        image = ddim_step(acp, eps_hat, t, image, num_inference_steps, eta=eta, z=z,
                          clip_sample=sched_cfg.get("clip_sample", True))
    return image if return_raw else postprocess(image)


def batch_sampling(sample_n: int, run_pipeline, init: Optional[torch.Tensor] = None, max_batch_n: int = 256,
                   rng=None) -> np.ndarray:
    """model.py:469-489.  `run_pipeline(batch_size, generator, init)` -> NHWC ndarray."""
    if init is None:
        if sample_n > max_batch_n:
            replica, residual = sample_n // max_batch_n, sample_n % max_batch_n
            sizes = [max_batch_n] * replica + ([residual] if residual > 0 else [])
        else:
            sizes = [sample_n]
        chunks = [None] * len(sizes)
    else:
        chunks = torch.split(init, max_batch_n)
        sizes = [len(c) for c in chunks]
    return np.concatenate([run_pipeline(bs, rng, chunks[i]) for i, bs in enumerate(sizes)])


def cosine_lr_lambda(step: int, warmup: int, total: int, num_cycles: float = 0.5) -> float:
    """D/optimization.py:134-138."""
    if step < warmup:
        return float(step) / float(max(1, warmup))
    progress = float(step - warmup) / float(max(1, total - warmup))
    return max(0.0, 0.5 * (1.0 + math.cos(math.pi * float(num_cycles) * 2.0 * progress)))


# Measurement tail (baddiffusion.py:533-546) -----------------------------------------------------
def ssim_torchmetrics(preds: torch.Tensor, target: torch.Tensor, data_range: float = 1.0, kernel_size: int = 11,
                      sigma: float = 1.5, k1: float = 0.01, k2: float = 0.03) -> torch.Tensor:
    """torchmetrics.StructuralSimilarityIndexMeasure(data_range=1.0)(preds, target) with its defaults, restated from the
    published algorithm of torchmetrics.functional.image.ssim (`_ssim_update`, gaussian_kernel=True, reduction
    'elementwise_mean'): reflect-pad by 5, depthwise 11x11 Gaussian filter of (x, y, xx, yy, xy), SSIM map, crop the
    padded border again, mean per image, mean over images.  torchmetrics is not installed here and not vendored by the
    reference (requirements.txt): PARITY UNPINNED for this function beyond this restatement."""
    B, C, H, W = preds.shape
    dist = torch.arange((1 - kernel_size) / 2, (1 + kernel_size) / 2, 1, dtype=preds.dtype)
    g = torch.exp(-torch.pow(dist / sigma, 2) / 2)
    g = (g / g.sum()).unsqueeze(0)
    kernel = torch.matmul(g.t(), g).expand(C, 1, kernel_size, kernel_size)
    pad = (kernel_size - 1) // 2
    c1, c2 = (k1 * data_range) ** 2, (k2 * data_range) ** 2
    p = F.pad(preds, (pad, pad, pad, pad), mode="reflect")
    t = F.pad(target, (pad, pad, pad, pad), mode="reflect")
    out = F.conv2d(torch.cat((p, t, p * p, t * t, p * t)), kernel, groups=C)
    mu_p, mu_t, pp, tt, pt = out.split(B)
    mu_pp, mu_tt, mu_pt = mu_p.pow(2), mu_t.pow(2), mu_p * mu_t
    s_p, s_t, s_pt = pp - mu_pp, tt - mu_tt, pt - mu_pt
    upper, lower = 2 * s_pt + c2, s_p + s_t + c2
    ssim_full = ((2 * mu_pt + c1) * upper) / ((mu_pp + mu_tt + c1) * lower)
    ssim = ssim_full[..., pad:-pad, pad:-pad]
    return ssim.reshape(B, -1).mean(-1).mean()


def backdoor_metrics(samples_u8_nhwc, target_chw: torch.Tensor):
    """baddiffusion.py:539-546 on the uint8 samples the PNG files would hold: ToTensor (u8/255, CHW), the target as
    (y/2+0.5).clamp(0,1) repeated, nn.MSELoss and SSIM.  Returns (mse, ssim) as floats."""
    x = torch.as_tensor(samples_u8_nhwc).permute(0, 3, 1, 2).to(torch.float32).div(255)
    y = (target_chw / 2 + 0.5).clamp(0, 1).unsqueeze(0).expand_as(x).contiguous()
    return float(F.mse_loss(x, y)), float(ssim_torchmetrics(x, y))
