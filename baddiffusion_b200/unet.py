"""Drop-in `UNet2DModel` (the unconditional DDPM UNet of the vendored diffusers, D/models/unet_2d.py:82-326)
whose forward / backward run on the hand-written sm_100a kernels of libb200bd.

What is kept from the reference: constructor arguments (= `config.json` keys),
`forward(sample, timestep, class_labels=None, return_dict=True)` returning an object with `.sample` (or a
1-tuple), `.config`, `.in_channels` / `.sample_size`, `.device` / `.dtype`, `parameters()` and a `state_dict()`
with the reference's exact key names and OIHW shapes, `save_pretrained` / `from_pretrained` with
`unet/config.json` + `unet/diffusion_pytorch_model.bin`.

What is different underneath: every parameter is a *view* of ONE flat fp32 buffer laid out for the kernels
(conv weights stored [kh][kw][O][I] and exposed as permuted OIHW views; q/k/v and all time_emb_proj matrices
adjacent so they run as single GEMMs); gradients are views of one flat fp32 buffer (one NCCL all-reduce, one
fused Adam launch).  There is no PyTorch compute fallback.
"""
from __future__ import annotations

import math
import os
from collections import OrderedDict
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple, Union

import torch
import torch.nn as nn

from . import config_utils as CU

WEIGHTS_NAME = "diffusion_pytorch_model.bin"  # D/utils/constants.py:22
CONFIG_NAME = "config.json"

_SUPPORTED_DOWN = ("DownBlock2D", "AttnDownBlock2D")
_SUPPORTED_UP = ("UpBlock2D", "AttnUpBlock2D")


@dataclass
class UNet2DOutput:  # D/models/unet_2d.py:28-36
    sample: torch.Tensor


def topology(cfg) -> dict:
    """Block list implied by the UNet2DModel constructor (D/models/unet_2d.py:147-210)."""
    boc = list(cfg["block_out_channels"])
    L = cfg["layers_per_block"]
    down, out_c = [], boc[0]
    for i, typ in enumerate(cfg["down_block_types"]):
        in_c, out_c = out_c, boc[i]
        down.append(dict(type=typ, resnets=[(in_c if j == 0 else out_c, out_c) for j in range(L)],
                         attn=typ == "AttnDownBlock2D", down=i != len(boc) - 1, channels=out_c))
    rev = list(reversed(boc))
    up, out_c = [], rev[0]
    for i, typ in enumerate(cfg["up_block_types"]):
        prev_out, out_c = out_c, rev[i]
        in_c = rev[min(i + 1, len(boc) - 1)]
        res = []
        for j in range(L + 1):  # D/models/unet_2d_blocks.py:1893-1896
            skip_c = in_c if j == L else out_c
            res_in = prev_out if j == 0 else out_c
            res.append((res_in, skip_c, out_c))
        up.append(dict(type=typ, resnets=res, attn=typ == "AttnUpBlock2D", up=i != len(boc) - 1, channels=out_c))
    return dict(temb=boc[0] * 4, down=down, up=up, mid=boc[-1])


def _resnet_entries(p, cin, cout, temb):
    e = [(p + "norm1.weight", (cin,)), (p + "norm1.bias", (cin,)), (p + "conv1.weight", (cout, cin, 3, 3)),
         (p + "conv1.bias", (cout,)), (p + "time_emb_proj.weight", (cout, temb)), (p + "time_emb_proj.bias", (cout,)),
         (p + "norm2.weight", (cout,)), (p + "norm2.bias", (cout,)), (p + "conv2.weight", (cout, cout, 3, 3)),
         (p + "conv2.bias", (cout,))]
    if cin != cout:  # D/models/resnet.py:541-549
        e += [(p + "conv_shortcut.weight", (cout, cin, 1, 1)), (p + "conv_shortcut.bias", (cout,))]
    return e


def _attn_entries(p, c):
    e = [(p + "group_norm.weight", (c,)), (p + "group_norm.bias", (c,))]
    for n in ("query", "key", "value", "proj_attn"):
        e += [(p + n + ".weight", (c, c)), (p + n + ".bias", (c,))]
    return e


def param_entries(cfg) -> "OrderedDict[str, tuple]":
    """state_dict keys in the reference's registration order with torch (OIHW) shapes."""
    topo = topology(cfg)
    boc, temb = list(cfg["block_out_channels"]), topo["temb"]
    e = [("conv_in.weight", (boc[0], cfg["in_channels"], 3, 3)), ("conv_in.bias", (boc[0],)),
         ("time_embedding.linear_1.weight", (temb, boc[0])), ("time_embedding.linear_1.bias", (temb,)),
         ("time_embedding.linear_2.weight", (temb, temb)), ("time_embedding.linear_2.bias", (temb,))]
    for i, b in enumerate(topo["down"]):
        if b["attn"]:
            for j in range(len(b["resnets"])):
                e += _attn_entries(f"down_blocks.{i}.attentions.{j}.", b["channels"])
        for j, (ci, co) in enumerate(b["resnets"]):
            e += _resnet_entries(f"down_blocks.{i}.resnets.{j}.", ci, co, temb)
        if b["down"]:
            e += [(f"down_blocks.{i}.downsamplers.0.conv.weight", (b["channels"],) * 2 + (3, 3)),
                  (f"down_blocks.{i}.downsamplers.0.conv.bias", (b["channels"],))]
    for i, b in enumerate(topo["up"]):
        if b["attn"]:
            for j in range(len(b["resnets"])):
                e += _attn_entries(f"up_blocks.{i}.attentions.{j}.", b["channels"])
        for j, (ri, sk, co) in enumerate(b["resnets"]):
            e += _resnet_entries(f"up_blocks.{i}.resnets.{j}.", ri + sk, co, temb)
        if b["up"]:
            e += [(f"up_blocks.{i}.upsamplers.0.conv.weight", (b["channels"],) * 2 + (3, 3)),
                  (f"up_blocks.{i}.upsamplers.0.conv.bias", (b["channels"],))]
    if cfg.get("add_attention", True):
        e += _attn_entries("mid_block.attentions.0.", topo["mid"])
    e += _resnet_entries("mid_block.resnets.0.", topo["mid"], topo["mid"], temb)
    e += _resnet_entries("mid_block.resnets.1.", topo["mid"], topo["mid"], temb)
    e += [("conv_norm_out.weight", (boc[0],)), ("conv_norm_out.bias", (boc[0],)),
          ("conv_out.weight", (cfg["out_channels"], boc[0], 3, 3)), ("conv_out.bias", (cfg["out_channels"],))]
    return OrderedDict(e)


def resnet_prefixes(cfg) -> List[str]:
    """ResnetBlock2D prefixes in EXECUTION order (the order of the fused time_emb_proj GEMM's output columns)."""
    topo = topology(cfg)
    out = []
    for i, b in enumerate(topo["down"]):
        out += [f"down_blocks.{i}.resnets.{j}." for j in range(len(b["resnets"]))]
    out += ["mid_block.resnets.0.", "mid_block.resnets.1."]
    for i, b in enumerate(topo["up"]):
        out += [f"up_blocks.{i}.resnets.{j}." for j in range(len(b["resnets"]))]
    return out


class FlatLayout:
    """Offsets of every parameter inside the flat buffer.  Region A (tensor-core GEMM operands, shadowed in fp16):
    3x3/1x1 conv weights in packed [kh][kw][O][I] order, fused [q;k;v] and proj matrices, all time_emb_proj
    matrices back to back.  Region B (fp32 only): biases, norm affine params, conv_in / conv_out, timestep MLP."""

    ALIGN = 64  # elements: 256 B in fp32, 128 B in fp16 (TMA base alignment)

    def __init__(self, cfg):
        self.entries = param_entries(cfg)
        self.offset: Dict[str, int] = {}
        order_a: List[str] = []
        order_b: List[str] = []
        names = list(self.entries)
        tproj_w = [p + "time_emb_proj.weight" for p in resnet_prefixes(cfg)]
        tproj_b = [p + "time_emb_proj.bias" for p in resnet_prefixes(cfg)]
        conv1_b = [p + "conv1.bias" for p in resnet_prefixes(cfg)]  # same column order: d(conv1.bias) == d(time_emb_proj.bias)
        special = set(tproj_w) | set(tproj_b) | set(conv1_b)
        for k in names:
            if k in special:
                continue
            shp = self.entries[k]
            is_gemm_w = k.endswith(".weight") and len(shp) in (2, 4) and not k.startswith(("conv_in.", "conv_out.", "time_embedding."))
            (order_a if is_gemm_w else order_b).append(k)
        off = 0

        def place(k, align=True):
            nonlocal off
            if align:
                off = (off + self.ALIGN - 1) // self.ALIGN * self.ALIGN
            self.offset[k] = off
            off += int(math.prod(self.entries[k]))

        for k in order_a:
            # query/key/value (and their biases below) stay adjacent without padding: C*C is a multiple of ALIGN
            place(k)
        off = (off + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        self.tproj_w_offset = off
        for k in tproj_w:
            place(k, align=False)
        self.tproj_rows = sum(self.entries[k][0] for k in tproj_w)
        off = (off + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        self.n_gemm = off  # end of region A
        self.tproj_b_offset = off
        for k in tproj_b:
            place(k, align=False)
        off = (off + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        self.conv1_b_offset = off
        for k in conv1_b:
            place(k, align=False)
        for k in order_b:
            # q/k/v biases adjacent (3C contiguous): do not pad between them
            tail = k.split(".")[-2]
            place(k, align=tail not in ("key", "value"))
        self.total = (off + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        # column offset of each resnet inside the fused time_emb_proj output
        self.tproj_col: Dict[str, int] = {}
        c = 0
        for p in resnet_prefixes(cfg):
            self.tproj_col[p] = c
            c += self.entries[p + "time_emb_proj.weight"][0]

    def view(self, flat: torch.Tensor, key: str) -> torch.Tensor:
        """torch-shaped (OIHW for convs) view of `key` inside `flat`."""
        shp = self.entries[key]
        n = int(math.prod(shp))
        seg = flat[self.offset[key]: self.offset[key] + n]
        if len(shp) == 4:
            O, I, kh, kw = shp
            return seg.view(kh, kw, O, I).permute(2, 3, 0, 1)
        return seg.view(*shp)

    def packed(self, flat: torch.Tensor, key: str) -> torch.Tensor:
        """kernel-side view: conv weights as [taps][O][I], everything else as stored."""
        shp = self.entries[key]
        n = int(math.prod(shp))
        seg = flat[self.offset[key]: self.offset[key] + n]
        if len(shp) == 4:
            O, I, kh, kw = shp
            return seg.view(kh * kw, O, I)
        return seg.view(*shp)


class UNet2DModel(nn.Module):
    config_name = CONFIG_NAME

    def __init__(self, sample_size: Optional[Union[int, Tuple[int, int]]] = None, in_channels: int = 3,
                 out_channels: int = 3, center_input_sample: bool = False, time_embedding_type: str = "positional",
                 freq_shift: int = 0, flip_sin_to_cos: bool = True,
                 down_block_types: Tuple[str] = ("DownBlock2D", "AttnDownBlock2D", "AttnDownBlock2D", "AttnDownBlock2D"),
                 up_block_types: Tuple[str] = ("AttnUpBlock2D", "AttnUpBlock2D", "AttnUpBlock2D", "UpBlock2D"),
                 block_out_channels: Tuple[int] = (224, 448, 672, 896), layers_per_block: int = 2,
                 mid_block_scale_factor: float = 1, downsample_padding: int = 1, act_fn: str = "silu",
                 attention_head_dim: Optional[int] = 8, norm_num_groups: int = 32, norm_eps: float = 1e-5,
                 resnet_time_scale_shift: str = "default", add_attention: bool = True,
                 class_embed_type: Optional[str] = None, num_class_embeds: Optional[int] = None):
        super().__init__()
        cfg = CU.capture_init_args(self, UNet2DModel.__init__, (), {k: v for k, v in locals().items()
                                                                    if k not in ("self", "__class__")})
        self.sample_size = sample_size
        # same argument validation as the reference (unet_2d.py:112-120)
        if len(down_block_types) != len(up_block_types):
            raise ValueError(f"Must provide the same number of `down_block_types` as `up_block_types`. "
                             f"`down_block_types`: {down_block_types}. `up_block_types`: {up_block_types}.")
        if len(block_out_channels) != len(down_block_types):
            raise ValueError(f"Must provide the same number of `block_out_channels` as `down_block_types`. "
                             f"`block_out_channels`: {block_out_channels}. `down_block_types`: {down_block_types}.")
        # the hot path is the unconditional DDPM UNet; everything else in the block zoo is out of scope
        for t in down_block_types:
            if t not in _SUPPORTED_DOWN:
                raise NotImplementedError(f"down block {t} is outside the BadDiffusion hot path")
        for t in up_block_types:
            if t not in _SUPPORTED_UP:
                raise NotImplementedError(f"up block {t} is outside the BadDiffusion hot path")
        if time_embedding_type != "positional" or resnet_time_scale_shift != "default" or class_embed_type is not None \
                or num_class_embeds is not None or act_fn not in ("silu", "swish"):
            raise NotImplementedError("only the positional-embedding, unconditional, SiLU UNet2DModel is implemented")
        if isinstance(sample_size, (tuple, list)) and sample_size[0] != sample_size[1]:
            raise NotImplementedError("square samples only")
        for c in block_out_channels:
            if c % norm_num_groups or c % 8:
                raise ValueError("block_out_channels must be multiples of norm_num_groups and of 8")

        self.layout = FlatLayout(cfg)
        flat = torch.zeros(self.layout.total, dtype=torch.float32)
        self._flat = flat
        self._flat_grad: Optional[torch.Tensor] = None
        self._flat16: Optional[torch.Tensor] = None
        self._flat16_version = -1
        self._param_version = 0
        self._engines = {}
        self._names = list(self.layout.entries)
        self._init_weights()
        for key in self._names:
            self._register(key, nn.Parameter(self.layout.view(flat, key)))

    # ------------------------------------------------------------------ parameters as views of one buffer
    def _register(self, key: str, p: nn.Parameter):
        mod = self
        parts = key.split(".")
        for name in parts[:-1]:
            if name not in mod._modules:
                mod.add_module(name, nn.Module())
            mod = mod._modules[name]
        mod.register_parameter(parts[-1], p)

    def _param(self, key: str) -> nn.Parameter:
        mod = self
        parts = key.split(".")
        for name in parts[:-1]:
            mod = mod._modules[name]
        return mod._parameters[parts[-1]]

    def _init_weights(self):
        """torch default initialisers (kaiming-uniform a=sqrt(5) for conv / linear, ones / zeros for GroupNorm)."""
        for key, shp in self.layout.entries.items():
            v = self.layout.view(self._flat, key)
            leaf = key.split(".")[-2]
            if "norm" in leaf:
                v.fill_(1.0 if key.endswith("weight") else 0.0)
                continue
            wkey = key[: key.rfind(".")] + ".weight"
            wshape = self.layout.entries[wkey]
            fan_in = int(math.prod(wshape[1:]))
            bound = 1.0 / math.sqrt(fan_in)
            v.copy_(torch.empty(shp).uniform_(-bound, bound))

    def _apply(self, fn, recurse=True):
        """Keep the one-flat-buffer invariant under .to() / .cuda() / .float(): move the buffer, re-point the views."""
        new_flat = fn(self._flat)
        if not new_flat.is_floating_point() or new_flat.dtype != torch.float32:
            raise NotImplementedError("master parameters are fp32 (the kernels keep their own fp16 operand copy)")
        self._flat = new_flat.contiguous()
        for key in self._names:
            p = self._param(key)
            p.data = self.layout.view(self._flat, key)
            p.grad = None
        self._flat_grad = None
        self._flat16 = None
        self._engines = {}
        self.mark_params_changed()
        return self

    def mark_params_changed(self):
        self._param_version += 1

    @property
    def flat_params(self) -> torch.Tensor:
        return self._flat

    def flat_grads(self, attach: bool = True) -> torch.Tensor:
        """The flat fp32 gradient buffer; every parameter's .grad is a view of it."""
        if self._flat_grad is None or self._flat_grad.device != self._flat.device:
            self._flat_grad = torch.zeros_like(self._flat)
        if attach:
            for key in self._names:
                p = self._param(key)
                if p.grad is None or p.grad.data_ptr() != self.layout.view(self._flat_grad, key).data_ptr():
                    p.grad = self.layout.view(self._flat_grad, key)
        return self._flat_grad

    def flat_half(self) -> torch.Tensor:
        """fp16 shadow of the GEMM-operand region, refreshed when the parameters changed."""
        from . import ops

        if self._flat16 is None or self._flat16.device != self._flat.device:
            self._flat16 = torch.empty(self.layout.n_gemm, dtype=torch.float16, device=self._flat.device)
            self._flat16_version = -1
        ver = (self._param_version, tuple(self._param(k)._version for k in self._names[:4]), self._flat._version)
        if ver != self._flat16_version:
            ops.cast_f32_to_f16(self._flat[: self.layout.n_gemm], self._flat16)
            self._flat16_version = ver
        return self._flat16

    def load_state_dict(self, state_dict, strict: bool = True):
        r = super().load_state_dict(state_dict, strict=strict)
        self.mark_params_changed()
        return r

    # ------------------------------------------------------------------ reference attribute surface
    @property
    def in_channels(self):  # deprecated accessor used at baddiffusion.py:410,512
        return self.config.in_channels

    @property
    def device(self) -> torch.device:
        return self._flat.device

    @property
    def dtype(self) -> torch.dtype:
        return torch.float32

    # ------------------------------------------------------------------ forward
    def engine(self, batch: int, train: bool):
        from .engine import UNetEngine

        key = (batch, bool(train))
        if key not in self._engines:
            self._engines[key] = UNetEngine(self, batch, train)
        return self._engines[key]

    def forward(self, sample: torch.Tensor, timestep: Union[torch.Tensor, float, int],
                class_labels: Optional[torch.Tensor] = None, return_dict: bool = True):
        """D/models/unet_2d.py:229-326."""
        if not sample.is_cuda or not self._flat.is_cuda:
            raise RuntimeError("baddiffusion_b200.UNet2DModel runs on CUDA (sm_100a) only -- there is no CPU path; "
                               "move the model and inputs to the GPU")
        if self.config.center_input_sample:
            sample = 2 * sample - 1.0
        B = sample.shape[0]
        timesteps = timestep
        if not torch.is_tensor(timesteps):
            timesteps = torch.tensor([timesteps], dtype=torch.long, device=sample.device)
        elif timesteps.dim() == 0:
            timesteps = timesteps[None].to(sample.device)
        timesteps = (timesteps * torch.ones(B, dtype=timesteps.dtype, device=timesteps.device)).to(
            device=sample.device, dtype=torch.long)
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        x = sample.to(torch.float32).contiguous()
        if need_grad:
            out = _UNetFunction.apply(self, x, timesteps, self._grad_anchor())
        else:
            out = self.engine(B, False).forward(x, timesteps).clone()
        if not return_dict:
            return (out,)
        return UNet2DOutput(sample=out)

    def _grad_anchor(self):
        if getattr(self, "_anchor", None) is None or self._anchor.device != self._flat.device:
            self._anchor = torch.zeros(1, device=self._flat.device, requires_grad=True)
        return self._anchor

    # ------------------------------------------------------------------ checkpoint layout (Appendix D)
    def save_pretrained(self, save_directory: str, **kwargs):
        """D/models/modeling_utils.py:245-303: config.json + diffusion_pytorch_model.bin (fp32 state_dict, OIHW)."""
        os.makedirs(save_directory, exist_ok=True)
        CU.save_config(self.config, "UNet2DModel", save_directory, CONFIG_NAME)
        sd = OrderedDict((k, v.detach().cpu().clone(memory_format=torch.contiguous_format)) for k, v in self.state_dict().items())
        torch.save(sd, os.path.join(save_directory, WEIGHTS_NAME))

    @classmethod
    def from_config(cls, config, **kwargs):
        d = dict(config)
        d.update(kwargs)
        return cls(**CU.filter_init_kwargs(cls, d))

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path: str, subfolder: Optional[str] = None, **kwargs):
        d = pretrained_model_name_or_path if subfolder is None else os.path.join(pretrained_model_name_or_path, subfolder)
        if not os.path.isdir(d):
            raise EnvironmentError(f"{d} is not a local directory (hub downloads are out of scope: no network)")
        model = cls.from_config(CU.load_config(d, CONFIG_NAME))
        path = os.path.join(d, WEIGHTS_NAME)
        if not os.path.isfile(path):
            raise EnvironmentError(f"Error no file named {WEIGHTS_NAME} found in directory {d}.")
        model.load_state_dict(torch.load(path, map_location="cpu"), strict=True)
        model.eval()  # D/models/modeling_utils.py:626 (quirk Q13)
        return model


class _UNetFunction(torch.autograd.Function):
    """Autograd bridge for API compatibility (`loss.backward()` after `p_losses_diffuser`): forward and backward
    are the engine's kernel sequences; parameter gradients are accumulated straight into the flat gradient
    buffer that every `param.grad` views."""

    @staticmethod
    def forward(ctx, model: UNet2DModel, x, timesteps, anchor):
        eng = model.engine(x.shape[0], True)
        out = eng.forward(x, timesteps)
        ctx.model, ctx.eng = model, eng
        return out.clone()

    @staticmethod
    def backward(ctx, grad_out):
        model, eng = ctx.model, ctx.eng
        fresh = any(model._param(k).grad is None for k in model._names)
        g = model.flat_grads(attach=True)
        if fresh:
            g.zero_()
        eng.backward(grad_out.contiguous().float(), g, loss_scale=1.0)
        return None, None, None, None
