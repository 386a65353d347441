"""Backdoor trigger / target tensors and the poisoned-batch protocol of the reference's dataset.py.

`Backdoor` (dataset.py:376-655) builds the trigger `g` and target `y` tensors; the per-sample blend
(dataset.py:275-276,288-315) is NOT done here on the host: `PoisonedBatch` only carries the raw images and the
`is_poison` flags and the fused batch-prep kernel applies mask/blend/add_noise/target in one pass.

Dataset download / PIL decode (HF `datasets`, 8 DataLoader workers) is out of scope: the
`SyntheticDataset` below produces the synthetic tensors of SURVEY.md 8(d).
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Optional, Tuple, Union

import numpy as np
import torch
import torch.nn.functional as F

DEFAULT_VMIN = float(-1.0)
DEFAULT_VMAX = float(1.0)
_ASSETS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets", "backdoor_assets.npz")


def normalize(x, vmin_in=None, vmax_in=None, vmin_out=0, vmax_out=1, eps=1e-5):
    """util.py:83-111 (quirk Q7: the divisor is max-min+1e-5, so [0,1] maps to [-1, 1-2e-5])."""
    if vmax_out is None and vmin_out is None:
        return x
    min_x = x.min() if vmin_in is None else vmin_in
    max_x = x.max() if vmax_in is None else vmax_in
    if vmax_out is None:
        vmax_out = max_x
    if vmin_out is None:
        vmin_out = min_x
    return ((x - min_x) / (max_x - min_x + eps)) * (vmax_out - vmin_out) + vmin_out


class Backdoor:
    GREY_BG_RATIO = 0.3
    STOP_SIGN_IMG = "static/stop_sign_wo_bg.png"
    CAT_IMG = "static/cat_wo_bg.png"
    GLASSES_IMG = "static/glasses.png"
    HAT_IMG = "static/fedora-hat.png"

    TARGET_TG, TARGET_CORNER, TARGET_SHIFT, TARGET_HAT, TARGET_CAT = "TRIGGER", "CORNER", "SHIFT", "HAT", "CAT"
    TRIGGER_GAP_X = TRIGGER_GAP_Y = 2
    TRIGGER_NONE = "NONE"
    TRIGGER_SM_BOX, TRIGGER_XSM_BOX, TRIGGER_XXSM_BOX, TRIGGER_XXXSM_BOX, TRIGGER_BIG_BOX = (
        "SM_BOX", "XSM_BOX", "XXSM_BOX", "XXXSM_BOX", "BIG_BOX")
    TRIGGER_BOX_18, TRIGGER_BOX_14, TRIGGER_BOX_11, TRIGGER_BOX_8, TRIGGER_BOX_4 = (
        "BOX_18", "BOX_14", "BOX_11", "BOX_8", "BOX_4")
    TRIGGER_GLASSES = "GLASSES"
    TRIGGER_STOP_SIGN_18, TRIGGER_STOP_SIGN_14, TRIGGER_STOP_SIGN_11, TRIGGER_STOP_SIGN_8, TRIGGER_STOP_SIGN_4 = (
        "STOP_SIGN_18", "STOP_SIGN_14", "STOP_SIGN_11", "STOP_SIGN_8", "STOP_SIGN_4")

    def __init__(self, root: str, static_root: Optional[str] = None):
        """`static_root`: directory that contains the reference's `static/` bitmaps (defaults to the cwd like the
        reference).  When a bitmap is not on disk the packaged pre-rendered tensors (assets/backdoor_assets.npz,
        produced from the reference by scripts/make_goldens.py) are used for the shipped sizes."""
        self.__root = root
        self.__static_root = static_root if static_root is not None else os.environ.get("BD_STATIC_ROOT", ".")

    # -- helpers (dataset.py:420-450)
    def __load_bitmap(self, rel_path: str, size, channel: int) -> Optional[torch.Tensor]:
        path = os.path.join(self.__static_root, rel_path)
        if not os.path.isfile(path):
            return None
        from PIL import Image
        from torchvision import transforms

        img = Image.open(path)
        img = img.convert("RGB") if channel == 3 else img.convert("L")
        img = transforms.Resize(size)(img)
        return normalize(x=transforms.ToTensor()(img), vmin_in=0.0, vmax_in=1.0, vmin_out=DEFAULT_VMIN, vmax_out=DEFAULT_VMAX)

    @staticmethod
    def __asset(name: str) -> torch.Tensor:
        if os.path.isfile(_ASSETS):
            with np.load(_ASSETS) as z:
                if name in z.files:
                    return torch.from_numpy(z[name].copy())
        raise FileNotFoundError(f"bitmap for {name} not found: put the reference's static/ directory under "
                                f"BD_STATIC_ROOT (or pass static_root=)")

    @staticmethod
    def __bg2grey(trig, vmin, vmax):
        thres = (vmax - vmin) * Backdoor.GREY_BG_RATIO + vmin
        trig[trig <= thres] = thres
        return trig

    @staticmethod
    def __box(b1, b2, channel, image_size, vmin, val):
        shape = (image_size, image_size) if isinstance(image_size, int) else tuple(image_size)
        trig = torch.full(size=(channel, *shape), fill_value=float(vmin))
        trig[:, b1[0]:b2[0], b1[1]:b2[1]] = val
        return trig

    @staticmethod
    def __trig_box_coord(x: int, y: int):
        if x < 0 or y < 0:
            raise ValueError("Argument x, y should > 0")
        return (-(y + Backdoor.TRIGGER_GAP_Y), -(x + Backdoor.TRIGGER_GAP_X)), (-Backdoor.TRIGGER_GAP_Y, -Backdoor.TRIGGER_GAP_X)

    def __img_trigger(self, rel_path, name, image_size, channel, trigger_sz, vmin, x=None, y=None):
        """dataset.py:472-497."""
        trig = self.__load_bitmap(rel_path, trigger_sz, channel)
        if trig is None:
            return Backdoor.__asset(f"trigger_{name}_{image_size}")
        l_pad = t_pad = int((image_size - trigger_sz) / 2)
        r_pad = image_size - trigger_sz - l_pad
        b_pad = image_size - trigger_sz - t_pad
        residual = image_size - trigger_sz
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
        trig = F.pad(trig, (l_pad, r_pad, t_pad, b_pad), value=vmin)
        trig[trig >= 0.999] = vmin
        return trig

    def get_trigger(self, type: str, channel: int, image_size: int, vmin=DEFAULT_VMIN, vmax=DEFAULT_VMAX) -> torch.Tensor:
        """dataset.py:526-597 (MNIST / FashionMNIST triggers need a dataset download and are out of scope)."""
        white = {self.TRIGGER_SM_BOX: 14, self.TRIGGER_XSM_BOX: 11, self.TRIGGER_XXSM_BOX: 8, self.TRIGGER_XXXSM_BOX: 4,
                 self.TRIGGER_BIG_BOX: 18}
        grey = {self.TRIGGER_BOX_18: 18, self.TRIGGER_BOX_14: 14, self.TRIGGER_BOX_11: 11, self.TRIGGER_BOX_8: 8,
                self.TRIGGER_BOX_4: 4}
        if type in white or type in grey:
            k = white.get(type, grey.get(type))
            b1, b2 = Backdoor.__trig_box_coord(k, k)
            val = vmax if type in white else (vmin + vmax) / 2
            return Backdoor.__box(b1, b2, channel, image_size, vmin, val)
        if type == self.TRIGGER_GLASSES:
            return self.__img_trigger(self.GLASSES_IMG, "GLASSES", image_size, channel, int(image_size * 0.625), vmin)
        if type.startswith("STOP_SIGN_"):
            return self.__img_trigger(self.STOP_SIGN_IMG, type, image_size, channel, int(type.split("_")[2]), vmin, x=-2, y=-2)
        if type == self.TRIGGER_NONE:
            return torch.full(size=(channel, image_size, image_size), fill_value=float(vmin))
        if type in ("FASHION", "FASHION_EZ", "MNIST", "MNIST_EZ"):
            raise NotImplementedError(f"trigger {type} needs a dataset download (out of scope: no network)")
        raise ValueError(f"Trigger type {type} isn't found")

    def get_target(self, type: str, trigger: torch.Tensor = None, dx: int = -5, dy: int = -3, vmin=DEFAULT_VMIN,
                   vmax=DEFAULT_VMAX) -> torch.Tensor:
        """dataset.py:627-655."""
        channel, image_size = trigger.shape[-3], list(trigger.shape[-2:])
        if type == self.TARGET_TG:
            return Backdoor.__bg2grey(trigger.clone().detach(), vmin, vmax)
        if type == self.TARGET_SHIFT:
            return Backdoor.__bg2grey(torch.roll(trigger.clone().detach(), shifts=(0, dy, dx), dims=(0, 1, 2)), vmin, vmax)
        if type == self.TARGET_CORNER:
            return Backdoor.__bg2grey(Backdoor.__box((None, None), (10, 10), channel, image_size, vmin, (vmin + vmax) / 2), vmin, vmax)
        if type in (self.TARGET_HAT, self.TARGET_CAT):
            img = self.__load_bitmap(self.HAT_IMG if type == self.TARGET_HAT else self.CAT_IMG, image_size, channel)
            if img is None:
                return Backdoor.__asset(f"target_{type}_{image_size[0]}")
            return Backdoor.__bg2grey(img, vmin, vmax)
        if type == "SHOE":
            raise NotImplementedError("target SHOE needs a dataset download (out of scope: no network)")
        raise NotImplementedError(f"Target type {type} isn't found")


def get_mask(trigger: torch.Tensor, vmin: float = DEFAULT_VMIN) -> torch.Tensor:
    """DatasetLoader.get_mask, dataset.py:275-276."""
    return torch.where(trigger > vmin, 0, 1)


@dataclass
class PoisonedBatch:
    """What the training step consumes instead of the reference's materialised {pixel_values, target} pair:
    raw images + per-sample poison flags; R and x0 are formed inside the batch-prep kernel."""
    image: torch.Tensor       # (B,C,H,W) fp32 in [-1,1]
    is_poison: torch.Tensor   # (B,) uint8


class SyntheticDataset:
    """image = randn(B,3,S,S, gen(seed)).clamp(-1,1); is_poison[i] = (i mod round(1/rate) == 0)."""

    def __init__(self, image_size: int, channels: int = 3, poison_rate: float = 0.1, n: int = 50000, seed: int = 0,
                 trigger: str = Backdoor.TRIGGER_BOX_14, target: str = Backdoor.TARGET_HAT, static_root: Optional[str] = None):
        self.image_size, self.channels, self.poison_rate, self.n, self.seed = image_size, channels, poison_rate, n, seed
        bd = Backdoor(root="datasets", static_root=static_root)
        self.trigger = bd.get_trigger(type=trigger, channel=channels, image_size=image_size)
        self.target = bd.get_target(type=target, trigger=self.trigger)
        self.every = int(round(1.0 / poison_rate)) if poison_rate > 0 else 0

    def batch(self, batch_size: int, index: int = 0, pin: bool = True) -> PoisonedBatch:
        g = torch.Generator().manual_seed(self.seed + index)
        img = torch.randn(batch_size, self.channels, self.image_size, self.image_size, generator=g).clamp(-1, 1)
        isp = torch.tensor([(self.every > 0 and (i % self.every == 0)) for i in range(batch_size)], dtype=torch.uint8)
        if pin and torch.cuda.is_available():
            img, isp = img.pin_memory(), isp.pin_memory()
        return PoisonedBatch(img, isp)


# ------------------------------------------------------------------------------------------------
# GPU data path: the host hands over DECODED pixels and coins, the device does the rest
# ------------------------------------------------------------------------------------------------
def draw_flips(n: int, generator: Optional[torch.Generator] = None, p: float = 0.5) -> torch.Tensor:
    """The coins of the reference's `transforms.RandomHorizontalFlip()` (dataset.py:126-128): one `torch.rand(1) < p`
    per image in batch order (generator None = the global generator, exactly the stream torchvision consumes)."""
    return torch.tensor([bool(torch.rand(1, generator=generator) < p) for _ in range(n)], dtype=torch.uint8)


@dataclass
class PoisonedBatchU8:
    """Decoded uint8 NHWC pixels + h-flip coins + poison flags: what `Trainer.step_u8` consumes.  ToTensor, normalize
    (Q7), the flip, the mask/blend and add_noise all run inside `bd_batch_prep_u8` (4x fewer H2D bytes than fp32 NCHW,
    none of the reference's 8 PIL/torchvision worker processes past the decode)."""
    image_u8: torch.Tensor    # (B,H,W,C) uint8
    flip: torch.Tensor        # (B,) uint8
    is_poison: torch.Tensor   # (B,) uint8


def u8_batch_to_image(image_u8: torch.Tensor, flip: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Device-side equivalent of the reference's per-image transform chain for already-sized RGB images (dataset.py:
    120-136): (B,H,W,C) uint8 on the GPU -> (B,C,H,W) fp32 in [-1, 1-2e-5], bit-exact (`bd_batch_prep_u8`'s image_out)."""
    from . import ops
    from .schedulers import DDPMScheduler

    if image_u8.device.type != "cuda":
        raise RuntimeError("u8_batch_to_image runs on the GPU (no CPU path): move the uint8 batch to the device")
    B, H, W, C = image_u8.shape
    dev = image_u8.device
    sched = DDPMScheduler()
    sched._to_device(dev)
    out = torch.empty(B, C, H, W, device=dev)
    t = torch.zeros(B, dtype=torch.int64, device=dev)
    ops.batch_prep_u8(image_u8.contiguous(), None if flip is None else flip.to(dev, torch.uint8), None, None, None, t,
                      sched._alphas_dev, sched._acp_dev, noise=torch.zeros(B, C, H, W, device=dev), image_out=out)
    return out
