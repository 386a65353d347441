#!/usr/bin/env python
"""Round-2 fixtures, generated from the REFERENCE's own code like scripts/make_goldens.py (build container only):
    PYTHONDONTWRITEBYTECODE=1 python scripts/make_goldens_r2.py
  * data_path.npz -- SURVEY 8f n2: a synthetic decoded uint8 batch pushed through the reference DatasetLoader's own
    transform closures (dataset.py:120-136 `__get_transform`, :288-315 `clean_transforms` / `backdoor_transforms`),
    called on a DatasetLoader instance created WITHOUT its constructor (which would download the dataset).
"""
import os
import sys

import numpy as np
import torch
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
R = ref_shim.load()
DL = R.DatasetLoader


def loader_stub(name, S, trigger, target):
    dl = object.__new__(DL)
    for k, v in dict(name=name, dataset=name, channel=3, image_size=S, vmin=-1.0, vmax=1.0, trigger=trigger, target=target).items():
        setattr(dl, "_DatasetLoader__" + k, v)
    return dl


out = {}
bd = R.Backdoor(root="/tmp/bd_datasets")
for tag, name, S, B, trig_k, targ_k in (("cifar", DL.CIFAR10, 32, 12, "BOX_14", "HAT"), ("celeba", DL.CELEBA, 64, 6, "GLASSES", "CAT")):
    with ref_shim.chdir_ref():
        import contextlib, io
        with contextlib.redirect_stdout(io.StringIO()):
            trigger = bd.get_trigger(type=trig_k, channel=3, image_size=S)
            target = bd.get_target(type=targ_k, trigger=trigger)
    out[f"{tag}/trigger"], out[f"{tag}/target_tensor"] = trigger.numpy(), target.numpy()
    dl = loader_stub(name, S, trigger, target)
    rng = np.random.RandomState(5 + S)
    u8 = rng.randint(0, 256, size=(B, S, S, 3), dtype=np.uint8)
    u8[0, :4, :4] = 0
    u8[0, 4:8, :4] = 255          # both ends of the value range are present
    key = "img" if name == DL.CIFAR10 else "image"
    for clean in (True, False):
        fn = DL._DatasetLoader__transform_generator(dl, name, clean)
        torch.manual_seed(1234)   # RandomHorizontalFlip draws torch.rand(1) per image from the global generator
        ex = fn({key: [Image.fromarray(a) for a in u8]})
        kind = "clean" if clean else "backdoor"
        out[f"{tag}/{kind}/image"] = ex[DL.IMAGE].numpy()
        out[f"{tag}/{kind}/pixel_values"] = ex[DL.PIXEL_VALUES].numpy()
        out[f"{tag}/{kind}/target"] = ex[DL.TARGET].numpy()
    # which samples were flipped (same seed -> same coins for both closures): recover from the clean output
    torch.manual_seed(1234)
    flips = np.array([bool(torch.rand(1) < 0.5) for _ in range(B)])
    # sanity: un-flipping restores the plain normalisation of the uint8 data
    img = out[f"{tag}/clean/image"]
    plain = R.normalize(vmin_in=0, vmax_in=1, vmin_out=-1.0, vmax_out=1.0, x=torch.from_numpy(u8).permute(0, 3, 1, 2).float().div(255)).numpy()
    for b in range(B):
        assert np.array_equal(img[b, :, :, ::-1] if flips[b] else img[b], plain[b]), (tag, b)
    out[f"{tag}/u8"] = u8
    out[f"{tag}/flips"] = flips
    out[f"{tag}/seed"] = np.int64(1234)
    print(tag, "flips", flips.astype(int))
np.savez_compressed(os.path.join(OUT, "data_path.npz"), **out)
print("wrote data_path", {k: v.shape for k, v in out.items()})
