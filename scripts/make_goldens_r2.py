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

# ---------------------------------------------------------------------------------------------
# the sampler behind --sched DPM_SOLVER_* / UNIPC / PNDM / DEIS / HEUN / LMSD.  model.py:598-630 pairs all
# of them with the reference's patched PNDMPipeline, which rebuilds a PNDMScheduler from the given scheduler's config.
# Fixtures: (a) single PNDMScheduler steps (step_prk / step_plms incl. the skip_prk_steps branches) on random tensors,
# (b) the pipeline on the tiny UNet, built exactly like model.py does for DPM_SOLVER_PP_O2-SCHED (clip on / off).
# ---------------------------------------------------------------------------------------------
from oracle import torch_ref as O  # noqa: E402
from diffusers.schedulers.scheduling_pndm import PNDMScheduler  # noqa: E402  (reference tree, via the shim's package stubs)
from diffusers.schedulers.scheduling_dpmsolver_multistep import DPMSolverMultistepScheduler  # noqa: E402
from diffusers.pipelines.pndm.pipeline_pndm import PNDMPipeline  # noqa: E402

pn = {}
gs = torch.Generator().manual_seed(17)
for skip in (False, True):
    for nsteps in (50, 20):
        s = PNDMScheduler(num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, skip_prk_steps=skip)
        s.set_timesteps(nsteps)
        x = torch.randn(2, 3, 8, 8, generator=gs)
        tag = f"steps_skip{int(skip)}_{nsteps}"
        pn[f"{tag}/x0"] = x
        pn[f"{tag}/timesteps"] = s.timesteps.numpy()
        n_run = 20 if not skip else 8
        eps_all, out_all = [], []
        for i, t in enumerate(s.timesteps[:n_run]):
            eps = torch.randn(2, 3, 8, 8, generator=gs)
            x = s.step(eps, t, x).prev_sample
            eps_all.append(eps)
            out_all.append(x)
        pn[f"{tag}/eps"] = torch.stack(eps_all)
        pn[f"{tag}/out"] = torch.stack(out_all)

cfg = O.TINY_CONFIG
m = R.UNet2DModel(**{k: (tuple(v) if isinstance(v, (list, tuple)) else v) for k, v in cfg.items()})
m.load_state_dict(O.make_state_dict(cfg, 0), strict=True)
m.eval()
init = torch.randn(4, 3, 32, 32, generator=torch.Generator().manual_seed(0))
for clip in (False, True):
    sched = DPMSolverMultistepScheduler(num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, solver_order=2,
                                        algorithm_type="dpmsolver++")
    pipe = PNDMPipeline(unet=m, scheduler=sched, clip_sample=clip)
    pipe.set_progress_bar_config(disable=True)
    res = pipe(batch_size=4, num_inference_steps=20, init=init, output_type=None, save_every_step=True)
    pn[f"pipe_clip{int(clip)}_20/images"] = res.images
    pn[f"pipe_clip{int(clip)}_20/movie_last3"] = np.stack(res.movie[-3:])
    assert type(pipe.scheduler).__name__ == "PNDMScheduler"
pn["pipe/init"] = init
np.savez_compressed(os.path.join(OUT, "pndm.npz"), **{k: (v.detach().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in pn.items()})
print("wrote pndm", {k: np.asarray(v).shape for k, v in pn.items()})
