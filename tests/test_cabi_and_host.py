"""CPU-only checks: the C-ABI library loads and exports exactly what include/b200bd.h declares (no compute
calls without a GPU); host-side scheduler scalars reproduce the reference fixtures when the kernel's arithmetic
is replayed in torch; config / checkpoint layout round-trips."""
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
T = torch.from_numpy


def _header_functions():
    hdr = open(os.path.join(ROOT, "include", "b200bd.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return set(re.findall(r"\b(bd_[a-z0-9_]+)\s*\(", hdr))


def test_library_exports_every_declared_symbol():
    from baddiffusion_b200 import _lib

    lib = _lib.load()
    declared = _header_functions()
    assert declared, "no declarations parsed"
    assert declared == set(_lib.EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.bd_version() >= 100
    assert isinstance(lib.bd_last_error(), bytes)


def test_product_fails_loudly_without_gpu():
    from baddiffusion_b200 import _lib

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.B200BDError):
        _lib.lib()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "baddiffusion_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
    for f in ("baddiffusion.py",):
        p = os.path.join(ROOT, f)
        if os.path.exists(p):
            assert not re.search(r"^\s*(from|import)\s+oracle\b", open(p).read(), flags=re.M), f


def _replay_ddpm(row, x, eps, z):
    sb, sa, c0, ct, sigma, clip, clipd, has_noise = [row[i] for i in range(8)]
    x0 = (x - sb * eps) / sa
    if clip > 0:
        x0 = x0.clamp(-float(clip), float(clip))
    out = c0 * x0 + ct * x
    if has_noise != 0:
        out = out + sigma * z
    if clipd > 0:
        out = out.clamp(-float(clipd), float(clipd))
    return out


def _replay_ddim(row, x, eps, z):
    sb, sa, sap, dirc, std, clip, reclip, _ = [row[i] for i in range(8)]
    x0 = (x - sb * eps) / sa
    pe = eps
    if clip > 0:
        x0 = x0.clamp(-float(clip), float(clip))
    if reclip != 0:
        pe = (x - sa * x0) / sb
    out = sap * x0 + dirc * pe
    if std > 0:
        out = out + std * z
    return out


def test_scheduler_host_scalars_reproduce_reference(golden):
    """The fused step kernels apply exactly this arithmetic; with the host scalars it is bit-identical to the
    reference's DDPMScheduler.step / DDIMScheduler.step outputs."""
    from baddiffusion_b200.schedulers import DDIMScheduler, DDPMScheduler

    g = golden("scheduler_steps")
    x, eps = T(g["x"]), T(g["eps"])
    z = torch.randn(x.shape, generator=torch.Generator().manual_seed(11))
    n = 0
    for key, val in g.items():
        if key.startswith("ddpm_fixed"):
            _, _, vt2, clip, nsteps, t = key.split("_")
            s = DDPMScheduler(variance_type="fixed_" + vt2, clip_sample=bool(int(clip)))
            s.set_timesteps(int(nsteps))
            out = _replay_ddpm(s.coef_row(int(t)), x, eps, z)
        elif key == "ddpm_clipdef_500":
            s = DDPMScheduler(clip_sample=False, clip_defense=True)
            s.set_timesteps(1000)
            out = _replay_ddpm(s.coef_row(500), x, eps, z)
        elif key.startswith("ddim_"):
            _, clip, nsteps, t, eta = key.split("_")
            s = DDIMScheduler(clip_sample=bool(int(clip)))
            s.set_timesteps(int(nsteps))
            out = _replay_ddim(s.coef_row(int(t), float(eta[3:])), x, eps, z)
        else:
            continue
        assert torch.equal(out, T(val)), key
        n += 1
    assert n >= 40


def test_scheduler_surface():
    from baddiffusion_b200.schedulers import DDIMScheduler, DDPMScheduler

    s = DDPMScheduler(variance_type="fixed_large")
    assert s.alphas_cumprod.shape == (1000,) and s.alphas.dtype == torch.float32
    assert abs(float(s.alphas_cumprod[0]) - 0.99990) < 1e-6 and abs(float(s.alphas_cumprod[999]) - 4.0358e-5) < 1e-8
    assert s.num_train_timesteps == 1000 and len(s) == 1000
    s.config.clip_sample = False  # model.py:640 mutates the config
    assert s.config["clip_sample"] is False
    s.set_timesteps(50)
    assert s.timesteps[0] == 980 and int(s.previous_timestep(980)) == 960
    with pytest.raises(ValueError):
        s.set_timesteps(2000)
    with pytest.raises(ValueError):
        s.set_timesteps(10, timesteps=[5, 3])
    with pytest.raises(ValueError):
        s.set_timesteps(timesteps=[3, 5])
    s.set_timesteps(timesteps=[999, 500, 0])
    assert int(s.previous_timestep(500)) == 0 and int(s.previous_timestep(0)) == -1
    d = DDIMScheduler.from_config(s.config)  # DDIMPipeline.__init__ converts like this (pipeline_ddim.py:40)
    assert d.config.clip_sample is False and "variance_type" not in d.config
    with pytest.raises(ValueError):
        d.step(torch.zeros(1), 0, torch.zeros(1))  # set_timesteps not called
    x0, noise = torch.randn(3, 3, 8, 8), torch.randn(3, 3, 8, 8)
    t = torch.tensor([0, 500, 999])
    from oracle import torch_ref as O

    assert torch.equal(s.add_noise(x0, noise, t), O.add_noise(s.alphas_cumprod, x0, noise, t))


def test_scheduler_config_roundtrip(tmp_path):
    import json

    from baddiffusion_b200.schedulers import DDPMScheduler

    s = DDPMScheduler(variance_type="fixed_large", clip_defense=True, clip_defense_range=0.8)
    s.save_pretrained(str(tmp_path))
    j = json.load(open(tmp_path / "scheduler_config.json"))
    assert j["_class_name"] == "DDPMScheduler" and j["_diffusers_version"] == "0.16.0.dev0"
    assert len([k for k in j if not k.startswith("_")]) == 14  # Appendix D: 14 DDPMScheduler args
    s2 = DDPMScheduler.from_pretrained(str(tmp_path))
    assert dict(s2.config) == dict(s.config)


@pytest.mark.parametrize("arch", ["DDPM-CIFAR10-32", "DDPM-CELEBA-HQ-256"])
def test_flat_layout_gradient_buckets_are_contiguous(arch):
    """The data-parallel trainer all-reduces the gradient buffer in address ranges that become final part-way through
    backward (up blocks first, then mid + down blocks 1.., then the rest): FlatLayout must keep the tensor-core weights
    of each of those layer groups contiguous, with no other parameter inside the range."""
    import math

    from baddiffusion_b200.model import DiffuserModelSched
    from baddiffusion_b200.unet import FlatLayout
    from baddiffusion_b200.unet import UNet2DModel

    lay = UNet2DModel(**DiffuserModelSched.ARCH[arch]).layout
    assert isinstance(lay, FlatLayout)

    def ranges(pred):
        segs = sorted((lay.offset[k], lay.offset[k] + math.prod(lay.entries[k])) for k in lay.offset
                      if pred(k) and lay.offset[k] + math.prod(lay.entries[k]) <= lay.tproj_w_offset)
        out = []
        for lo, hi in segs:
            if out and lo - out[-1][1] < lay.ALIGN:
                out[-1][1] = hi
            else:
                out.append([lo, hi])
        return out

    up = ranges(lambda k: k.startswith("up_blocks."))
    assert len(up) == 1
    nd = sum(1 for k in lay.offset if k.endswith("resnets.0.conv1.weight") and k.startswith("down_blocks."))
    late = tuple(f"down_blocks.{i}." for i in range(1, nd)) + ("mid_block.",)
    lt = ranges(lambda k: k.startswith(late))
    assert 1 <= len(lt) <= 2
    for (lo, hi), pref in [(up[0], ("up_blocks.",))] + [(r, late) for r in lt]:
        for k, o in lay.offset.items():
            assert k.startswith(pref) or not (lo <= o < hi), (k, lo, hi)
    covered = sum(hi - lo for lo, hi in up + lt)
    assert covered > 0.8 * lay.n_gemm   # the overlapped buckets carry most of the gradient bytes
