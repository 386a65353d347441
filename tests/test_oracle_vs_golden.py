"""Pins the CPU oracle (oracle/torch_ref.py) against
  (1) fixtures produced by the reference's own code (scripts/make_goldens.py -> tests/golden), and
  (2) the reference's hard-coded known-answer values from its vendored test-suite
      (diffusers/tests/schedulers/test_scheduler_ddpm.py:62-100, test_scheduler_ddim.py:46-54,94-140,
       diffusers/tests/models/test_layers_utils.py:92-117).
CPU only; runs in the `-m "not gpu"` suite."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import torch_ref as O

T = torch.from_numpy


def test_param_inventory_matches_reference():
    with open(os.path.join(os.path.dirname(__file__), "golden", "param_inventory.json")) as f:
        inv = json.load(f)
    for name, cfg in (("tiny", O.TINY_CONFIG), ("cifar10", O.CIFAR10_CONFIG), ("celebahq", O.CELEBAHQ_CONFIG)):
        mine = {k: list(v) for k, v in O.unet_param_shapes(cfg).items()}
        assert mine == inv[name]
    assert sum(int(np.prod(v)) for v in inv["cifar10"].values()) == 35_746_307
    assert sum(int(np.prod(v)) for v in inv["celebahq"].values()) == 113_673_219


def test_backdoor_tensors(golden):
    g = golden("backdoor_tensors")
    for S in (32, 256):
        for kind in ("BOX_14", "BOX_8", "SM_BOX", "NONE"):
            assert torch.equal(O.get_trigger(kind, S), T(g[f"trigger_{kind}_{S}"]))
        base = O.get_trigger("BOX_14", S)
        for kind in ("CORNER", "TRIGGER", "SHIFT"):
            assert torch.equal(O.get_target(kind, base), T(g[f"target_{kind}_{S}"]))
    t = T(g["trigger_BOX_14_32"])
    assert int((t > -1).sum()) == 588 and set(t.unique().tolist()) == {-1.0, 0.0}
    hat = T(g["target_HAT_32"])
    assert abs(float(hat.min()) + 0.4) < 1e-6 and float(hat.max()) < 1.0


@pytest.mark.skipif(not os.path.isdir("/root/reference/static"), reason="reference bitmaps not mounted")
def test_backdoor_bitmap_tensors(golden):
    g = golden("backdoor_tensors")
    sd = "/root/reference/static"
    for S in (32, 256):
        assert torch.equal(O.get_trigger("GLASSES", S, sd), T(g[f"trigger_GLASSES_{S}"]))
        assert torch.equal(O.get_trigger("STOP_SIGN_14", S, sd), T(g[f"trigger_STOP_SIGN_14_{S}"]))
        base = O.get_trigger("BOX_14", S)
        assert torch.equal(O.get_target("HAT", base, sd), T(g[f"target_HAT_{S}"]))
        assert torch.equal(O.get_target("CAT", base, sd), T(g[f"target_CAT_{S}"]))


def test_q_sample_bit_exact(golden):
    g = golden("q_sample")
    bt = golden("backdoor_tensors")
    betas, alphas, acp = O.beta_tables()
    assert torch.equal(alphas, T(g["alphas"])) and torch.equal(acp, T(g["alphas_cumprod"]))
    R, x0 = O.poison_blend(T(g["image"]), T(g["is_poison"]), T(bt["trigger_BOX_14_32"]), T(bt["target_HAT_32"]))
    assert torch.equal(R, T(g["R"])) and torch.equal(x0, T(g["x0"]))
    xn, tgt = O.q_sample(alphas, acp, x0, R, T(g["t"]), T(g["noise"]))
    assert torch.equal(xn, T(g["x_noisy"])) and torch.equal(tgt, T(g["target"]))
    clean = ~T(g["is_poison"])
    assert torch.equal(tgt[clean], T(g["noise"])[clean])  # clean rows: target == eps exactly
    # rho_t spot values, SURVEY 8(a) a6
    rho = (1 - alphas ** 0.5) * (1 - acp) ** 0.5 / (1 - alphas)
    for t, v in ((0, 0.0050004), (10, 0.023417), (500, 0.48137), (999, 0.50251)):
        assert abs(float(rho[t]) - v) < 5e-6


def test_scheduler_steps_bit_exact(golden):
    g = golden("scheduler_steps")
    _, _, acp = O.beta_tables()
    x, eps = T(g["x"]), T(g["eps"])
    n = 0
    for key, val in g.items():
        if key.startswith("ddpm_fixed"):
            _, _, vt2, clip, nsteps, t = key.split("_")
            z = torch.randn(x.shape, generator=torch.Generator().manual_seed(11))
            out = O.ddpm_step(acp, eps, int(t), x, z, int(nsteps), variance_type="fixed_" + vt2,
                              clip_sample=bool(int(clip)))
        elif key == "ddpm_clipdef_500":
            z = torch.randn(x.shape, generator=torch.Generator().manual_seed(11))
            out = O.ddpm_step(acp, eps, 500, x, z, 1000, clip_sample=False, clip_defense=True)
        elif key.startswith("ddim_"):
            _, clip, nsteps, t, eta = key.split("_")
            eta = float(eta[3:])
            z = torch.randn(x.shape, generator=torch.Generator().manual_seed(11)) if eta > 0 else None
            out = O.ddim_step(acp, eps, int(t), x, int(nsteps), eta=eta, z=z, clip_sample=bool(int(clip)))
        else:
            continue
        assert torch.equal(out, T(val)), key
        n += 1
    assert n >= 40


# --- reference KATs (hard-coded numbers from the reference's own tests) ---------------------------
def _dummy_sample_deter():  # T/schedulers/test_schedulers.py:225-236
    batch_size, num_channels, height, width = 4, 3, 8, 8
    num_elems = batch_size * num_channels * height * width
    sample = torch.arange(num_elems).reshape(num_channels, height, width, batch_size) / num_elems
    return sample.permute(3, 0, 1, 2)


def _dummy_model(sample, t):  # T/schedulers/test_schedulers.py:238-243
    return sample * t / (t + 1)


def test_kat_ddpm_variance():
    _, _, acp = O.beta_tables()

    def var(t):
        prev = acp[t - 1] if t > 0 else torch.tensor(1.0)
        return torch.clamp((1 - prev) / (1 - acp[t]) * (1 - acp[t] / prev), min=1e-20)

    assert abs(float(var(0)) - 0.0) < 1e-5
    assert abs(float(var(487)) - 0.00979) < 1e-5
    assert abs(float(var(999)) - 0.02) < 1e-5


def test_kat_ddpm_full_loop():
    """T/schedulers/test_scheduler_ddpm.py:71-100: sum 258.9606, mean 0.3372."""
    _, _, acp = O.beta_tables()
    sample = _dummy_sample_deter()
    gen = torch.manual_seed(0)
    for t in reversed(range(1000)):
        residual = _dummy_model(sample, t)
        z = torch.randn(residual.shape, generator=gen) if t > 0 else None
        sample = O.ddpm_step(acp, residual, t, sample, z, 1000, variance_type="fixed_small", clip_sample=True)
    assert abs(float(sample.abs().sum()) - 258.9606) < 1e-2
    assert abs(float(sample.abs().mean()) - 0.3372) < 1e-3


def test_kat_ddim_timesteps_and_loop():
    """T/schedulers/test_scheduler_ddim.py:94-140: 10 steps, eta 0 -> 172.0067 / 0.223967."""
    assert O.timesteps_for(10).tolist() == [900, 800, 700, 600, 500, 400, 300, 200, 100, 0]
    assert O.timesteps_for(50)[0] == 980
    _, _, acp = O.beta_tables()
    sample = _dummy_sample_deter()
    for t in O.timesteps_for(10):
        residual = _dummy_model(sample, int(t))
        sample = O.ddim_step(acp, residual, int(t), sample, 10, eta=0.0, clip_sample=True)
    assert abs(float(sample.abs().sum()) - 172.0067) < 1e-2
    assert abs(float(sample.abs().mean()) - 0.223967) < 1e-3


def test_kat_sinusoid():
    """T/models/test_layers_utils.py:92-117 (embedding_dim 32, t = arange(10))."""
    t = torch.arange(10)
    e = O.timestep_embedding(t, 32, flip_sin_to_cos=False, freq_shift=1)
    assert (e[:, 0] - torch.sin(t.float())).abs().max() < 1e-5
    e1 = O.timestep_embedding(t, 64, flip_sin_to_cos=False, freq_shift=1)
    assert abs(float(e1[1, 0]) - 0.8415) < 1e-3 and abs(float(e1[1, 32]) - 0.5403) < 1e-3


def test_layers(golden):
    g = golden("layers")
    x, temb = T(g["x"]), T(g["temb_in"])
    for tag in ("same", "shortcut"):
        sd = {k.split("/", 1)[1]: T(v) for k, v in g.items() if k.startswith(f"resnet_{tag}/") and not k.endswith("/out")}
        out = O.resnet_block(sd, "", x, temb, 32, 1e-6)
        assert (out - T(g[f"resnet_{tag}/out"])).abs().max() < 1e-5
    for tag, hd in (("1head", None), ("8dim", 8)):
        sd = {k.split("/", 1)[1]: T(v) for k, v in g.items() if k.startswith(f"attn_{tag}/") and not k.endswith("/out")}
        out = O.attention_block(sd, "", x, 32, 1e-6, hd)
        assert (out - T(g[f"attn_{tag}/out"])).abs().max() < 1e-5
    for pad in (0, 1):
        sd = {k.split("/", 1)[1]: T(v) for k, v in g.items() if k.startswith(f"down_pad{pad}/") and not k.endswith("/out")}
        assert (O.downsample(sd, "", x, pad) - T(g[f"down_pad{pad}/out"])).abs().max() < 1e-5
    sd = {k.split("/", 1)[1]: T(v) for k, v in g.items() if k.startswith("up/") and not k.endswith("/out")}
    assert (O.upsample(sd, "", x) - T(g["up/out"])).abs().max() < 1e-5
    tt = T(g["temb/t"])
    assert torch.allclose(O.timestep_embedding(tt, 128, False, 1), T(g["temb/flip0_shift1"]), atol=1e-6)
    assert torch.allclose(O.timestep_embedding(tt, 128, True, 0), T(g["temb/flip1_shift0"]), atol=1e-6)


@pytest.mark.parametrize("name,cfg", [("tiny", O.TINY_CONFIG), ("cifar10", O.CIFAR10_CONFIG)])
def test_unet_forward_and_loss(golden, name, cfg):
    g = golden(f"unet_{name}")
    sd = O.make_state_dict(cfg, 0)
    image, t = T(g["image"]), T(g["t"])
    with torch.no_grad():
        eps_hat = O.unet_forward(sd, cfg, image, t)
        eps37 = O.unet_forward(sd, cfg, image, 37)
    assert ((eps_hat - T(g["eps_hat"])) ** 2).mean() < 1e-10
    assert ((eps37 - T(g["eps_hat_t37"])) ** 2).mean() < 1e-10
    bt = golden("backdoor_tensors")
    S = cfg["sample_size"]
    R, x0 = O.poison_blend(image, T(g["is_poison"]), T(bt[f"trigger_BOX_14_{S}"]), T(bt[f"target_HAT_{S}"]))
    _, alphas, acp = O.beta_tables()
    if name == "tiny":  # backward through the oracle (autograd) vs the reference's gradients
        sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        loss = O.p_losses(sdg, cfg, alphas, acp, x0, R, t, T(g["noise"]))
        loss.backward()
        assert abs(float(loss) - float(g["loss"])) < 1e-6 * max(1.0, abs(float(g["loss"])))
        for k, v in sdg.items():
            ref = T(g["grad/" + k])
            assert (v.grad - ref).abs().max() <= 1e-5 * max(1.0, float(ref.abs().max())), k
    else:
        with torch.no_grad():
            loss = O.p_losses(sd, cfg, alphas, acp, x0, R, t, T(g["noise"]))
        assert abs(float(loss) - float(g["loss"])) < 1e-5 * max(1.0, abs(float(g["loss"])))


def test_pipelines_tiny(golden):
    g = golden("pipelines_tiny")
    cfg = O.TINY_CONFIG
    sd = O.make_state_dict(cfg, 0)
    init, bd_init = T(g["init"]), T(g["bd_init"])
    bt = golden("backdoor_tensors")
    assert torch.equal(bd_init, init + T(bt["trigger_BOX_14_32"])[None])  # quirk Q8
    tol = 2e-4  # 25 chained UNet evaluations; fp32 summation-order noise only
    for vt in ("fixed_small", "fixed_large"):
        for clip in (True, False):
            out = O.ddpm_pipeline(sd, cfg, dict(variance_type=vt, clip_sample=clip), 6,
                                  generator=torch.Generator().manual_seed(3), num_inference_steps=25, init=init)
            assert np.abs(out - g[f"ddpm_{vt}_{int(clip)}_25"]).max() < tol
    sc = dict(variance_type="fixed_large", clip_sample=True)
    out = O.ddpm_pipeline(sd, cfg, sc, 3, generator=torch.Generator().manual_seed(9), num_inference_steps=10)
    assert np.abs(out - g["ddpm_fresh_10"]).max() < tol
    out = O.ddpm_pipeline(sd, cfg, sc, 6, generator=torch.Generator().manual_seed(3), num_inference_steps=25,
                          init=bd_init)
    assert np.abs(out - g["ddpm_backdoor_25"]).max() < tol
    run = lambda bs, rng, chunk: O.ddpm_pipeline(sd, cfg, sc, bs, generator=rng, num_inference_steps=20, init=chunk)
    out = O.batch_sampling(6, run, init=init, max_batch_n=4, rng=torch.Generator().manual_seed(13))
    assert np.abs(out - g["batch_sampling_6_by_4"]).max() < tol
    # eta=0 DDIM on a random-init UNet amplifies fp32 summation-order noise ~2x per step (measured),
    # so the deterministic chains get a looser max-abs bound and a tight mean-abs bound.
    out = O.ddim_pipeline(sd, cfg, sc, 6, num_inference_steps=8, init=init)
    assert np.abs(out - g["ddim_8"]).max() < 1e-2 and np.abs(out - g["ddim_8"]).mean() < 1e-4
    out = O.ddim_pipeline(sd, cfg, sc, 6, num_inference_steps=10, init=bd_init)
    assert np.abs(out - g["ddim_10_backdoor"]).max() < 1e-2 and np.abs(out - g["ddim_10_backdoor"]).mean() < 1e-4
    out = O.ddim_pipeline(sd, cfg, sc, 6, num_inference_steps=10, init=init, eta=1.0,
                          generator=torch.Generator().manual_seed(21))
    assert np.abs(out - g["ddim_10_eta1"]).max() < 1e-2 and np.abs(out - g["ddim_10_eta1"]).mean() < 1e-4


def test_cosine_lr(golden):
    lrs = golden("cosine_lr")["lrs"]
    mine = np.array([2e-4 * O.cosine_lr_lambda(i, 500, 2000) for i in range(2000)])
    assert np.abs(mine - lrs).max() < 1e-12


def test_u8_data_path_matches_reference_transforms(golden):
    """oracle restatement of ToTensor + normalize + RandomHorizontalFlip + mask/blend vs the reference
    DatasetLoader's own transform closures (fixture: scripts/make_goldens_r2.py)."""
    g = golden("data_path")
    for tag in ("cifar", "celeba"):
        u8, flips = g[f"{tag}/u8"], g[f"{tag}/flips"]
        coins = O.draw_flips(len(u8), generator=torch.Generator().manual_seed(int(g[f"{tag}/seed"])))
        assert np.array_equal(coins.numpy(), flips)     # global-generator stream == explicit generator with that seed
        img = O.u8_batch_to_image(u8, flips)
        assert np.array_equal(img.numpy(), g[f"{tag}/clean/image"]) and np.array_equal(img.numpy(), g[f"{tag}/backdoor/image"])
        trig, targ = torch.from_numpy(g[f"{tag}/trigger"]), torch.from_numpy(g[f"{tag}/target_tensor"])
        B = len(u8)
        R, x0 = O.poison_blend(img, torch.ones(B, dtype=torch.bool), trig, targ)
        assert np.array_equal(R.numpy(), g[f"{tag}/backdoor/pixel_values"]) and np.array_equal(x0.numpy(), g[f"{tag}/backdoor/target"])
        R, x0 = O.poison_blend(img, torch.zeros(B, dtype=torch.bool), trig, targ)
        assert np.array_equal(R.numpy(), g[f"{tag}/clean/pixel_values"]) and np.array_equal(x0.numpy(), g[f"{tag}/clean/target"])
