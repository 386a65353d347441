#!/usr/bin/env python
"""Generates tests/golden/*.npz from the REFERENCE's own code (IBM/BadDiffusion + vendored diffusers),
imported read-only from /root/reference through oracle/ref_shim.py.

Run in the build container only (the reference tree is not on the GPU box):
    PYTHONDONTWRITEBYTECODE=1 python scripts/make_goldens.py
The fixtures are small, committed, and are what both the oracle restatement and the CUDA path are
checked against.  Weights come from oracle.torch_ref.make_state_dict (per-key seeded generators) and are
loaded into the reference modules with load_state_dict(strict=True), so no weights are stored.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim, torch_ref as O  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
ASSETS = os.path.join(ROOT, "baddiffusion_b200", "assets")
os.makedirs(OUT, exist_ok=True)
os.makedirs(ASSETS, exist_ok=True)
R = ref_shim.load()
torch.set_num_threads(8)


def cfg_kwargs(cfg):
    return {k: (tuple(v) if isinstance(v, (list, tuple)) else v) for k, v in cfg.items()}


def ref_unet(cfg, seed):
    m = R.UNet2DModel(**cfg_kwargs(cfg))
    sd = O.make_state_dict(cfg, seed)
    ref_keys = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    mine = {k: tuple(v) for k, v in O.unet_param_shapes(cfg).items()}
    assert ref_keys == mine, "param inventory differs from the reference"
    m.load_state_dict(sd, strict=True)
    return m.eval(), sd


def synth_batch(B, S, seed=0, poison_every=10):
    """SURVEY.md section 8(d) synthetic inputs."""
    image = torch.randn(B, 3, S, S, generator=torch.Generator().manual_seed(seed)).clamp(-1, 1)
    is_poison = torch.tensor([(i % poison_every == 0) if poison_every else False for i in range(B)])
    t = torch.randint(0, 1000, (B,), generator=torch.Generator().manual_seed(seed + 1))
    noise = torch.randn(B, 3, S, S, generator=torch.Generator().manual_seed(seed + 2))
    return image, is_poison, t, noise


def save(name, **arrs):
    arrs = {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in arrs.items()}
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrs)
    print("wrote", name, {k: v.shape for k, v in arrs.items()})


# ---------------------------------------------------------------------------------------------
# 0. parameter inventories (state_dict keys + shapes, Appendix D)
# ---------------------------------------------------------------------------------------------
inv = {}
for name, cfg in (("tiny", O.TINY_CONFIG), ("cifar10", O.CIFAR10_CONFIG), ("celebahq", O.CELEBAHQ_CONFIG)):
    m = R.UNet2DModel(**cfg_kwargs(cfg))
    inv[name] = {k: list(v.shape) for k, v in m.state_dict().items()}
    assert {k: tuple(v) for k, v in inv[name].items()} == {k: tuple(v) for k, v in O.unet_param_shapes(cfg).items()}
    print(name, len(inv[name]), "tensors", sum(int(np.prod(v)) for v in inv[name].values()), "params")
    del m
with open(os.path.join(OUT, "param_inventory.json"), "w") as f:
    json.dump(inv, f)

# ---------------------------------------------------------------------------------------------
# 1. Backdoor trigger / target tensors (dataset.py Backdoor) -- also shipped as product assets
# ---------------------------------------------------------------------------------------------
bd = R.Backdoor(root="/tmp/bd_datasets")
trig = {}
with ref_shim.chdir_ref():
    import io, contextlib
    with contextlib.redirect_stdout(io.StringIO()):
        for S in (32, 256):
            for kind in ("BOX_14", "BOX_8", "SM_BOX", "STOP_SIGN_14", "GLASSES", "NONE"):
                trig[f"trigger_{kind}_{S}"] = bd.get_trigger(type=kind, channel=3, image_size=S)
            base = trig[f"trigger_BOX_14_{S}"]
            for kind in ("HAT", "CAT", "CORNER", "TRIGGER", "SHIFT"):
                trig[f"target_{kind}_{S}"] = bd.get_target(type=kind, trigger=base)
save("backdoor_tensors", **trig)
# product assets: the two configurations of BASELINE.json (BOX_14->HAT @32, GLASSES->CAT @256) + extras
np.savez_compressed(os.path.join(ASSETS, "backdoor_assets.npz"),
                    **{k: v.numpy() for k, v in trig.items()
                       if k in ("target_HAT_32", "target_CAT_32", "target_HAT_256", "target_CAT_256",
                                "trigger_GLASSES_32", "trigger_GLASSES_256", "trigger_STOP_SIGN_14_32",
                                "trigger_STOP_SIGN_14_256")})

# ---------------------------------------------------------------------------------------------
# 2. poison blend + q_sample_diffuser (dataset.py:288-315, loss.py:257-285)
# ---------------------------------------------------------------------------------------------
sched = R.DDPMScheduler(num_train_timesteps=1000, beta_schedule="linear", variance_type="fixed_large")
image, is_poison, t, noise = synth_batch(16, 32, seed=0, poison_every=4)
g, y = trig["trigger_BOX_14_32"], trig["target_HAT_32"]
mask = torch.where(g > -1.0, 0, 1)  # DatasetLoader.get_mask, dataset.py:275-276
R_ref = torch.where(is_poison.view(-1, 1, 1, 1), mask * image + (1 - mask) * g, torch.zeros_like(image))
x0_ref = torch.where(is_poison.view(-1, 1, 1, 1), y.expand_as(image), image)
x_noisy, target = R.q_sample_diffuser(sched, x_start=x0_ref, R=R_ref, timesteps=t, noise=noise)
save("q_sample", image=image, is_poison=is_poison, t=t, noise=noise, R=R_ref, x0=x0_ref, x_noisy=x_noisy,
     target=target, alphas=sched.alphas, alphas_cumprod=sched.alphas_cumprod)

# ---------------------------------------------------------------------------------------------
# 3. scheduler steps
# ---------------------------------------------------------------------------------------------
gs = torch.Generator().manual_seed(7)
xs = torch.randn(4, 3, 32, 32, generator=gs)
es = torch.randn(4, 3, 32, 32, generator=gs)
steps = {}
for vt in ("fixed_small", "fixed_large"):
    for clip in (True, False):
        for nsteps in (1000, 50):
            s = R.DDPMScheduler(variance_type=vt, clip_sample=clip)
            s.set_timesteps(nsteps)
            for t_ in (int(s.timesteps[0]), int(s.timesteps[len(s.timesteps) // 2]), int(s.timesteps[-2]), 0):
                gen = torch.Generator().manual_seed(11)
                out = s.step(es, t_, xs, generator=gen).prev_sample
                steps[f"ddpm_{vt}_{int(clip)}_{nsteps}_{t_}"] = out
s = R.DDPMScheduler(variance_type="fixed_small", clip_sample=False, clip_defense=True, clip_defense_range=1.0)
s.set_timesteps(1000)
steps["ddpm_clipdef_500"] = s.step(es, 500, xs, generator=torch.Generator().manual_seed(11)).prev_sample
for clip in (True, False):
    for nsteps in (50, 10):
        s = R.DDIMScheduler(clip_sample=clip)
        s.set_timesteps(nsteps)
        for t_ in (int(s.timesteps[0]), int(s.timesteps[len(s.timesteps) // 2]), 0):
            steps[f"ddim_{int(clip)}_{nsteps}_{t_}_eta0"] = s.step(es, t_, xs, eta=0.0).prev_sample
            steps[f"ddim_{int(clip)}_{nsteps}_{t_}_eta1"] = s.step(
                es, t_, xs, eta=1.0, generator=torch.Generator().manual_seed(11)).prev_sample
save("scheduler_steps", x=xs, eps=es, z_seed=11, **steps)

# ---------------------------------------------------------------------------------------------
# 4. UNet forward + p_losses_diffuser fwd/bwd (tiny and CIFAR10 configs)
# ---------------------------------------------------------------------------------------------
for name, cfg, B in (("tiny", O.TINY_CONFIG, 4), ("cifar10", O.CIFAR10_CONFIG, 2)):
    m, sd = ref_unet(cfg, seed=0)
    S = cfg["sample_size"]
    image, is_poison, t, noise = synth_batch(B, S, seed=0, poison_every=2)
    g, y = trig[f"trigger_BOX_14_{S}"], trig[f"target_HAT_{S}"]
    Rr, x0 = O.poison_blend(image, is_poison, g, y)
    with torch.no_grad():
        eps_hat = m(image, t).sample
        eps_hat_scalar_t = m(image, 37).sample  # python-int timestep broadcast, unet_2d.py:255-261
    for p in m.parameters():
        p.requires_grad_(True)
    loss = R.p_losses_diffuser(sched, model=m, x_start=x0, R=Rr, timesteps=t, noise=noise, loss_type="l2")
    loss.backward()
    grads = {k: p.grad for k, p in m.named_parameters()}
    gstats = {}
    for k, gv in grads.items():
        gstats["gnorm/" + k] = gv.norm()
        gstats["ghead/" + k] = gv.flatten()[:16]
    extra = {}
    if name == "tiny":
        extra = {"grad/" + k: gv for k, gv in grads.items()}
    else:
        for k in ("conv_in.weight", "conv_out.weight", "mid_block.attentions.0.query.weight",
                  "down_blocks.1.resnets.0.conv_shortcut.weight", "up_blocks.3.resnets.2.norm1.weight",
                  "time_embedding.linear_1.bias", "down_blocks.0.downsamplers.0.conv.bias",
                  "up_blocks.0.upsamplers.0.conv.bias"):
            extra["grad/" + k] = grads[k]
    save(f"unet_{name}", image=image, is_poison=is_poison, t=t, noise=noise, eps_hat=eps_hat,
         eps_hat_t37=eps_hat_scalar_t, loss=loss.detach(), **gstats, **extra)
    del m

# ---------------------------------------------------------------------------------------------
# 5. pipelines + batch_sampling (tiny UNet so it runs in seconds)
# ---------------------------------------------------------------------------------------------
m, sd = ref_unet(O.TINY_CONFIG, seed=0)
pipes = {}
noise16 = torch.randn(6, 3, 32, 32, generator=torch.Generator().manual_seed(0))
bd_init = noise16 + trig["trigger_BOX_14_32"][None]  # baddiffusion.py:417,515 (quirk Q8)
for vt in ("fixed_small", "fixed_large"):
    for clip in (True, False):
        s = R.DDPMScheduler(variance_type=vt, clip_sample=clip)
        pipe = R.DDPMPipeline(unet=m, scheduler=s)
        pipe.set_progress_bar_config(disable=True)
        out = pipe(batch_size=6, generator=torch.Generator().manual_seed(3), num_inference_steps=25, init=noise16,
                   output_type=None)
        pipes[f"ddpm_{vt}_{int(clip)}_25"] = out.images
s = R.DDPMScheduler(variance_type="fixed_large", clip_sample=True)
pipe = R.DDPMPipeline(unet=m, scheduler=s)
pipe.set_progress_bar_config(disable=True)
pipes["ddpm_nogen_init_1000"] = pipe(batch_size=2, generator=torch.Generator().manual_seed(5),
                                     num_inference_steps=1000, init=noise16[:2], output_type=None).images
out = pipe(batch_size=3, generator=torch.Generator().manual_seed(9), num_inference_steps=10, output_type=None,
           save_every_step=True)
pipes["ddpm_fresh_10"] = out.images
pipes["ddpm_fresh_10_movie"] = np.stack(out.movie)
pipes["ddpm_backdoor_25"] = pipe(batch_size=6, generator=torch.Generator().manual_seed(3), num_inference_steps=25,
                                 init=bd_init, output_type=None).images
# batch_sampling shares ONE rng across chunks (model.py:469-489)
pipes["batch_sampling_6_by_4"] = R.batch_sampling(
    6, _p := (lambda **kw: pipe(num_inference_steps=20, **kw)), init=noise16, max_batch_n=4,
    rng=torch.Generator().manual_seed(13))
dpipe = R.DDIMPipeline(unet=m, scheduler=s)
dpipe.set_progress_bar_config(disable=True)
pipes["ddim_8"] = dpipe(batch_size=6, num_inference_steps=8, init=noise16, output_type=None).images
pipes["ddim_10_backdoor"] = dpipe(batch_size=6, num_inference_steps=10, init=bd_init, output_type=None).images
pipes["ddim_10_eta1"] = dpipe(batch_size=6, num_inference_steps=10, init=noise16, eta=1.0,
                              generator=torch.Generator().manual_seed(21), output_type=None).images
save("pipelines_tiny", init=noise16, bd_init=bd_init, **pipes)

# ---------------------------------------------------------------------------------------------
# 6. layer-level fixtures (ResnetBlock2D, AttentionBlock incl. multi-head, Down/Upsample2D, temb)
# ---------------------------------------------------------------------------------------------
torch.manual_seed(0)
layers = {}
x = torch.randn(2, 64, 16, 16)
temb = torch.randn(2, 128)
for tag, (ci, co) in (("same", (64, 64)), ("shortcut", (64, 96))):
    blk = R.ResnetBlock2D(in_channels=ci, out_channels=co, temb_channels=128, eps=1e-6, groups=32)
    for k, v in blk.state_dict().items():
        layers[f"resnet_{tag}/{k}"] = v
    layers[f"resnet_{tag}/out"] = blk(x, temb)
for tag, hd in (("1head", None), ("8dim", 8)):
    blk = R.AttentionBlock(64, num_head_channels=hd, eps=1e-6, norm_num_groups=32)
    for k, v in blk.state_dict().items():
        layers[f"attn_{tag}/{k}"] = v
    layers[f"attn_{tag}/out"] = blk(x)
for pad in (0, 1):
    blk = R.Downsample2D(64, use_conv=True, out_channels=64, padding=pad, name="op")
    for k, v in blk.state_dict().items():
        layers[f"down_pad{pad}/{k}"] = v
    layers[f"down_pad{pad}/out"] = blk(x)
blk = R.Upsample2D(64, use_conv=True, out_channels=64)
for k, v in blk.state_dict().items():
    layers[f"up/{k}"] = v
layers["up/out"] = blk(x)
tt = torch.tensor([0, 1, 37, 500, 999])
layers["temb/t"] = tt
layers["temb/flip0_shift1"] = R.get_timestep_embedding(tt, 128, flip_sin_to_cos=False, downscale_freq_shift=1)
layers["temb/flip1_shift0"] = R.get_timestep_embedding(tt, 128, flip_sin_to_cos=True, downscale_freq_shift=0)
save("layers", x=x, temb_in=temb, **{k: v.detach() for k, v in layers.items()})

# ---------------------------------------------------------------------------------------------
# 7. cosine LR schedule (D/optimization.py:109-141)
# ---------------------------------------------------------------------------------------------
opt = torch.optim.Adam([torch.nn.Parameter(torch.zeros(1))], lr=2e-4)
ls = R.get_cosine_schedule_with_warmup(opt, num_warmup_steps=500, num_training_steps=2000)
lrs = []
for i in range(2000):
    lrs.append(ls.get_last_lr()[0])
    opt.step()
    ls.step()
save("cosine_lr", lrs=np.array(lrs, dtype=np.float64))
print("done")
