#!/usr/bin/env python
"""Benchmark of the BadDiffusion hot path (BASELINE.json): poisoned DDPM training step and DDPM / DDIM sampling.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host CPU cores
    python bench.py --workload {cifar_train|celeba_train|ddim_sample|ddpm_sample}   # which BASELINE config is the line

Workloads (config.workload), all with synthetic inputs and random-init weights of the named architecture:
  cifar_train  (default; BASELINE configs[1], N>1: configs[2]) DDPM-CIFAR10-32 UNet2DModel (35.7 M params), batch 128 per
               GPU, poison_rate 0.1, BOX_14 -> HAT.  A "step" is one full training step: batch-prep -> UNet fwd -> MSE ->
               UNet bwd -> (all-reduce) -> clip + Adam.  Weak scaling: the per-GPU batch stays 128.
  celeba_train (configs[3]) DDPM-CELEBA-HQ-256 UNet2DModel (113.7 M params), batch 4 per GPU (32 over 8 GPUs), GLASSES -> CAT.
  ddim_sample  (configs[4]) DDIM-SCHED, 50 steps, CIFAR10-32 UNet, 256 clean + 256 backdoor-init samples per GPU
               (eval_max_batch 2048 over 8 GPUs), through batch_sampling / DDIMPipeline; no collective.
  ddpm_sample  DDPM-SCHED, 1000 steps (all of them run, nothing extrapolated), 256 samples per GPU.
The default line carries the other three as sub-records ("celebahq_256", "sampling") measured at the same N, plus
`roofline` (FLOP-dominant conv kernel), `roofline_time_dominant` (GroupNorm backward, HBM-bound class), `cpu_baseline`
(oracle port on the host cores) and `gpu_incumbent` (the same reference modules in PyTorch-eager CUDA: cuDNN / cuBLAS under
fp16 autocast + channels_last, N = 1 only).

One JSON line is printed by rank 0 (see the keys below); everything else goes to stderr.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

UNIT = "images/s"
GFLOP_FWD = {"cifar": 12.444, "celeba": 497.03}   # SURVEY.md 8(d) / BASELINE.md 2: UNet forward per image, 2*MAC
WORKLOADS = ("cifar_train", "celeba_train", "ddim_sample", "ddpm_sample")
METRICS = {"cifar_train": "train_images_per_sec", "celeba_train": "train_images_per_sec",
           "ddim_sample": "ddim_50_step_samples_per_sec", "ddpm_sample": "ddpm_1000_step_samples_per_sec"}
UNITS = {"cifar_train": "images/s", "celeba_train": "images/s", "ddim_sample": "samples/s", "ddpm_sample": "samples/s"}
WORKLOAD_TEXT = {
    "cifar_train": "DDPM-CIFAR10-32 poisoned train step (BASELINE configs[1]; N>1: configs[2]): p_losses_diffuser fwd/bwd over "
                   "the google/ddpm-cifar10-32 UNet2DModel topology (35.7 M parameters, random init) + clip + Adam, synthetic 3x32x32",
    "celeba_train": "DDPM-CELEBA-HQ-256 poisoned train step (BASELINE configs[3]): p_losses_diffuser fwd/bwd over the "
                    "google/ddpm-ema-celebahq-256 UNet2DModel topology (113.7 M parameters, random init) + clip + Adam, "
                    "GLASSES -> CAT, synthetic 3x256x256",
    "ddim_sample": "DDIM-SCHED sampling (BASELINE configs[4]): 50 steps, CIFAR10-32 UNet, clean + backdoor (noise + trigger) "
                   "init, eval_max_batch 256 per GPU, batch_sampling -> DDIMPipeline",
    "ddpm_sample": "DDPM-SCHED sampling: 1000 steps, CIFAR10-32 UNet, 256 samples per GPU, batch_sampling -> DDPMPipeline",
}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


def ncu_traffic(name="roofline_kernel_traffic.json"):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of a roofline kernel, from the committed `ncu --set full`
    capture (profiles/<name>; written by scripts/ncu_traffic.py), or None."""
    p = os.path.join(ROOT, "profiles", name)
    try:
        return float(json.load(open(p))["traffic_bytes_per_launch"])
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                       "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for n, v in zip(names, r[5:9]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        if sm:
            sm.sort()
            out = {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}
        return out


# ----------------------------------------------------------------------------------------------------------
# reference arm: the reference algorithm (oracle port of the vendored-diffusers UNet + loss.p_losses_diffuser +
# baddiffusion.py:593-615 train step, DDIM / DDPM pipelines) on the host cores.  /root/reference is not on the GPU box
# and the reference is a Python repo, so kind = "port".
# ----------------------------------------------------------------------------------------------------------
def _stats(times):
    s = sorted(times)
    return {"min_s": s[0], "median_s": s[len(s) // 2], "mean_s": sum(s) / len(s), "n": len(s)}


def cpu_train_steps(batch, steps, warmup, threads=None, budget_s=None, arch="cifar"):
    import numpy as np
    import torch

    from oracle import torch_ref as O

    threads = threads or len(os.sched_getaffinity(0))
    torch.set_num_threads(threads)
    cfg, S, trig_kind, targ_kind = ((O.CIFAR10_CONFIG, 32, "BOX_14", "HAT") if arch == "cifar"
                                    else (O.CELEBAHQ_CONFIG, 256, "GLASSES", "CAT"))
    sd = {k: v.clone().requires_grad_(True) for k, v in O.make_state_dict(cfg, 0).items()}
    params = list(sd.values())
    opt = torch.optim.Adam(params, lr=2e-4)
    _, alphas, acp = O.beta_tables()
    z = np.load(os.path.join(ROOT, "baddiffusion_b200", "assets", "backdoor_assets.npz"))
    trig = O.get_trigger(trig_kind, S) if trig_kind.startswith("BOX") else torch.from_numpy(z[f"trigger_{trig_kind}_{S}"])
    targ = torch.from_numpy(z[f"target_{targ_kind}_{S}"])
    g = torch.Generator().manual_seed(0)
    times = []
    t_begin = time.perf_counter()
    for i in range(warmup + steps):
        image = torch.randn(batch, 3, S, S, generator=g).clamp(-1, 1)
        isp = torch.tensor([j % 10 == 0 for j in range(batch)])
        t0 = time.perf_counter()
        R, x0 = O.poison_blend(image, isp, trig, targ)
        noise = torch.randn(image.shape)
        t = torch.randint(0, 1000, (batch,))
        loss = O.p_losses(sd, cfg, alphas, acp, x0, R, t, noise)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(params, 1.0)
        opt.step()
        opt.zero_grad()
        _ = loss.item()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        if budget_s is not None and time.perf_counter() - t_begin > budget_s and len(times) >= 1:
            break
    return times, threads


def cpu_sample_steps(batch, nsteps, ddim, threads=None):
    """Seconds per denoise step of the oracle's DDPM / DDIM loop (UNet forward + scheduler step) on the host cores."""
    import torch

    from oracle import torch_ref as O

    threads = threads or len(os.sched_getaffinity(0))
    torch.set_num_threads(threads)
    cfg = O.CIFAR10_CONFIG
    sd = O.make_state_dict(cfg, 0)
    _, _, acp = O.beta_tables()
    x = torch.randn(batch, 3, 32, 32, generator=torch.Generator().manual_seed(0))
    ts = O.timesteps_for(50 if ddim else 1000)
    times = []
    with torch.no_grad():
        for i in range(nsteps + 1):
            t = int(ts[i])
            t0 = time.perf_counter()
            e = O.unet_forward(sd, cfg, x, t)
            if ddim:
                x = O.ddim_step(acp, e, t, x, 50)
            else:
                x = O.ddpm_step(acp, e, t, x, torch.randn(x.shape), 1000, variance_type="fixed_large", clip_sample=True)
            if i > 0:
                times.append(time.perf_counter() - t0)
    return times, threads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = args.workload
    if wl in ("cifar_train", "celeba_train"):
        # the arm's own --steps / --warmup at the FULL per-GPU batch of the workload, bounded to a few minutes of CPU time
        batch = args.ref_batch or (128 if wl == "cifar_train" else 4)
        times, threads = cpu_train_steps(batch, args.steps, min(args.warmup, 2), budget_s=float(args.ref_budget_s),
                                         arch="cifar" if wl == "cifar_train" else "celeba")
        total = sum(times)
        value = batch * len(times) / total
        sample = (f"{len(times)} full train steps (fwd+bwd+clip+Adam) at the workload's per-GPU batch {batch}, fp32, torch CPU "
                  f"(min {min(times):.2f} s, median {sorted(times)[len(times) // 2]:.2f} s per step)")
        ms = 1e3 * total / len(times)
        nsteps = len(times)
    else:
        ddim = wl == "ddim_sample"
        batch, n = 32, max(2, min(args.steps, 8))
        times, threads = cpu_sample_steps(batch, n, ddim)
        per = sum(times) / len(times)
        value = batch / (per * (50 if ddim else 1000))
        sample = (f"{len(times)} denoise steps (UNet forward + scheduler step) at batch {batch}, fp32 torch CPU; samples/s = "
                  f"batch / (s_per_step x {50 if ddim else 1000} steps)")
        ms = 1e3 * per
        nsteps = len(times)
    line = {
        "impl": "reference", "metric": METRICS[wl], "value": value, "unit": UNITS[wl], "n_gpus": args.gpus, "steps": nsteps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD_TEXT[wl]},
        "cpu_baseline": {"value": value, "unit": UNITS[wl], "cores": threads, "kind": "port", "sample": sample,
                         **_stats(times)},
        "e2e": {"value": value, "unit": UNITS[wl], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------
class Ctx:
    """Process-wide state of the GPU arm (rank / world, barrier, max-over-ranks)."""

    def __init__(self):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.pg = None
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            self.pg = dist.group.WORLD

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v):
        if self.world == 1:
            return v
        t = self.torch.tensor([v], device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t)


def measure_train(ctx, arch, K, W, e2e_steps=None):
    """Training throughput of one architecture at N GPUs: (1) device-resident, (2) end to end through Trainer.step with
    pinned host batches.  Returns a dict on every rank (times are the max over ranks)."""
    torch = ctx.torch
    from baddiffusion_b200 import _lib
    from baddiffusion_b200.dataset import SyntheticDataset
    from baddiffusion_b200.model import DiffuserModelSched
    from baddiffusion_b200.schedulers import DDPMScheduler
    from baddiffusion_b200.train import Trainer
    from baddiffusion_b200.unet import UNet2DModel

    if arch == "cifar":
        name, S, B, trig, targ, lr, vt = "DDPM-CIFAR10-32", 32, 128, "BOX_14", "HAT", 2e-4, "fixed_large"
    else:
        name, S, B, trig, targ, lr, vt = "DDPM-CELEBA-HQ-256", 256, 4, "GLASSES", "CAT", 8e-5, "fixed_small"
    torch.manual_seed(0)
    model = UNet2DModel(**DiffuserModelSched.ARCH[name]).cuda()
    sched = DDPMScheduler(variance_type=vt, clip_sample=True)
    ds = SyntheticDataset(S, 3, poison_rate=0.1, seed=1000 * ctx.rank, trigger=trig, target=targ)
    tr = Trainer(model, sched, B, ds.trigger, ds.target, lr=lr, total_steps=50 * 469, warmup_steps=500,
                 process_group=ctx.pg, seed=1234 + ctx.rank)
    host = [ds.batch(B, index=i) for i in range(4)]
    tr.load_batch(host[0].image, host[0].is_poison)
    for _ in range(max(W, 3)):
        tr.t.copy_(torch.randint(0, 1000, (B,), device="cuda"))
        tr.step_resident(True)
    ctx.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        tr.t.copy_(torch.randint(0, 1000, (B,), device="cuda"))  # baddiffusion.py:600 (GPU RNG draw of t)
        tr.step_resident(True)
    e1.record()
    ctx.barrier()
    ms = ctx.max_over_ranks(e0.elapsed_time(e1))
    loss_resident = float(tr.loss)
    assert _lib.lib().bd_umma_error() == 0, "tcgen05 pipeline time-out"
    # end to end through the public API: pinned host batch -> H2D -> step -> loss D2H, every step
    K2 = e2e_steps or K
    for i in range(2):
        tr.step(host[i % 4].image, host[i % 4].is_poison)
        tr.loss_item()
    ctx.barrier()
    e0.record()
    last = 0.0
    for i in range(K2):
        hb = host[i % 4]
        tr.step(hb.image, hb.is_poison)   # H2D of the pinned batch (copy stream) -> forward/backward -> optimizer
        last = tr.loss_item()             # D2H of THIS step's loss, every step
    e1.record()
    ctx.barrier()
    ms_e2e = ctx.max_over_ranks(e0.elapsed_time(e1))
    out = {"arch": name, "model": model, "sched": sched, "per_gpu_batch": B, "ms_per_step": ms / K,
           "value": K * B * ctx.world / (ms / 1e3), "e2e_value": K2 * B * ctx.world / (ms_e2e / 1e3),
           "e2e_ms_per_step": ms_e2e / K2, "h2d": host[0].image.numel() * 4 + host[0].is_poison.numel(), "d2h": 4,
           "launches_per_step": int(tr.launches_per_step), "loss_resident": loss_resident, "loss_e2e": last,
           "loss_scale": tr.loss_scale, "allreduce_bytes": int(tr.gflat.numel()) * 4 if ctx.world > 1 else 0}
    del tr
    torch.cuda.empty_cache()
    return out


def measure_sampling(ctx, model, sched, ddim, SB=256, reps=1):
    """configs[4]-style sampling at N GPUs through the public drivers: every rank denoises SB clean-init + SB
    backdoor-init samples (DDIM-50) or SB samples (DDPM-1000) of its own, no collective; timed on the device on every
    rank, max over ranks.  `value`: init already resident in HBM, images left on the device (pipeline loop only);
    `e2e`: batch_sampling() -- host init -> H2D -> loop -> finalize -> D2H numpy images."""
    torch = ctx.torch
    from baddiffusion_b200 import ops
    from baddiffusion_b200.dataset import Backdoor
    from baddiffusion_b200.model import batch_sampling, shard_for_rank
    from baddiffusion_b200.pipelines import DDIMPipeline, DDPMPipeline

    pipe = (DDIMPipeline if ddim else DDPMPipeline)(unet=model, scheduler=sched)
    pipe.set_progress_bar_config(disable=True)
    nsteps = 50 if ddim else 1000
    trig = Backdoor(root="datasets").get_trigger(type="BOX_14", channel=3, image_size=32)
    N = SB * ctx.world
    noise = torch.randn(N, 3, 32, 32, generator=torch.Generator().manual_seed(0))
    lo, hi = shard_for_rank(N, ctx.rank, ctx.world)      # this rank's slice of the global sample set (model.py:478)
    inits = [noise[lo:hi].pin_memory()] + ([(noise[lo:hi] + trig[None]).pin_memory()] if ddim else [])
    run = lambda init: batch_sampling(hi - lo, lambda **kw: pipe(num_inference_steps=nsteps, **kw), init=init,
                                      max_batch_n=SB, rng=None)
    # warm-up: graph capture + one full chain (the captured graph is keyed on the length of the coefficient table)
    pipe(batch_size=hi - lo, num_inference_steps=nsteps, init=inits[0], output_type=None)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # (1) device-resident loop
    dev_inits = [i.cuda() for i in inits]
    before = ops.launch_count()
    ctx.barrier()
    e0.record()
    for _ in range(reps):
        for di in dev_inits:
            x = di.clone()
            pipe.scheduler.set_timesteps(nsteps)
            ts = [int(t) for t in pipe.scheduler.timesteps]
            table = pipe.scheduler.coef_table(0.0, False, ts) if ddim else pipe.scheduler.coef_table(ts)
            pipe._run_loop(x, ts, table, None, ddim, [False] * len(ts) if ddim else [t > 0 for t in ts], False, [])
    e1.record()
    ctx.barrier()
    launches = ops.launch_count() - before
    ms = ctx.max_over_ranks(e0.elapsed_time(e1))
    # (2) end to end
    ctx.barrier()
    t0 = time.perf_counter()
    e0.record()
    imgs = [run(i) for i in inits]
    e1.record()
    ctx.barrier()
    wall = time.perf_counter() - t0
    ms_e2e = ctx.max_over_ranks(e0.elapsed_time(e1))
    n_local = (hi - lo) * len(inits)
    total = n_local * ctx.world
    assert all(im.shape == (hi - lo, 32, 32, 3) and float(im.min()) >= 0.0 and float(im.max()) <= 1.0 for im in imgs)
    pk = peaks()
    steps_run = reps * len(inits) * nsteps
    return {"sampler": "DDIM-SCHED" if ddim else "DDPM-SCHED", "steps_per_sample": nsteps, "samples_per_gpu": n_local,
            "inits": ["clean", "backdoor (noise + trigger)"] if ddim else ["clean"], "batch": hi - lo,
            "value": reps * total / (ms / 1e3), "ms_per_denoise_step": ms / steps_run, "denoise_steps_timed": steps_run,
            "e2e_value": total / (ms_e2e / 1e3), "e2e_wall_s": wall, "h2d": n_local * 3 * 32 * 32 * 4, "d2h": n_local * 3 * 32 * 32 * 4,
            "launches": int(launches),
            "fwd_tensor_frac_of_sustained": (hi - lo) / (ms / steps_run / 1e3) * GFLOP_FWD["cifar"] / 1e3 / pk["tf_sust"]}


def roofline_conv(ctx, B=128):
    """FLOP-dominant kernel: persistent 3x3 conv 128->128 @32x32 (7 fwd + 7 dgrad-shaped launches per CIFAR step), timed
    alone with CUDA events after an L2 flush."""
    torch = ctx.torch
    from baddiffusion_b200 import _lib, ops

    pk = peaks()
    H, C = 32, 128
    x = torch.randn(B, H, H, C, device="cuda").half()
    w = (torch.randn(9, C, C, device="cuda") / 34).half()
    y = torch.empty(B, H, H, C, dtype=torch.half, device="cuda")
    bias = torch.zeros(C, device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(5):
        ops.conv_fwd(x, w, y, ksize=3, bias=bias, impl=_lib.BD_IMPL_UMMA)
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    reps, tot = 10, 0.0
    for _ in range(reps):
        flush.zero_()  # > L2 (126 MB): the next launch reads its operands from HBM
        e0.record()
        ops.conv_fwd(x, w, y, ksize=3, bias=bias, impl=_lib.BD_IMPL_UMMA)
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    k_ms = tot / reps
    flops = 2.0 * B * H * H * C * C * 9
    achieved = flops / (k_ms / 1e3) / 1e12
    return {"bound": "tensor", "kernel": "umma_conv3p_kernel (persistent 3x3 conv, halo reuse) 128->128 @32x32, B=128",
            "achieved": achieved, "peak": pk["tf_burst"], "unit": "TFLOP/s", "frac": achieved / pk["tf_burst"],
            "traffic": ncu_traffic(), "peak_source": f"{pk['src']} (burst, kernel timed alone)", "ms_per_launch": k_ms,
            "algorithmic_flops_per_launch": flops,
            "algorithmic_bytes_per_launch": 2.0 * (2 * B * H * H * C + 9 * C * C),
            "conv_hbm_gbs_algorithmic": 2.0 * (2 * B * H * H * C + 9 * C * C) / (k_ms / 1e3) / 1e9}


def roofline_groupnorm(ctx, B=128):
    """Time-dominant kernel CLASS of the step (GroupNorm fwd/bwd, HBM-bound by construction): the fused backward at
    32x32x128 -- reads x, dy, add_dx and writes dx (fp16): 8 B/element -- timed alone after an L2 flush."""
    torch = ctx.torch
    from baddiffusion_b200 import ops

    pk = peaks()
    H, C, G = 32, 128, 32
    x = torch.randn(B, H, H, C, device="cuda").half()
    dy = torch.randn(B, H, H, C, device="cuda").half()
    add = torch.randn(B, H, H, C, device="cuda").half()
    dx = torch.empty_like(x)
    y = torch.empty_like(x)
    gamma, beta = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
    stats = torch.empty(B, G, 2, device="cuda")
    work = torch.empty(ops.gn_workspace_floats(B, C), device="cuda")
    parts = torch.empty(B, 2 * C, device="cuda")
    gsum = torch.empty(B, C, device="cuda")
    dg, db = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    res = {}
    for which in ("fwd", "bwd"):
        tot, reps = 0.0, 10
        for i in range(reps + 3):
            flush.zero_()
            e0.record()
            if which == "fwd":
                ops.groupnorm_fwd(x, y, gamma, beta, stats, work, G, 1e-6, True)
            else:
                ops.groupnorm_bwd(x, dy, dx, gamma, beta, stats, dg, db, work, G, True, add_dx=add, gsum=gsum, parts=parts)
            e1.record()
            torch.cuda.synchronize()
            if i >= 3:
                tot += e0.elapsed_time(e1)
        res[which] = tot / reps
    n = B * H * H * C
    bwd_bytes, fwd_bytes = 8.0 * n, 4.0 * n
    ach = bwd_bytes / (res["bwd"] / 1e3) / 1e9
    return {"bound": "hbm", "kernel": "GroupNorm+SiLU backward (gn_bwd_* kernels) 32x32x128, B=128, with add_dx", "achieved": ach,
            "peak": pk["hbm"], "unit": "GB/s", "frac": ach / pk["hbm"], "traffic": ncu_traffic("roofline_gn_bwd_traffic.json"),
            "peak_source": f"{pk['src']} (copy bandwidth, kernel timed alone)", "ms_per_launch": res["bwd"],
            "algorithmic_bytes_per_launch": bwd_bytes,
            "forward": {"ms_per_launch": res["fwd"], "algorithmic_bytes_per_launch": fwd_bytes,
                        "achieved": fwd_bytes / (res["fwd"] / 1e3) / 1e9, "frac": fwd_bytes / (res["fwd"] / 1e3) / 1e9 / pk["hbm"]}}


def gpu_incumbent(ctx, B=128, steps=10, warmup=3):
    """Library incumbent on the same GPU (SURVEY 8d / BASELINE.md 3 'Also timed'): the reference's modules as restated by
    the oracle, run by PyTorch eager on CUDA -- cuDNN convolutions / cuBLAS GEMMs under fp16 autocast with channels_last
    activations and weights, GradScaler, clip_grad_norm_, torch.optim.Adam -- same B = 128 poisoned train step."""
    torch = ctx.torch
    import numpy as np

    from oracle import torch_ref as O

    cfg = O.CIFAR10_CONFIG
    dev = torch.device("cuda")
    sd = {}
    for k, v in O.make_state_dict(cfg, 0).items():
        v = v.to(dev)
        if v.dim() == 4:
            v = v.contiguous(memory_format=torch.channels_last)
        sd[k] = v.requires_grad_(True)
    params = list(sd.values())
    opt = torch.optim.Adam(params, lr=2e-4)
    scaler = torch.amp.GradScaler("cuda")
    _, alphas, acp = (t.to(dev) for t in O.beta_tables())
    z = np.load(os.path.join(ROOT, "baddiffusion_b200", "assets", "backdoor_assets.npz"))
    trig, targ = O.get_trigger("BOX_14", 32).to(dev), torch.from_numpy(z["target_HAT_32"]).to(dev)
    image = torch.randn(B, 3, 32, 32, device=dev).clamp(-1, 1)
    isp = torch.tensor([j % 10 == 0 for j in range(B)], device=dev)
    torch.backends.cudnn.benchmark = True

    def step():
        R, x0 = O.poison_blend(image, isp, trig, targ)
        noise = torch.randn(image.shape, device=dev)
        t = torch.randint(0, 1000, (B,), device=dev)
        with torch.autocast("cuda", dtype=torch.float16):
            x_noisy, target = O.q_sample(alphas, acp, x0, R, t, noise)
            eps = O.unet_forward(sd, cfg, x_noisy.contiguous(memory_format=torch.channels_last), t)
            loss = torch.nn.functional.mse_loss(target, eps.float())
        scaler.scale(loss).backward()
        scaler.unscale_(opt)
        torch.nn.utils.clip_grad_norm_(params, 1.0)
        scaler.step(opt)
        scaler.update()
        opt.zero_grad(set_to_none=True)
        return loss

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    # forward-only (sampling incumbent), B = 256
    x = torch.randn(256, 3, 32, 32, device=dev).contiguous(memory_format=torch.channels_last)
    tt = torch.full((256,), 500, device=dev)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        for _ in range(3):
            O.unet_forward(sd, cfg, x, tt)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            O.unet_forward(sd, cfg, x, tt)
        e1.record()
        torch.cuda.synchronize()
    fwd_ms = e0.elapsed_time(e1) / 10
    out = {"what": "reference modules (oracle restatement) in PyTorch-eager CUDA: cuDNN/cuBLAS, fp16 autocast, channels_last, "
                   "GradScaler + clip + torch.optim.Adam; B=128 CIFAR10-32 poisoned train step",
           "train_images_per_sec": B / (ms / 1e3), "ms_per_step": ms, "loss": float(loss),
           "unet_forward_b256_ms": fwd_ms, "ddim_50_step_samples_per_sec": 256 / (fwd_ms * 50 / 1e3),
           "ddpm_1000_step_samples_per_sec": 256 / (fwd_ms * 1000 / 1e3), "torch": torch.__version__,
           "cudnn": torch.backends.cudnn.version()}
    del sd, params, opt
    torch.cuda.empty_cache()
    return out


def run_ours(args):
    ctx = Ctx()
    torch = ctx.torch
    from baddiffusion_b200 import _lib

    _lib.lib()
    wl, K, W = args.workload, args.steps, args.warmup
    pk = peaks()
    sampler = ClockSampler(ctx.local) if ctx.rank == 0 else None
    extras = {}

    def sub_train(arch, steps, warm):
        r = measure_train(ctx, arch, steps, warm, e2e_steps=steps)
        model, sched = r.pop("model"), r.pop("sched")
        r["step_tensor_frac_of_sustained"] = r["value"] / ctx.world * 3 * GFLOP_FWD[arch] / 1e3 / pk["tf_sust"]
        return r, model, sched

    def guarded(name, fn):
        try:
            return fn()
        except Exception as e:  # sub-records are reported, never fatal for the main metric
            log(f"[bench] sub-record {name} failed: {e!r}")
            return {"error": repr(e)}

    if wl in ("cifar_train", "celeba_train"):
        arch = "cifar" if wl == "cifar_train" else "celeba"
        main, model, sched = sub_train(arch, K, W)
        clocks = sampler.stop() if sampler else None
        value, ms_per_step, e2e = main["value"], main["ms_per_step"], {
            "value": main["e2e_value"], "unit": UNITS[wl], "h2d_bytes_per_step": main["h2d"], "d2h_bytes_per_step": main["d2h"],
            "ms_per_step": main["e2e_ms_per_step"],
            "note": "separate timed loop: Trainer.step(pinned host batch) + Trainer.loss_item() every step; the H2D copy runs on a "
                    "copy stream ordered after the last reader of the input buffers, so it overlaps the previous step's optimizer"}
        launches = main["launches_per_step"] * K
        steps_reported = K
        if wl == "cifar_train" and not args.no_extras:
            from baddiffusion_b200.schedulers import DDPMScheduler

            extras["sampling"] = {
                "ddim_50": guarded("ddim", lambda: measure_sampling(ctx, model, sched, True)),
                "ddpm_1000": guarded("ddpm", lambda: measure_sampling(ctx, model, DDPMScheduler(variance_type="fixed_large", clip_sample=True), False)),
            }
            del model
            torch.cuda.empty_cache()
            if not args.no_celeba:
                def celeba():
                    r, m2, _ = sub_train("celeba", 10, 3)
                    del m2
                    r["workload"] = WORKLOAD_TEXT["celeba_train"]
                    return r
                extras["celebahq_256"] = guarded("celeba", celeba)
                torch.cuda.empty_cache()
    else:
        from baddiffusion_b200.model import DiffuserModelSched
        from baddiffusion_b200.schedulers import DDPMScheduler
        from baddiffusion_b200.unet import UNet2DModel

        torch.manual_seed(0)
        model = UNet2DModel(**DiffuserModelSched.ARCH["DDPM-CIFAR10-32"]).cuda()
        sched = DDPMScheduler(variance_type="fixed_large", clip_sample=True)
        main = measure_sampling(ctx, model, sched, wl == "ddim_sample", reps=max(1, K if wl == "ddim_sample" else 1))
        clocks = sampler.stop() if sampler else None
        value, ms_per_step = main["value"], main["ms_per_denoise_step"]
        e2e = {"value": main["e2e_value"], "unit": UNITS[wl], "h2d_bytes_per_step": main["h2d"], "d2h_bytes_per_step": main["d2h"],
               "note": "a step of the e2e leg = one batch_sampling() call per init: H2D init, all denoise steps, D2H images"}
        launches = main["launches"]
        steps_reported = main["denoise_steps_timed"]

    line = None
    if ctx.rank == 0:
        roof = roof_gn = cpu = inc = None
        if wl in ("cifar_train", "celeba_train"):
            roof = guarded("roofline", lambda: roofline_conv(ctx))
            roof["step_tensor_frac_of_sustained"] = main["step_tensor_frac_of_sustained"]
            roof_gn = guarded("roofline_gn", lambda: roofline_groupnorm(ctx))
        else:
            roof = {"bound": "tensor", "kernel": "whole denoise step (UNet forward + fused scheduler step), CIFAR10-32 UNet",
                    "achieved": main["fwd_tensor_frac_of_sustained"] * pk["tf_sust"], "peak": pk["tf_sust"], "unit": "TFLOP/s",
                    "frac": main["fwd_tensor_frac_of_sustained"], "traffic": None,
                    "peak_source": f"{pk['src']} (sustained, timed inside the loop)"}
        if ctx.world == 1 and not args.no_cpu_baseline:
            if wl == "cifar_train":
                times, threads = cpu_train_steps(32, 4, 1, budget_s=40.0)
                cpu = {"value": 32 * len(times) / sum(times), "unit": UNITS[wl], "cores": threads, "kind": "port",
                       "sample": f"{len(times)} full train steps at batch 32 (of the 128-image batch), fp32 torch CPU", **_stats(times)}
            elif wl == "celeba_train":
                times, threads = cpu_train_steps(1, 2, 1, budget_s=60.0, arch="celeba")
                cpu = {"value": len(times) / sum(times), "unit": UNITS[wl], "cores": threads, "kind": "port",
                       "sample": f"{len(times)} full train steps at batch 1 (of the 4-image batch), fp32 torch CPU", **_stats(times)}
            else:
                ddim = wl == "ddim_sample"
                times, threads = cpu_sample_steps(32, 4, ddim)
                per = sum(times) / len(times)
                cpu = {"value": 32 / (per * (50 if ddim else 1000)), "unit": UNITS[wl], "cores": threads, "kind": "port",
                       "sample": f"{len(times)} denoise steps at batch 32, fp32 torch CPU, scaled to {50 if ddim else 1000} steps",
                       **_stats(times)}
            if wl == "cifar_train" and not args.no_extras:
                # sampling CPU baseline: a few denoise steps at B = 32
                ts, th = cpu_sample_steps(32, 3, True)
                per = sum(ts) / len(ts)
                cpu["sampling"] = {"s_per_denoise_step_b32": per, "ddim_50_step_samples_per_sec": 32 / (per * 50),
                                   "ddpm_1000_step_samples_per_sec": 32 / (per * 1000), "cores": th, **_stats(ts)}
        if ctx.world == 1 and wl == "cifar_train" and not args.no_incumbent:
            inc = guarded("gpu_incumbent", lambda: gpu_incumbent(ctx))
        cfgd = {"workload": WORKLOAD_TEXT[wl], "per_gpu_batch": main.get("per_gpu_batch", main.get("batch")),
                "precision": "fp16 operands, fp32 accumulate/master/GroupNorm/softmax, loss scaling",
                "l2": "working set (activations+grads, >2 GB/step) exceeds the 126 MB L2; no explicit flush",
                "parallelism": f"dp{ctx.world}" if ctx.world > 1 else "single"}
        if wl in ("cifar_train", "celeba_train"):
            cfgd.update({"global_batch": main["per_gpu_batch"] * ctx.world, "poison_rate": 0.1,
                         "trigger": "BOX_14" if wl == "cifar_train" else "GLASSES", "target": "HAT" if wl == "cifar_train" else "CAT"})
        line = {
            "metric": METRICS[wl], "value": value, "unit": UNITS[wl], "n_gpus": ctx.world, "steps": steps_reported, "warmup": W,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16",
            "data": "synthetic", "config": cfgd, "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
            "roofline": roof, "roofline_time_dominant": roof_gn, "cpu_baseline": cpu, "gpu_incumbent": inc,
        }
        if wl in ("cifar_train", "celeba_train"):
            line["launches_per_step"] = main["launches_per_step"]
            line["loss"] = {"resident_last": main["loss_resident"], "e2e_last": main["loss_e2e"], "loss_scale": main["loss_scale"]}
            if ctx.world > 1:
                line["allreduce_bytes_per_step"] = main["allreduce_bytes"]
        else:
            line["sampling_detail"] = main
        line.update(extras)
        print(json.dumps(line), flush=True)
    if ctx.world > 1:
        ctx.dist.barrier()
        ctx.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cifar_train", choices=list(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-incumbent", action="store_true", help="skip the PyTorch-eager CUDA incumbent leg")
    ap.add_argument("--no-celeba", action="store_true", help="skip the CelebA-HQ-256 sub-record of the default line")
    ap.add_argument("--no-extras", action="store_true", help="default line only: no sampling / CelebA sub-records")
    ap.add_argument("--ref-batch", type=int, default=0, help="reference arm: override the per-step batch (tests)")
    ap.add_argument("--ref-budget-s", type=float, default=150.0, help="reference arm: CPU seconds after which it stops stepping")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
