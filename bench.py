#!/usr/bin/env python
"""Benchmark of the BadDiffusion hot path (BASELINE.json): poisoned DDPM training step, images/sec.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host CPU cores

Workload (config.workload): BASELINE.json configs[1] -- DDPM-CIFAR10-32 UNet2DModel (35.7 M params), batch 128 per
GPU, poison_rate 0.1, trigger BOX_14 -> target HAT, synthetic 3x32x32 data, random-init weights.  A "step" is one
full training step: batch-prep -> UNet fwd -> MSE -> UNet bwd -> (all-reduce) -> clip + Adam.  Weak scaling: the
per-GPU batch stays 128 (N=8 is BASELINE configs[2], global batch 1024).

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

METRIC = "train_images_per_sec"
UNIT = "images/s"
FWD_GFLOP_PER_IMG = 12.444      # CIFAR10-32 UNet forward, 2*MAC
TRAIN_GFLOP_PER_IMG = 37.33     # 3 x forward


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the roofline kernel, from the committed
    `ncu --set full` capture (profiles/roofline_kernel_traffic.json; written by scripts/ncu_traffic.py), or None."""
    p = os.path.join(ROOT, "profiles", "roofline_kernel_traffic.json")
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
# baddiffusion.py:593-615 train step) on the host cores.  /root/reference is not on the GPU box and the reference is
# a Python repo, so kind = "port".
# ----------------------------------------------------------------------------------------------------------
def cpu_train_steps(batch, steps, warmup, threads=None, budget_s=None):
    import torch

    from oracle import torch_ref as O

    threads = threads or len(os.sched_getaffinity(0))
    torch.set_num_threads(threads)
    cfg = O.CIFAR10_CONFIG
    sd = {k: v.clone().requires_grad_(True) for k, v in O.make_state_dict(cfg, 0).items()}
    params = list(sd.values())
    opt = torch.optim.Adam(params, lr=2e-4)
    _, alphas, acp = O.beta_tables()
    trig = O.get_trigger("BOX_14", 32)
    bt_path = os.path.join(ROOT, "baddiffusion_b200", "assets", "backdoor_assets.npz")
    import numpy as np

    targ = torch.from_numpy(np.load(bt_path)["target_HAT_32"])
    g = torch.Generator().manual_seed(0)
    times = []
    t_begin = time.perf_counter()
    for i in range(warmup + steps):
        image = torch.randn(batch, 3, 32, 32, generator=g).clamp(-1, 1)
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


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample_b = 8
    times, threads = cpu_train_steps(sample_b, args.steps, args.warmup)
    total = sum(times)
    value = sample_b * len(times) / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(times),
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "DDPM-CIFAR10-32 poisoned train step (BOX_14->HAT, poison_rate 0.1), synthetic 3x32x32",
                   "per_gpu_batch": 128, "sample": f"{sample_b} images per step of the 128-image batch"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{len(times)} full train steps (fwd+bwd+clip+Adam) at batch {sample_b}, fp32, torch CPU"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    pg = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        pg = dist.group.WORLD
    from baddiffusion_b200 import _lib, ops
    from baddiffusion_b200.dataset import SyntheticDataset
    from baddiffusion_b200.model import DiffuserModelSched
    from baddiffusion_b200.pipelines import DDPMPipeline
    from baddiffusion_b200.schedulers import DDPMScheduler
    from baddiffusion_b200.train import Trainer
    from baddiffusion_b200.unet import UNet2DModel

    _lib.lib()
    B, K, W = args.batch, args.steps, args.warmup
    torch.manual_seed(0)
    model = UNet2DModel(**DiffuserModelSched.ARCH["DDPM-CIFAR10-32"]).cuda()
    sched = DDPMScheduler(variance_type="fixed_large", clip_sample=True)
    ds = SyntheticDataset(32, 3, poison_rate=0.1, seed=1000 * rank)
    tr = Trainer(model, sched, B, ds.trigger, ds.target, lr=2e-4, total_steps=50 * 469, warmup_steps=500,
                 process_group=pg, seed=1234 + rank)
    host = [ds.batch(B, index=i) for i in range(4)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    # ---- (1) device-resident throughput: inputs already in HBM when the timed region starts
    tr.load_batch(host[0].image, host[0].is_poison)
    # clocks / throttle reasons are sampled from the warm-up to the end of the second timed region (a 0.2 s timed loop
    # alone is shorter than nvidia-smi's polling latency); every sample is taken under the same load
    sampler = ClockSampler(local) if rank == 0 else None
    for _ in range(max(W, 3)):
        tr.t.copy_(torch.randint(0, 1000, (B,), device="cuda"))
        tr.step_resident(True)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        tr.t.copy_(torch.randint(0, 1000, (B,), device="cuda"))  # baddiffusion.py:600 (GPU RNG draw of t)
        tr.step_resident(True)
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    value = K * B * world / (ms / 1e3)
    loss_resident = float(tr.loss)
    assert _lib.lib().bd_umma_error() == 0, "tcgen05 pipeline time-out"

    # ---- (2) end to end through the public API: pinned host batch -> H2D -> step -> loss D2H, every step
    for i in range(2):
        tr.step(host[i % 4].image, host[i % 4].is_poison).item()
    barrier()
    e0.record()
    last = 0.0
    for i in range(K):
        hb = host[i % 4]
        last = tr.step(hb.image, hb.is_poison).item()
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if sampler else None
    e2e_value = K * B * world / (ms_e2e / 1e3)
    h2d = host[0].image.numel() * 4 + host[0].is_poison.numel()

    line = None
    if rank == 0:
        pk = peaks()
        # ---- (3) roofline of the dominant kernel: 3x3 conv 128->128 @32x32 (7 fwd + 7 dgrad-shaped launches/step)
        H, C = 32, 128
        x = torch.randn(B, H, H, C, device="cuda").half()
        w = (torch.randn(9, C, C, device="cuda") / 34).half()
        y = torch.empty(B, H, H, C, dtype=torch.half, device="cuda")
        bias = torch.zeros(C, device="cuda")
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
        roof = {"bound": "tensor", "kernel": "umma_conv3p_kernel (persistent 3x3 conv, halo reuse) 128->128 @32x32, B=128",
                "achieved": achieved, "peak": pk["tf_burst"], "unit": "TFLOP/s", "frac": achieved / pk["tf_burst"],
                "traffic": ncu_traffic(), "peak_source": f"{pk['src']} (burst, kernel timed alone)", "ms_per_launch": k_ms,
                "algorithmic_flops_per_launch": flops,
                "algorithmic_bytes_per_launch": 2.0 * (2 * B * H * H * C + 9 * C * C),
                "conv_hbm_gbs_algorithmic": 2.0 * (2 * B * H * H * C + 9 * C * C) / (k_ms / 1e3) / 1e9,
                "step_tensor_frac_of_sustained": value / world * TRAIN_GFLOP_PER_IMG / 1e3 / pk["tf_sust"]}
        del x, w, y, flush

        # ---- (4) DDPM sampling (second half of the BASELINE metric): per-step time of the captured sampling graph
        samp = None
        try:
            pipe = DDPMPipeline(unet=model, scheduler=sched)
            pipe.set_progress_bar_config(disable=True)
            SB, nst = 256, 20
            init = torch.randn(SB, 3, 32, 32)
            pipe(batch_size=SB, num_inference_steps=nst, init=init, output_type=None)  # capture + warm-up
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            pipe(batch_size=SB, num_inference_steps=nst, init=init, output_type=None)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            samp = {"batch": SB, "measured_steps": nst, "ms_per_denoise_step": 1e3 * dt / nst,
                    "ddpm_1000_step_samples_per_sec": SB / (dt / nst * 1000),
                    "ddim_50_step_samples_per_sec": SB / (dt / nst * 50),
                    "fwd_tensor_frac_of_sustained": SB / (dt / nst) * FWD_GFLOP_PER_IMG / 1e3 / pk["tf_sust"]}
        except Exception as e:  # sampling is reported, never fatal for the train metric
            samp = {"error": repr(e)}

        # ---- (4b) BASELINE configs[3] shape: DDPM-CELEBA-HQ-256 UNet (113.7 M params), 256x256, per-GPU batch 4
        celeba = None
        launches_per_step, loss_scale = int(tr.launches_per_step), tr.loss_scale
        if world == 1 and not args.no_celeba:
            try:
                tr = None
                torch.cuda.empty_cache()
                cm = UNet2DModel(**DiffuserModelSched.ARCH["DDPM-CELEBA-HQ-256"]).cuda()
                cds = SyntheticDataset(256, 3, poison_rate=0.1, seed=7)
                CB = 4
                ctr = Trainer(cm, DDPMScheduler(variance_type="fixed_small", clip_sample=True), CB, cds.trigger, cds.target,
                              lr=8e-5, total_steps=1000, warmup_steps=10)
                hb = cds.batch(CB, index=0)
                ctr.load_batch(hb.image, hb.is_poison)
                for _ in range(3):
                    ctr.step_resident(True)
                torch.cuda.synchronize()
                e0.record()
                for _ in range(10):
                    ctr.step_resident(True)
                e1.record()
                torch.cuda.synchronize()
                cms = e0.elapsed_time(e1) / 10
                celeba = {"workload": "DDPM-CELEBA-HQ-256 poisoned train step, synthetic 3x256x256, per-GPU batch 4",
                          "ms_per_step": cms, "train_images_per_sec": CB / (cms / 1e3),
                          "step_tensor_frac_of_sustained": CB / (cms / 1e3) * 1491.1 / 1e3 / pk["tf_sust"],
                          "loss": float(ctr.loss)}
                del ctr, cm
                torch.cuda.empty_cache()
            except Exception as e:
                celeba = {"error": repr(e)}

        # ---- (5) CPU baseline: the oracle port on this box's host cores, bounded sample
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            times, threads = cpu_train_steps(8, 3, 1, budget_s=45.0)
            cpu = {"value": 8 * len(times) / sum(times), "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": f"{len(times)} full train steps at batch 8 (of the 128-image batch), fp32 torch CPU"}

        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16",
            "data": "synthetic",
            "config": {"workload": "DDPM-CIFAR10-32 poisoned train step (BASELINE configs[1]; N>1: configs[2]): p_losses_diffuser "
                                   "fwd/bwd over the google/ddpm-cifar10-32 UNet2DModel topology (35.7 M parameters, random init) "
                                   "+ clip + Adam, synthetic 3x32x32",
                       "per_gpu_batch": B, "global_batch": B * world, "poison_rate": 0.1, "trigger": "BOX_14",
                       "target": "HAT", "precision": "fp16 operands, fp32 accumulate/master/GroupNorm/softmax, loss scaling",
                       "l2": "working set (activations+grads, >2 GB/step) exceeds the 126 MB L2; no explicit flush",
                       "parallelism": f"dp{world}" if world > 1 else "single"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / K},
            "gpu_launches": launches_per_step * K,
            "launches_per_step": launches_per_step,
            "roofline": roof, "cpu_baseline": cpu, "sampling": samp, "celebahq_256": celeba,
            "loss": {"resident_last": loss_resident, "e2e_last": last, "loss_scale": loss_scale},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=128, help="per-GPU batch")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-celeba", action="store_true", help="skip the CelebA-HQ-256 shaped extra measurement")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
