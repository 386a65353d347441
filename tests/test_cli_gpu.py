"""End-to-end CLI run on the GPU: train a few steps from a synthetic checkpoint, checkpoint layout, resume, sampling,
measure (the modes of baddiffusion.py:16-20)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ, **(env or {}))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "baddiffusion.py")] + args, capture_output=True, text=True, env=e,
                       cwd=ROOT, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    return r


def test_train_resume_sample_measure(tmp_path):
    sys.path.insert(0, ROOT)
    from baddiffusion_b200.model import DiffuserModelSched

    ck = str(tmp_path / "ckpts" / "DDPM-CIFAR10-32")
    DiffuserModelSched.new_synthetic_checkpoint("DDPM-CIFAR10-32", ck, seed=0)
    res = str(tmp_path / "res")
    # batch 64 of the effective 128: two micro-batches per optimizer step (baddiffusion.py:195-217); 2 epochs x 2 batches
    _run(["--project", "t", "--mode", "train", "--dataset", "CIFAR10", "--batch", "64", "--epoch", "2", "--poison_rate", "0.1",
          "--trigger", "BOX_14", "--target", "HAT", "--ckpt", ck, "--fclip", "o", "-o", "--result", res,
          "--dataset_size", "128", "--save_image_epochs", "100", "--save_model_epochs", "1"])
    out = [d for d in os.listdir(res) if d.startswith("res_")]
    assert len(out) == 1
    out = os.path.join(res, out[0])
    for f in ("args.json", "config.json", "model_index.json", "unet/config.json", "unet/diffusion_pytorch_model.bin",
              "scheduler/scheduler_config.json", "data.ckpt", "ckpt/optimizer.bin", "samples", "backdoor_samples"):
        assert os.path.exists(os.path.join(out, f)), f
    import torch

    st = torch.load(os.path.join(out, "data.ckpt"))
    assert st == {"epoch": 1, "step": 4}, st       # the LAST epoch, not the start epoch (baddiffusion.py:643)
    opt = torch.load(os.path.join(out, "ckpt", "optimizer.bin"))
    assert int(opt["step"]) + int(opt["state"][4]) == 2   # 4 micro-batches = 2 optimizer steps (taken or skipped by the scaler)
    assert any(n.endswith("_sample_t0.png") for n in os.listdir(os.path.join(out, "samples")))
    r = _run(["--mode", "resume", "--ckpt", out, "--max_steps", "6"])
    st = torch.load(os.path.join(out, "data.ckpt"))
    assert st["epoch"] == 1 and st["step"] == 6, st   # quirk Q12: the saved epoch is re-run, from the saved step count
    _run(["--mode", "sampling", "--ckpt", out, "--fclip", "w", "--sched", "UNIPC-SCHED"])    # model.py:616-618: PNDMPipeline
    _run(["--mode", "sampling", "--ckpt", out, "--fclip", "w", "--sched", "DDIM-SCHED"])
    assert any(n.startswith("final") for n in os.listdir(os.path.join(out, "samples")))
    env = {"BD_MEASURE_N": "8"}
    r = _run(["--mode", "measure", "--ckpt", out, "--fclip", "o", "--eval_max_batch", "8", "--sched", "DDIM-SCHED"], env)
    score = json.load(open(os.path.join(out, "score.json")))
    assert score["MSE"] is not None and score["MSE"] >= 0.0 and -1.0 <= score["SSIM"] <= 1.0
    # the scores are those of the PNG files on disk (baddiffusion.py:539-546 re-reads them): oracle on the files
    import numpy as np
    from PIL import Image
    from baddiffusion_b200.dataset import Backdoor
    from oracle import torch_ref as O

    d = os.path.join(out, "measure", "backdoor_noclip")
    u8 = np.stack([np.asarray(Image.open(os.path.join(d, f"{i}.png")).convert("RGB")) for i in range(8)])
    bd = Backdoor(root="datasets")
    targ = bd.get_target(type="HAT", trigger=bd.get_trigger(type="BOX_14", channel=3, image_size=32))
    mse_ref, ssim_ref = O.backdoor_metrics(u8, targ)
    assert abs(score["MSE"] - mse_ref) <= 1e-6 * max(1.0, mse_ref) and abs(score["SSIM"] - ssim_ref) <= 1e-5, (score, mse_ref, ssim_ref)
    assert len(os.listdir(os.path.join(out, "measure", "clean_noclip"))) == 8
