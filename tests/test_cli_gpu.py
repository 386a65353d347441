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
    _run(["--project", "t", "--mode", "train", "--dataset", "CIFAR10", "--batch", "16", "--epoch", "1", "--poison_rate", "0.1",
          "--trigger", "BOX_14", "--target", "HAT", "--ckpt", ck, "--fclip", "o", "-o", "--result", res,
          "--dataset_size", "64", "--max_steps", "3", "--save_image_epochs", "100", "--save_model_epochs", "1"])
    out = [d for d in os.listdir(res) if d.startswith("res_")]
    assert len(out) == 1
    out = os.path.join(res, out[0])
    for f in ("args.json", "config.json", "model_index.json", "unet/config.json", "unet/diffusion_pytorch_model.bin",
              "scheduler/scheduler_config.json", "data.ckpt", "ckpt/optimizer.bin", "samples", "backdoor_samples"):
        assert os.path.exists(os.path.join(out, f)), f
    _run(["--mode", "resume", "--ckpt", out, "--max_steps", "5"])
    _run(["--mode", "sampling", "--ckpt", out, "--fclip", "w", "--sched", "DDIM-SCHED"])
    assert any(n.startswith("final") for n in os.listdir(os.path.join(out, "samples")))
    env = {"BD_MEASURE_N": "8"}
    r = _run(["--mode", "measure", "--ckpt", out, "--fclip", "o", "--eval_max_batch", "8", "--sched", "DDIM-SCHED"], env)
    score = json.load(open(os.path.join(out, "score.json")))
    assert score["MSE"] is not None and score["MSE"] >= 0.0
    assert len(os.listdir(os.path.join(out, "measure", "clean_noclip"))) == 8
