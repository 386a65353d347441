"""CLI surface of baddiffusion.py (CPU): flags, per-mode allow-lists, result naming, JSON artefacts, poison split."""
import json
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import baddiffusion as cli  # noqa: E402


def test_flags_and_naming(tmp_path):
    a = cli.parse_args(["--project", "default", "--mode", "train", "--dataset", "CIFAR10", "--batch", "128", "--epoch", "50",
                        "--poison_rate", "0.1", "--trigger", "BOX_14", "--target", "HAT", "--ckpt", "DDPM-CIFAR10-32",
                        "--fclip", "o", "-o", "--gpu", "0", "--result", str(tmp_path)])
    cfg = cli.setup(a)
    assert cfg.output_dir.endswith("res_DDPM-CIFAR10-32_CIFAR10_ep50_c1.0_p0.1_BOX_14-HAT")  # baddiffusion.py:130-134
    assert cfg.learning_rate == 2e-4 and cfg.lr_warmup_steps == 500 and cfg.measure_sample_n == 2048
    saved = json.load(open(os.path.join(cfg.output_dir, "args.json")))
    assert saved["trigger"] == "BOX_14" and saved["mode"] == "train"
    assert os.path.isfile(os.path.join(cfg.output_dir, "config.json"))
    # sampling mode reloads the run's args.json and rejects train-only flags (baddiffusion.py:163-175)
    b = cli.parse_args(["--mode", "sampling", "--ckpt", cfg.output_dir, "--fclip", "w", "--eval_max_batch", "64"])
    c2 = cli.setup(b)
    assert c2.dataset == "CIFAR10" and c2.fclip == "w" and c2.trigger == "BOX_14"
    with pytest.raises(NotImplementedError):
        cli.setup(cli.parse_args(["--mode", "sampling", "--ckpt", cfg.output_dir, "--epoch", "3"]))
    with pytest.raises(NotImplementedError):
        cli.setup(cli.parse_args(["--mode", "train", "--dataset", "CIFAR10", "--ckpt", "x", "--sample_ep", "3"]))
    with pytest.raises(SystemExit):
        cli.parse_args(["--mode", "bogus"])


def test_celeba_defaults(tmp_path):
    a = cli.parse_args(["--mode", "train", "--dataset", "CELEBA-HQ", "--ckpt", "DDPM-CELEBA-HQ-256", "--trigger", "GLASSES",
                        "--target", "CAT", "--result", str(tmp_path), "--batch", "16"])
    cfg = cli.setup(a)
    assert cfg.learning_rate == 8e-5 and cli.DATASETS[cfg.dataset] == (256, 3)
    assert cfg.gradient_accumulation_steps == 4   # batch_256 = 64 (baddiffusion.py:108,205,217)


def test_batch_accumulation_and_lr_rules(tmp_path):
    """baddiffusion.py:194-223: effective batch 128 (32x32) / 64 (CELEBA, CELEBA-HQ); --batch must divide it; CELEBA is
    in the 256 group for the default LR; an existing output directory needs --overwrite."""
    base = ["--mode", "train", "--ckpt", "x", "--result", str(tmp_path)]
    cfg = cli.setup(cli.parse_args(base + ["--dataset", "CIFAR10", "--batch", "32"]))
    assert cfg.gradient_accumulation_steps == 4 and cfg.learning_rate == 2e-4
    with pytest.raises(ValueError):   # the default batch (512) is rejected like in the reference
        cli.setup(cli.parse_args(base + ["--dataset", "CIFAR10", "-o"]))
    with pytest.raises(ValueError):
        cli.setup(cli.parse_args(base + ["--dataset", "CIFAR10", "--batch", "48", "-o"]))
    with pytest.raises(ValueError):
        cli.setup(cli.parse_args(base + ["--dataset", "CELEBA-HQ", "--batch", "128", "-o"]))
    cfg = cli.setup(cli.parse_args(base + ["--dataset", "CELEBA", "--batch", "64"]))
    assert cfg.learning_rate == 8e-5 and cfg.gradient_accumulation_steps == 1
    with pytest.raises(ValueError, match="overwrite"):
        cli.setup(cli.parse_args(base + ["--dataset", "CELEBA", "--batch", "64"]))
    cli.setup(cli.parse_args(base + ["--dataset", "CELEBA", "--batch", "64", "-o"]))


def test_poison_split_and_rank_shards(tmp_path):
    a = cli.parse_args(["--mode", "train", "--dataset", "CIFAR10", "--ckpt", "x", "--batch", "8", "--poison_rate", "0.1",
                        "--trigger", "BOX_14", "--target", "CORNER", "--result", str(tmp_path), "--dataset_size", "160"])
    cfg = cli.setup(a)
    os.environ["WORLD_SIZE"] = "2"
    try:
        d = cli.Data(cfg)
    finally:
        os.environ.pop("WORLD_SIZE")
    assert int(d.poison.sum()) == 16 and d.num_batch == 10
    b0 = list(d.epoch_batches(0, 0, 2))
    b1 = list(d.epoch_batches(0, 1, 2))
    assert len(b0) == len(b1) == 10 and b0[0].image.shape == (8, 3, 32, 32) and b0[0].is_poison.dtype == torch.uint8
    assert not torch.equal(b0[0].image, b1[0].image)
    assert d.trigger.shape == (3, 32, 32) and float(d.target.min()) == pytest.approx(-0.4)
