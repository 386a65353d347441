#!/usr/bin/env python
"""BadDiffusion command line (drop-in for the reference's baddiffusion.py on the B200-native hot path).

Same flags (baddiffusion.py:53-82), modes (:16-20), TrainingConfig defaults (:84-128), result-folder naming
(:130-134), JSON artefacts (args.json / config.json / sampling.json / measure.json / score.json) and directory
layout (ckpt/, data.ckpt, epochs/epN, samples/, backdoor_samples/, measure/...) as the reference; underneath, the
train loop is baddiffusion_b200.train.Trainer (fused CUDA graphs, one NCCL all-reduce per step under torchrun) and
sampling is the CUDA-graph DDPM/DDIM pipeline.

Out of scope here: HF-hub checkpoint / dataset download, wandb / tensorboard trackers, FID
(pytorch-fid + Inception weights are not available offline).  Data comes from `--dataset` as
  * a local tensor file  $BD_DATA_DIR/<DATASET>.pt  (uint8 NHWC or float NCHW in [-1,1]) if present, else
  * the synthetic generator of SURVEY.md 8(d).
Launch multi-GPU runs with:  torchrun --nproc-per-node N --master-addr 127.0.0.1 baddiffusion.py --mode train ...
"""
import argparse
import json
import os
import sys
import traceback
from dataclasses import asdict, dataclass

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from baddiffusion_b200.dataset import Backdoor, PoisonedBatch, SyntheticDataset, normalize  # noqa: E402
from baddiffusion_b200.model import (DiffuserModelSched, backdoor_metrics, batch_sampling, batch_sampling_save,  # noqa: E402
                                     batch_sampling_u8, save_imgs, save_imgs_u8, shard_for_rank)

MODE_TRAIN, MODE_RESUME, MODE_SAMPLING, MODE_MEASURE, MODE_TRAIN_MEASURE = "train", "resume", "sampling", "measure", "train+measure"
DATASETS = {"MNIST": (28, 1), "CIFAR10": (32, 3), "CELEBA": (64, 3), "CELEBA-HQ": (256, 3)}
DEFAULT_LEARNING_RATE_32, DEFAULT_LEARNING_RATE_256 = 2e-4, 8e-5
MODE_RESUME_OPTS = ["project", "mode", "gpu", "ckpt"]
MODE_SAMPLING_OPTS = ["project", "mode", "eval_max_batch", "gpu", "fclip", "ckpt", "sample_ep", "sched"]
MODE_MEASURE_OPTS = MODE_SAMPLING_OPTS
IGNORE_ARGS = ["overwrite", "is_save_all_model_epochs"]
SCHEDS = ["DDPM-SCHED", "DDIM-SCHED", "DPM_SOLVER_PP_O1-SCHED", "DPM_SOLVER_O1-SCHED", "DPM_SOLVER_PP_O2-SCHED",
          "DPM_SOLVER_O2-SCHED", "DPM_SOLVER_PP_O3-SCHED", "DPM_SOLVER_O3-SCHED", "UNIPC-SCHED", "PNDM-SCHED",
          "DEIS-SCHED", "HEUN-SCHED", "SCORE-SDE-VE-SCHED"]


def parse_args(argv=None):
    p = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    p.add_argument("--project", "-pj", type=str)
    p.add_argument("--mode", "-m", required=True, type=str, choices=[MODE_TRAIN, MODE_RESUME, MODE_SAMPLING, MODE_MEASURE, MODE_TRAIN_MEASURE])
    p.add_argument("--dataset", "-ds", type=str, choices=list(DATASETS))
    p.add_argument("--batch", "-b", type=int)
    p.add_argument("--sched", "-sc", type=str, choices=SCHEDS)
    p.add_argument("--eval_max_batch", "-eb", type=int)
    p.add_argument("--epoch", "-e", type=int)
    p.add_argument("--learning_rate", "-lr", type=float)
    p.add_argument("--clean_rate", "-cr", type=float)
    p.add_argument("--poison_rate", "-pr", type=float)
    p.add_argument("--trigger", "-tr", type=str)
    p.add_argument("--target", "-ta", type=str)
    p.add_argument("--dataset_load_mode", "-dlm", type=str, choices=["FIXED", "FLEX"])
    p.add_argument("--gpu", "-g", type=str)
    p.add_argument("--ckpt", "-c", type=str)
    p.add_argument("--overwrite", "-o", action="store_true")
    p.add_argument("--postfix", "-p", type=str)
    p.add_argument("--fclip", "-fc", type=str, choices=["w", "o"])
    p.add_argument("--save_image_epochs", "-sie", type=int)
    p.add_argument("--save_model_epochs", "-sme", type=int)
    p.add_argument("--is_save_all_model_epochs", "-isame", action="store_true")
    p.add_argument("--sample_ep", "-se", type=int)
    p.add_argument("--result", "-res", type=str)
    # additions (not in the reference): size of the synthetic epoch, steps cap for smoke runs
    p.add_argument("--dataset_size", type=int, default=None, help="synthetic images per epoch (default 50000)")
    p.add_argument("--max_steps", type=int, default=None, help="stop after this many optimizer steps")
    return p.parse_args(argv)


@dataclass
class TrainingConfig:  # baddiffusion.py:84-128
    project: str = "Default"
    batch: int = 512
    epoch: int = 50
    eval_max_batch: int = 256
    learning_rate: float = None
    clean_rate: float = 1.0
    poison_rate: float = 0.007
    trigger: str = Backdoor.TRIGGER_BOX_14
    target: str = Backdoor.TARGET_CORNER
    dataset_load_mode: str = "FIXED"
    gpu: str = "0"
    ckpt: str = None
    overwrite: bool = False
    postfix: str = ""
    fclip: str = "o"
    save_image_epochs: int = 20
    save_model_epochs: int = 5
    is_save_all_model_epochs: bool = False
    sample_ep: int = None
    result: str = "."
    eval_sample_n: int = 16
    measure_sample_n: int = 2048
    batch_32: int = 128
    batch_256: int = 64
    gradient_accumulation_steps: int = 1
    learning_rate_32_scratch: float = 2e-4
    learning_rate_256_scratch: float = 2e-5
    lr_warmup_steps: int = 500
    mixed_precision: str = "fp16"
    seed: int = 0
    dataset_path: str = "datasets"
    ckpt_dir: str = "ckpt"
    data_ckpt_dir: str = "data.ckpt"
    ep_model_dir: str = "epochs"
    ckpt_path: str = None
    data_ckpt_path: str = None


def naming_fn(config):  # baddiffusion.py:130-134
    add_on = f"_{config.postfix}" if config.postfix else ""
    ck = os.path.basename(os.path.normpath(str(config.ckpt)))  # local checkpoint directories: name the run after the id
    return f"res_{ck}_{config.dataset}_ep{config.epoch}_c{config.clean_rate}_p{config.poison_rate}_{config.trigger}-{config.target}{add_on}"


def setup(args):
    """baddiffusion.py:144-248 (argument overlay, result dir, args.json / config.json)."""
    args_d = {k: v for k, v in vars(args).items() if v is not None and v is not False or k == "mode"}
    config = TrainingConfig()
    if args.mode in (MODE_TRAIN, MODE_TRAIN_MEASURE):
        if args.sample_ep is not None:
            raise NotImplementedError("Argument 'sample_ep' shouldn't be used in mode train")
        for need in ("dataset", "ckpt"):
            if getattr(args, need) is None:
                raise ValueError(f"--{need} is required in mode {args.mode}")
    else:
        if args.ckpt is None:
            raise ValueError(f"--ckpt is required in mode {args.mode}")
        allowed = {MODE_RESUME: MODE_RESUME_OPTS, MODE_SAMPLING: MODE_SAMPLING_OPTS, MODE_MEASURE: MODE_MEASURE_OPTS}[args.mode]
        for k in args_d:
            if k not in allowed + IGNORE_ARGS + ["dataset_size", "max_steps", "result"]:
                raise NotImplementedError(f"Argument: {k}={args_d[k]} should not be set in mode {args.mode}")
        with open(os.path.join(args.ckpt, "args.json")) as f:  # reload the run's own arguments (:154-161)
            saved = json.load(f)
        for k, v in saved.items():
            if k not in args_d:
                args_d[k] = v
    for k, v in args_d.items():
        setattr(config, k, v)
    if getattr(config, "dataset", None) is None:
        raise ValueError("--dataset is required")
    size, _ = DATASETS[config.dataset]
    # gradient accumulation and default learning rate (:195-217): the effective batch is batch_32 (MNIST / CIFAR10) or
    # batch_256 (CELEBA, CELEBA-HQ); --batch is the micro-batch and must divide it
    small = config.dataset in ("CIFAR10", "MNIST")
    bs = config.batch_32 if small else config.batch_256
    if config.learning_rate is None:
        if config.ckpt is None:
            config.learning_rate = config.learning_rate_32_scratch if small else config.learning_rate_256_scratch
        else:
            config.learning_rate = DEFAULT_LEARNING_RATE_32 if small else DEFAULT_LEARNING_RATE_256
    if bs % config.batch != 0:
        raise ValueError(f"batch size {config.batch} should be divisible to {bs} for dataset {config.dataset}")
    if bs < config.batch:
        raise ValueError(f"batch size {config.batch} should be smaller or equal to {bs} for dataset {config.dataset}")
    config.gradient_accumulation_steps = int(bs // config.batch)
    if args.mode in (MODE_TRAIN, MODE_TRAIN_MEASURE):
        config.output_dir = os.path.join(config.result, naming_fn(config))
    else:
        config.output_dir = args.ckpt
    config.ckpt_path = os.path.join(config.output_dir, config.ckpt_dir)
    config.data_ckpt_path = os.path.join(config.output_dir, config.data_ckpt_dir)
    if args.mode in (MODE_TRAIN, MODE_TRAIN_MEASURE) and not config.overwrite and os.path.isdir(config.output_dir):
        raise ValueError(f"Output directory: {config.output_dir} has already been created, please set overwrite flag "
                         f"--overwrite or -o")  # :222-223
    if _rank() == 0:
        os.makedirs(config.output_dir, exist_ok=True)
        if args.mode in (MODE_TRAIN, MODE_TRAIN_MEASURE):
            with open(os.path.join(config.output_dir, "args.json"), "w") as f:
                json.dump({k: v for k, v in args_d.items()}, f, indent=4)
        with open(os.path.join(config.output_dir, "config.json" if args.mode not in (MODE_SAMPLING, MODE_MEASURE) else f"{args.mode}.json"), "w") as f:
            json.dump({k: v for k, v in vars(config).items() if not k.startswith("_")}, f, indent=4, default=str)
    return config


def _rank():
    return int(os.environ.get("RANK", "0"))


def _world():
    return int(os.environ.get("WORLD_SIZE", "1"))


class Data:
    """Poisoned dataset of the run: fixed poison split by index like DatasetLoader MODE_FIXED (dataset.py:162-201),
    from a local tensor file when available, synthetic otherwise."""

    def __init__(self, config):
        self.size, self.channel = DATASETS[config.dataset]
        self.batch = config.batch
        bd = Backdoor(root=config.dataset_path)
        self.trigger = bd.get_trigger(type=config.trigger, channel=self.channel, image_size=self.size)
        self.target = bd.get_target(type=config.target, trigger=self.trigger)
        path = os.path.join(os.environ.get("BD_DATA_DIR", config.dataset_path), f"{config.dataset}.pt")
        self.images = None
        if os.path.isfile(path):
            t = torch.load(path)
            if t.dtype == torch.uint8:  # NHWC uint8 -> NCHW [-1,1) with the reference's normalize (quirk Q7)
                t = normalize(t.permute(0, 3, 1, 2).float() / 255.0, vmin_in=0.0, vmax_in=1.0, vmin_out=-1.0, vmax_out=1.0)
            self.images = t.float()
            self.n = len(self.images)
        else:
            self.n = getattr(config, "dataset_size", None) or 50000
            self.synth = SyntheticDataset(self.size, self.channel, poison_rate=max(config.poison_rate, 1e-9), n=self.n,
                                          trigger=config.trigger, target=config.target)
        # fixed split: the first round(n * poison_rate) indices of a seeded permutation are poisoned
        g = torch.Generator().manual_seed(config.seed)
        perm = torch.randperm(self.n, generator=g)
        self.poison = torch.zeros(self.n, dtype=torch.bool)
        self.poison[perm[: int(round(self.n * config.poison_rate))]] = True
        self.num_batch = self.n // (self.batch * _world())

    def epoch_batches(self, epoch, rank, world):
        g = torch.Generator().manual_seed(1000 + epoch)
        order = torch.randperm(self.n, generator=g)
        per = self.batch * world
        for i in range(self.num_batch):
            idx = order[i * per + rank * self.batch: i * per + (rank + 1) * self.batch]
            if self.images is not None:
                img = self.images[idx]
                flip = torch.rand(len(idx), generator=g) < 0.5  # RandomHorizontalFlip (dataset.py:127-128, quirk Q6)
                img = torch.where(flip[:, None, None, None], img.flip(-1), img)
            else:
                gi = torch.Generator().manual_seed(int(idx[0]) * 7919 + epoch)
                img = torch.randn(len(idx), self.channel, self.size, self.size, generator=gi).clamp(-1, 1)
            isp = self.poison[idx].to(torch.uint8)
            if torch.cuda.is_available():
                img, isp = img.pin_memory(), isp.pin_memory()
            yield PoisonedBatch(img, isp)


def sampling(config, file_name, pipeline, data):
    """baddiffusion.py:366-419: 16 clean + 16 backdoor samples as 4x4 grids."""
    from PIL import Image

    def grid(images, rows, cols):
        w, h = images[0].size
        g = Image.new("RGB", (cols * w, rows * h))
        for i, im in enumerate(images):
            g.paste(im, box=(i % cols * w, i // cols * h))
        return g

    S, C = data.size, data.channel
    noise = torch.randn((config.eval_sample_n, C, S, S), generator=torch.Generator().manual_seed(config.seed))
    for name, init in (("samples", noise), ("backdoor_samples", noise + data.trigger[None])):  # quirk Q8
        # each gen_samples call re-seeds (:375): clean and backdoor samples see the SAME per-step noise stream
        res = pipeline(batch_size=config.eval_sample_n, generator=torch.Generator().manual_seed(config.seed), init=init,
                       output_type=None, save_every_step=True)
        d = os.path.join(config.output_dir, name)
        os.makedirs(d, exist_ok=True)
        tag = f"{file_name:04d}" if isinstance(file_name, int) else f"{file_name}"
        clip_tag = "_noclip" if config.fclip != "w" else ""
        grid(pipeline.numpy_to_pil(res.images), 4, 4).save(os.path.join(d, f"{tag}{clip_tag}.png"))
        grid(pipeline.numpy_to_pil(res.movie[0]), 4, 4).save(os.path.join(d, f"{tag}{clip_tag}_sample_t0.png"))
    _check_device_errors()


def measure(config, pipeline, data, rank, world):
    """baddiffusion.py:477-551 minus FID (pytorch-fid unavailable offline): MSE and SSIM of the backdoor samples to the
    target; samples are sharded across ranks, the only collective is the sum of three scalars in backdoor_metrics."""
    S, C, N = data.size, data.channel, int(os.environ.get("BD_MEASURE_N", config.measure_sample_n))
    tag = "_noclip" if config.fclip != "w" else ""
    root = os.path.join(config.output_dir, "measure" if config.sample_ep is None else f"measure/ep{config.sample_ep}")
    noise = torch.randn((N, C, S, S), generator=torch.Generator().manual_seed(config.seed))
    lo, hi = shard_for_rank(N, rank, world)
    rng = torch.Generator().manual_seed(config.seed + rank)
    batch_sampling_save(hi - lo, pipeline, os.path.join(root, f"clean{tag}"), init=noise[lo:hi],
                        max_batch_n=config.eval_max_batch, rng=rng, start_cnt=lo)
    # backdoor samples stay on the GPU as the uint8 pixels of the PNG files; MSE and SSIM come from one device kernel
    # (no save -> ImagePathDataset re-read as in baddiffusion.py:539), the files are still written for the user
    bd_u8 = batch_sampling_u8(hi - lo, pipeline, init=noise[lo:hi] + data.trigger[None], max_batch_n=config.eval_max_batch, rng=rng)
    save_imgs_u8(bd_u8, os.path.join(root, f"backdoor{tag}"), start_cnt=lo)
    mse_sc, ssim_sc = backdoor_metrics(bd_u8, data.target)
    score = {"FID": None, "MSE": mse_sc, "SSIM": ssim_sc,
             "note": "FID needs pytorch-fid + Inception weights (unavailable offline); MSE / SSIM: backdoor samples vs target, "
                     "computed on the device from the uint8 pixels of the saved PNGs (torchmetrics SSIM defaults)"}
    if rank == 0:
        with open(os.path.join(config.output_dir, "score.json"), "w") as f:
            json.dump(score, f, indent=4)
    return score


def _check_device_errors():
    """The tcgen05 pipelines flag a barrier time-out on the device instead of hanging; surface it on the host at the
    cheap cadences of the run (loss log, checkpoint, end of sampling) so a stalled kernel never trains on silently."""
    from baddiffusion_b200 import _lib

    if _lib.lib().bd_umma_error() != 0:
        raise RuntimeError("libb200bd: a tcgen05 pipeline barrier timed out (results of this run are invalid)")


def checkpoint(config, trainer, pipeline, epoch, step):
    """baddiffusion.py:558-570: pipeline in the diffusers layout + optimizer state + {'epoch','step'}."""
    torch.cuda.synchronize()
    _check_device_errors()
    pipeline.save_pretrained(config.output_dir)
    os.makedirs(config.ckpt_path, exist_ok=True)
    torch.save({"exp_avg": trainer.m.cpu(), "exp_avg_sq": trainer.v.cpu(), "state": trainer.state.cpu(),
                "step": trainer.step_dev.cpu(), "iter": trainer.iter_dev.cpu()}, os.path.join(config.ckpt_path, "optimizer.bin"))
    torch.save({"epoch": epoch, "step": step}, config.data_ckpt_path)
    if config.is_save_all_model_epochs:
        pipeline.save_pretrained(os.path.join(config.output_dir, config.ep_model_dir, f"ep{epoch}"))


def main(argv=None):
    args = parse_args(argv)
    world, rank = _world(), _rank()
    if args.gpu is not None and world == 1:
        os.environ.setdefault("CUDA_VISIBLE_DEVICES", args.gpu.split(",")[0])  # one process per GPU: use torchrun for more
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    pg = None
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
        pg = torch.distributed.group.WORLD
    config = setup(args)
    data = Data(config)
    clip = True if config.fclip == "w" else False
    ckpt = config.ckpt if args.mode in (MODE_TRAIN, MODE_TRAIN_MEASURE) else config.output_dir
    if config.sample_ep is not None and args.mode in (MODE_SAMPLING, MODE_MEASURE):
        ckpt = os.path.join(config.output_dir, config.ep_model_dir, f"ep{config.sample_ep}")
    model, noise_sched, get_pipeline = DiffuserModelSched.get_pretrained(
        ckpt=ckpt, clip_sample=clip, noise_sched_type=getattr(config, "sched", None) if args.mode in (MODE_SAMPLING, MODE_MEASURE) else None)
    model = model.cuda()
    pipeline = get_pipeline(model, noise_sched)
    pipeline.set_progress_bar_config(disable=rank != 0)

    if args.mode in (MODE_SAMPLING,):
        if rank == 0:
            sampling(config, "final" if config.sample_ep is None else config.sample_ep, pipeline, data)
        return 0
    if args.mode == MODE_MEASURE:
        print(json.dumps(measure(config, pipeline, data, rank, world)))
        return 0

    from baddiffusion_b200.schedulers import DDPMScheduler
    from baddiffusion_b200.train import Trainer

    train_sched = noise_sched if isinstance(noise_sched, DDPMScheduler) else DDPMScheduler.from_config(noise_sched.config)
    total = data.num_batch * config.epoch
    trainer = Trainer(model, train_sched, config.batch, data.trigger, data.target, lr=config.learning_rate,
                      total_steps=max(total, 1), warmup_steps=config.lr_warmup_steps, process_group=pg, seed=config.seed + rank,
                      accum_steps=config.gradient_accumulation_steps)
    cur_epoch, cur_step = 0, 0
    if args.mode == MODE_RESUME and os.path.isfile(config.data_ckpt_path):
        st = torch.load(config.data_ckpt_path)
        cur_epoch, cur_step = st["epoch"], st["step"]  # quirk Q12: the saved epoch is re-run
        opt = torch.load(os.path.join(config.ckpt_path, "optimizer.bin"))
        trainer.m.copy_(opt["exp_avg"]); trainer.v.copy_(opt["exp_avg_sq"]); trainer.state.copy_(opt["state"])
        trainer.step_dev.copy_(opt["step"]); trainer.iter_dev.copy_(opt["iter"])
    try:  # baddiffusion.py:572-645
        for epoch in range(cur_epoch, config.epoch):
            cur_epoch = epoch  # the `finally` checkpoint saves the epoch the loop was in (:643)
            for batch in data.epoch_batches(epoch, rank, world):
                loss = trainer.step(batch.image, batch.is_poison)
                cur_step += 1
                if rank == 0 and cur_step % 50 == 0:
                    print(f"epoch {epoch} step {cur_step} loss {float(loss):.5f} grad_norm {trainer.grad_norm:.4f} scale {trainer.loss_scale:g}", flush=True)
                    _check_device_errors()
                if args.max_steps and cur_step >= args.max_steps:
                    break
            if rank == 0:
                if (epoch + 1) % config.save_image_epochs == 0 or epoch == config.epoch - 1:
                    sampling(config, epoch, pipeline, data)
                if (epoch + 1) % config.save_model_epochs == 0 or epoch == config.epoch - 1:
                    checkpoint(config, trainer, pipeline, epoch, cur_step)
            if args.max_steps and cur_step >= args.max_steps:
                break
    except Exception:  # the reference swallows, prints and still checkpoints (:635-645)
        traceback.print_exc()
    finally:
        if rank == 0:
            checkpoint(config, trainer, pipeline, min(config.epoch - 1, cur_epoch), cur_step)
            sampling(config, "final", pipeline, data)
    if args.mode == MODE_TRAIN_MEASURE:
        print(json.dumps(measure(config, pipeline, data, rank, world)))
    if world > 1:
        torch.distributed.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
