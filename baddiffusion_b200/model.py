"""Drop-in for the live part of the reference's model.py: the batched sampling drivers (model.py:469-529) and the
checkpoint / scheduler / pipeline factory `DiffuserModelSched` (model.py:531-729).

Multi-GPU sampling: independent noise batches per rank, no collective -- `shard_for_rank` gives each
rank the contiguous slice of `init` it owns and the file-name offset that keeps the reference's `{i}.png`
numbering (model.py:525-526).
"""
from __future__ import annotations

import os
from typing import Optional, Union

import numpy as np
import torch

from .pipelines import DDIMPipeline, DDPMPipeline, PNDMPipeline
from .schedulers import DDIMScheduler, DDPMScheduler, PNDMScheduler
from .unet import UNet2DModel


def _batch_sizes(sample_n: int, init, max_batch_n: int):
    if init is None:
        if sample_n > max_batch_n:
            replica, residual = sample_n // max_batch_n, sample_n % max_batch_n
            return [max_batch_n] * replica + ([residual] if residual > 0 else []), None
        return [sample_n], None
    chunks = torch.split(init, max_batch_n)
    return [len(c) for c in chunks], chunks


def batch_sampling(sample_n: int, pipeline, init: torch.Tensor = None, max_batch_n: int = 256,
                   rng: torch.Generator = None):
    """model.py:469-489: one pipeline call per chunk, ONE rng shared by all chunks."""
    sizes, chunks = _batch_sizes(sample_n, init, max_batch_n)
    out = []
    for i, bs in enumerate(sizes):
        res = pipeline(batch_size=bs, generator=rng, init=None if chunks is None else chunks[i], output_type=None)
        out.append(res.images)
    return np.concatenate(out)


def save_imgs(imgs: np.ndarray, file_dir: Union[str, os.PathLike], file_name: Union[str, os.PathLike] = "",
              start_cnt: int = 0) -> None:
    """model.py:496-502."""
    from PIL import Image

    os.makedirs(file_dir, exist_ok=True)
    images = [Image.fromarray(image) for image in np.squeeze((imgs * 255).round().astype("uint8"))]
    for i, img in enumerate(images):
        img.save(os.path.join(file_dir, f"{file_name}{start_cnt + i}.png"))


def batch_sampling_save(sample_n: int, pipeline, path: Union[str, os.PathLike], init: torch.Tensor = None,
                        max_batch_n: int = 256, rng: torch.Generator = None, start_cnt: int = 0):
    """model.py:504-529 (start_cnt: this rank's offset when the sample set is sharded across GPUs)."""
    sizes, chunks = _batch_sizes(sample_n, init, max_batch_n)
    cnt = start_cnt
    for i, bs in enumerate(sizes):
        res = pipeline(batch_size=bs, generator=rng, init=None if chunks is None else chunks[i], output_type=None)
        save_imgs(imgs=res.images, file_dir=path, file_name="", start_cnt=cnt)
        cnt += bs
    return None


def batch_sampling_u8(sample_n: int, pipeline, init: torch.Tensor = None, max_batch_n: int = 256,
                      rng: torch.Generator = None) -> torch.Tensor:
    """`batch_sampling` whose result stays on the GPU as the uint8 NHWC pixels `save_imgs` would write (one pipeline
    call per chunk, ONE rng): input of `backdoor_metrics` -- no PNG encode / decode between sampling and scoring."""
    sizes, chunks = _batch_sizes(sample_n, init, max_batch_n)
    out = []
    for i, bs in enumerate(sizes):
        res = pipeline(batch_size=bs, generator=rng, init=None if chunks is None else chunks[i], output_type="u8")
        out.append(res.images)
    return torch.cat(out)


def save_imgs_u8(imgs_u8: torch.Tensor, file_dir: Union[str, os.PathLike], file_name: Union[str, os.PathLike] = "",
                 start_cnt: int = 0) -> None:
    """model.py:496-502 for samples that are already uint8 (device or host tensor)."""
    from PIL import Image

    os.makedirs(file_dir, exist_ok=True)
    for i, a in enumerate(imgs_u8.cpu().numpy()):
        Image.fromarray(np.squeeze(a)).save(os.path.join(file_dir, f"{file_name}{start_cnt + i}.png"))


def backdoor_metrics(samples_u8: torch.Tensor, target: torch.Tensor, process_group=None):
    """baddiffusion.py:533-546 on the device: MSE (nn.MSELoss) and SSIM (torchmetrics StructuralSimilarityIndexMeasure
    (data_range=1.0) defaults) of the generated backdoor samples against (target / 2 + 0.5).clamp(0, 1), from the uint8
    pixels the reference would re-read from disk.  One kernel (`bd_image_metrics`); sharded sampling sums the two fp64
    accumulators and the sample count across ranks (the only collective of the measurement path).  -> (mse, ssim)."""
    from . import ops

    B, H, W, C = samples_u8.shape
    acc = ops.image_metrics(samples_u8.contiguous(), target.to(samples_u8.device, torch.float32).contiguous())
    tot = torch.cat([acc, torch.tensor([float(B)], dtype=torch.float64, device=acc.device)])
    if process_group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()
                                     and torch.distributed.get_world_size() > 1):
        torch.distributed.all_reduce(tot, group=process_group)
    se, ss, n = (float(v) for v in tot.cpu())
    if n == 0:
        return float("nan"), float("nan")
    return se / (n * C * H * W), ss / (n * C * (H - 10) * (W - 10))


def shard_for_rank(n: int, rank: int, world: int):
    """Contiguous [lo, hi) slice of n samples owned by `rank` (mirrors torch.split ordering, model.py:478,513)."""
    per = (n + world - 1) // world
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


class DiffuserModelSched:
    """model.py:531-729.  Checkpoint ids that point at the HF hub are mapped to their names only; with no network
    the id must be a local directory in the diffusers layout (Appendix D)."""

    LR_SCHED_CKPT: str = "lr_sched.pth"
    OPTIM_CKPT: str = "optim.pth"
    SDE_VP: str = "SDE-VP"
    SDE_VE: str = "SDE-VE"
    SDE_LDM: str = "SDE-LDM"
    CLIP_SAMPLE_DEFAULT = False
    MODEL_DEFAULT: str = "DEFAULT"
    DDPM_32_DEFAULT: str = "DDPM-32-DEFAULT"
    DDPM_256_DEFAULT: str = "DDPM-256-DEFAULT"
    DDPM_CIFAR10_DEFAULT: str = "DDPM-CIFAR10-DEFAULT"
    DDPM_CELEBA_HQ_DEFAULT: str = "DDPM-CELEBA-HQ-DEFAULT"
    DDPM_CIFAR10_32: str = "DDPM-CIFAR10-32"
    DDPM_CELEBA_HQ_256: str = "DDPM-CELEBA-HQ-256"
    DDPM_SCHED = "DDPM-SCHED"
    DDIM_SCHED = "DDIM-SCHED"
    DPM_SOLVER_PP_O1_SCHED = "DPM_SOLVER_PP_O1-SCHED"
    DPM_SOLVER_O1_SCHED = "DPM_SOLVER_O1-SCHED"
    DPM_SOLVER_PP_O2_SCHED = "DPM_SOLVER_PP_O2-SCHED"
    DPM_SOLVER_O2_SCHED = "DPM_SOLVER_O2-SCHED"
    DPM_SOLVER_PP_O3_SCHED = "DPM_SOLVER_PP_O3-SCHED"
    DPM_SOLVER_O3_SCHED = "DPM_SOLVER_O3-SCHED"
    UNIPC_SCHED = "UNIPC-SCHED"
    PNDM_SCHED = "PNDM-SCHED"
    DEIS_SCHED = "DEIS-SCHED"
    HEUN_SCHED = "HEUN-SCHED"
    LMSD_SCHED = "LMSD-SCHED"
    # model.py:598-630: every one of these is sampled through PNDMPipeline (which rebuilds a PNDMScheduler)
    PNDM_FAMILY = (DPM_SOLVER_PP_O1_SCHED, DPM_SOLVER_O1_SCHED, DPM_SOLVER_PP_O2_SCHED, DPM_SOLVER_O2_SCHED,
                   DPM_SOLVER_PP_O3_SCHED, DPM_SOLVER_O3_SCHED, UNIPC_SCHED, PNDM_SCHED, DEIS_SCHED, HEUN_SCHED, LMSD_SCHED)

    HUB_IDS = {DDPM_CIFAR10_32: "google/ddpm-cifar10-32", DDPM_CELEBA_HQ_256: "google/ddpm-ema-celebahq-256"}

    # architectures of the two checkpoints BadDiffusion fine-tunes
    ARCH = {
        DDPM_CIFAR10_32: dict(
            sample_size=32, in_channels=3, out_channels=3, block_out_channels=(128, 256, 256, 256), layers_per_block=2,
            down_block_types=("DownBlock2D", "AttnDownBlock2D", "DownBlock2D", "DownBlock2D"),
            up_block_types=("UpBlock2D", "UpBlock2D", "AttnUpBlock2D", "UpBlock2D"), attention_head_dim=None,
            norm_eps=1e-6, downsample_padding=0, flip_sin_to_cos=False, freq_shift=1),
        DDPM_CELEBA_HQ_256: dict(
            sample_size=256, in_channels=3, out_channels=3, block_out_channels=(128, 128, 256, 256, 512, 512),
            layers_per_block=2, down_block_types=("DownBlock2D",) * 4 + ("AttnDownBlock2D", "DownBlock2D"),
            up_block_types=("UpBlock2D", "AttnUpBlock2D") + ("UpBlock2D",) * 4, attention_head_dim=None,
            norm_eps=1e-6, downsample_padding=0, flip_sin_to_cos=False, freq_shift=1),
    }
    SCHED = {DDPM_CIFAR10_32: dict(variance_type="fixed_large", clip_sample=True),
             DDPM_CELEBA_HQ_256: dict(variance_type="fixed_small", clip_sample=True)}

    @staticmethod
    def get_sample_clip(clip_sample: bool, clip_sample_default: bool) -> bool:
        return clip_sample if clip_sample is not None else clip_sample_default

    @staticmethod
    def _select_sched(noise_sched, noise_sched_type: Optional[str], clip_sample: Optional[bool]):
        """model.py:592-641 restricted to the samplers on the hot path."""
        if noise_sched_type in (None, DiffuserModelSched.DDPM_SCHED):
            get_pipeline = lambda unet, scheduler: DDPMPipeline(unet=unet, scheduler=scheduler)
            if not isinstance(noise_sched, DDPMScheduler):
                noise_sched = DDPMScheduler.from_config(noise_sched.config)
        elif noise_sched_type == DiffuserModelSched.DDIM_SCHED:
            noise_sched = DDIMScheduler.from_config(noise_sched.config)
            get_pipeline = lambda unet, scheduler: DDIMPipeline(unet=unet, scheduler=scheduler)
        elif noise_sched_type in DiffuserModelSched.PNDM_FAMILY:
            # model.py:598-630: DPM-Solver (++), UniPC, PNDM, DEIS, Heun and LMSD schedulers are all paired with the patched
            # PNDMPipeline, whose constructor rebuilds a PNDMScheduler from their config (pipeline_pndm.py:43) -- the shared
            # keys (1000 linear-beta train steps, 1e-4 .. 0.02, epsilon prediction) are all that survives the rebuild
            c = noise_sched.config
            noise_sched = PNDMScheduler(num_train_timesteps=c.num_train_timesteps, beta_start=c.beta_start, beta_end=c.beta_end)
            clip_used = clip_sample
            get_pipeline = lambda unet, scheduler: PNDMPipeline(unet=unet, scheduler=scheduler, clip_sample=bool(clip_used))
        else:
            raise NotImplementedError(f"noise scheduler {noise_sched_type} is outside the BadDiffusion hot path "
                                      "(DDPM, DDIM and the PNDM-pipeline family are implemented; SCORE-SDE-VE / EDM are not)")
        if clip_sample is not None:
            noise_sched.config.clip_sample = clip_sample  # model.py:639-641
        return noise_sched, get_pipeline

    @staticmethod
    def new_synthetic_checkpoint(arch: str, path: str, seed: int = 0):
        """Random-init weights of the named architecture saved in the diffusers layout: the stand-in for the hub
        checkpoints that cannot be downloaded here (SURVEY.md 8d 'synthetic checkpoint')."""
        g = torch.random.get_rng_state()
        torch.manual_seed(seed)
        unet = UNet2DModel(**DiffuserModelSched.ARCH[arch])
        torch.random.set_rng_state(g)
        sched = DDPMScheduler(**DiffuserModelSched.SCHED[arch])
        DDPMPipeline(unet=unet, scheduler=sched).save_pretrained(path)
        return path

    @staticmethod
    def get_pretrained(ckpt: str, clip_sample: bool = None, noise_sched_type: str = None):
        """model.py:700-725 -> (model, noise_sched, get_pipeline)."""
        path = ckpt
        if ckpt in DiffuserModelSched.HUB_IDS and not os.path.isdir(ckpt):
            local = os.environ.get("BD_CKPT_DIR")
            cand = os.path.join(local, ckpt) if local else None
            if cand and os.path.isdir(cand):
                path = cand
            else:
                raise EnvironmentError(
                    f"{ckpt} maps to the hub id {DiffuserModelSched.HUB_IDS[ckpt]}, which cannot be downloaded here; "
                    f"pass a local diffusers-layout directory or create one with "
                    f"DiffuserModelSched.new_synthetic_checkpoint('{ckpt}', <dir>) and set BD_CKPT_DIR")
        pipe = DDPMPipeline.from_pretrained(path)
        noise_sched, get_pipeline = DiffuserModelSched._select_sched(pipe.scheduler, noise_sched_type, clip_sample)
        return pipe.unet, noise_sched, get_pipeline

    @staticmethod
    def get_model_sched(image_size: int, channels: int, ckpt: str = MODEL_DEFAULT, noise_sched_type: str = None,
                        clip_sample: bool = None):
        """model.py:645-698.  The reference's train-from-scratch branch raises TypeError (quirk Q1); here it builds
        the model.py:657-679 topology with a working scheduler."""
        if ckpt in (DiffuserModelSched.MODEL_DEFAULT, DiffuserModelSched.DDPM_256_DEFAULT):
            unet = UNet2DModel(sample_size=image_size, in_channels=channels, out_channels=channels, layers_per_block=2,
                               block_out_channels=(128, 128, 256, 256, 512, 512),
                               down_block_types=("DownBlock2D",) * 4 + ("AttnDownBlock2D", "DownBlock2D"),
                               up_block_types=("UpBlock2D", "AttnUpBlock2D") + ("UpBlock2D",) * 4)
            sched = DDPMScheduler(num_train_timesteps=1000)
            noise_sched, get_pipeline = DiffuserModelSched._select_sched(sched, noise_sched_type, clip_sample)
            return unet, noise_sched, get_pipeline
        return DiffuserModelSched.get_pretrained(ckpt=ckpt, clip_sample=clip_sample, noise_sched_type=noise_sched_type)
