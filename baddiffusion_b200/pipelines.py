"""Drop-in DDPMPipeline / DDIMPipeline (the BadDiffusion-patched versions: `init=`, `start_from=`,
`save_every_step=`, `movie`; D/pipelines/ddpm/pipeline_ddpm.py:46-125, D/pipelines/ddim/pipeline_ddim.py:50-142).

The denoising loop is ONE captured CUDA graph (UNet forward + fused scheduler step + timestep advance) replayed
per step: the per-step scalars come from a device-resident coefficient table indexed by a device-side step
counter, so no host work or synchronisation happens inside the loop.  With a CPU `torch.Generator` the noise
stream is the reference's (CPU draw per step, copied to the device: quirk Q11) so samples are reproducible
"on identical seeds"; without a generator the noise is drawn in-kernel (Philox).
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass, field
from typing import List, Optional, Union

import numpy as np
import torch

from . import config_utils as CU
from . import ops
from .schedulers import DDIMScheduler, DDPMScheduler, PNDMScheduler
from .unet import UNet2DModel


@dataclass
class ImagePipelineOutput:  # D/pipelines/pipeline_utils.py:111-124 (patched: movie)
    images: Union[List, np.ndarray]
    movie: List = field(default_factory=list)


class _Pipeline:
    model_index_class = "DDPMPipeline"

    def __init__(self, unet: UNet2DModel, scheduler):
        self.unet, self.scheduler = unet, scheduler
        self._graphs = {}
        self._pb_kwargs = {}
        self.use_cuda_graph = os.environ.get("BD_NO_GRAPH", "0") != "1"

    # -- DiffusionPipeline surface used by BadDiffusion
    @property
    def device(self):
        return self.unet.device

    def to(self, device):
        self.unet.to(device)
        self._graphs = {}
        return self

    def set_progress_bar_config(self, **kwargs):
        self._pb_kwargs = kwargs

    def progress_bar(self, iterable):
        if self._pb_kwargs.get("disable", False):
            return iterable
        try:
            from tqdm.auto import tqdm

            return tqdm(iterable, **self._pb_kwargs)
        except Exception:
            return iterable

    @staticmethod
    def numpy_to_pil(images):
        from PIL import Image

        if images.ndim == 3:
            images = images[None, ...]
        images = (images * 255).round().astype("uint8")
        if images.shape[-1] == 1:
            return [Image.fromarray(image.squeeze(), mode="L") for image in images]
        return [Image.fromarray(image) for image in images]

    def save_pretrained(self, save_directory: str, **kwargs):
        """D/pipelines/pipeline_utils.py:527-612 layout: model_index.json + unet/ + scheduler/."""
        os.makedirs(save_directory, exist_ok=True)
        index = {"_class_name": self.model_index_class, "_diffusers_version": CU.DIFFUSERS_VERSION,
                 "scheduler": ["diffusers", type(self.scheduler).__name__], "unet": ["diffusers", "UNet2DModel"]}
        with open(os.path.join(save_directory, "model_index.json"), "w", encoding="utf-8") as f:
            f.write(json.dumps(index, indent=2, sort_keys=True) + "\n")
        self.unet.save_pretrained(os.path.join(save_directory, "unet"))
        self.scheduler.save_pretrained(os.path.join(save_directory, "scheduler"))

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path: str, **kwargs):
        d = pretrained_model_name_or_path
        if not os.path.isdir(d):
            raise EnvironmentError(f"{d} is not a local directory (hub downloads are out of scope: no network)")
        idx = CU.load_config(d, "model_index.json")
        sched_cls = {"DDPMScheduler": DDPMScheduler, "DDIMScheduler": DDIMScheduler, "PNDMScheduler": PNDMScheduler}.get(idx["scheduler"][1])
        if sched_cls is None:
            raise NotImplementedError(f"scheduler {idx['scheduler'][1]} is outside the BadDiffusion hot path")
        unet = UNet2DModel.from_pretrained(d, subfolder="unet")
        scheduler = sched_cls.from_pretrained(d, subfolder="scheduler")
        return cls(unet=unet, scheduler=scheduler)

    # -- the loop
    def _image_shape(self, batch_size):
        ss = self.unet.config.sample_size
        if isinstance(ss, int):
            return (batch_size, self.unet.config.in_channels, ss, ss)
        return (batch_size, self.unet.config.in_channels, *ss)

    @staticmethod
    def _post(image):
        """(image / 2 + 0.5).clamp(0, 1) -> cpu NHWC numpy, via the fused finalize kernel."""
        B, C, H, W = image.shape
        out = torch.empty(B, H, W, C, device=image.device)
        ops.finalize_images(image, out, None)
        return out.cpu().numpy()

    @staticmethod
    def _post_u8(image):
        """round((image / 2 + 0.5).clamp(0, 1) * 255) -> uint8 NHWC ON THE DEVICE: exactly the bytes `save_imgs` would
        put into the PNG files (model.py:499), kept for the device-side MSE / SSIM tail."""
        B, C, H, W = image.shape
        out = torch.empty(B, H, W, C, dtype=torch.uint8, device=image.device)
        ops.finalize_images(image, None, out)
        return out

    def _run_loop(self, image, timesteps, coef_table, generator, ddim, noise_steps, save_every_step, mov):
        """image: (B,C,H,W) fp32 cuda, updated in place.  noise_steps[i] says whether step i consumes noise.
        ddim: False = DDPM step, True = DDIM step, "pndm" = PNDM step (device state: accumulator, saved sample, 4 history slots)."""
        dev = image.device
        B = image.shape[0]
        eng = self.unet.engine(B, False)
        eng.refresh_weights()
        n = len(timesteps)
        ts_dev = torch.as_tensor(np.asarray(timesteps, dtype=np.int64)).to(dev)
        coef = coef_table.to(dev).contiguous()
        cpu_gen = generator is not None and generator.device.type == "cpu"
        if not cpu_gen and ddim != "pndm":
            # Philox step noise: the captured graph bakes (seed, offset), so every call draws ONE id from the caller's
            # generator (which advances it, like the reference's per-step randn calls would) and passes it through the
            # spare column of the coefficient table -- fresh noise per call, reproducible after re-seeding, no re-capture
            sid = float(int(torch.randint(0, 1 << 23, (1,), generator=generator, device=dev).item()))
            coef = coef.clone()
            if ddim:
                coef[:, 7] = sid
            else:
                coef[:, 7] = torch.where(coef[:, 7] != 0, coef[:, 7] + sid, coef[:, 7])
        seed = (generator.initial_seed() if generator is not None else torch.initial_seed()) & ((1 << 63) - 1)
        key = (B, ddim, cpu_gen)
        st = self._graphs.get(key)
        if st is None:
            st = dict(x=torch.empty_like(image), z=torch.empty_like(image) if cpu_gen else None,
                      t_vec=torch.zeros(B, dtype=torch.int64, device=dev), step=torch.zeros(1, dtype=torch.int32, device=dev),
                      graph=None, coef=None, ts=None, seed=None)
            self._graphs[key] = st
        st["x"].copy_(image)
        step_fn = ops.ddim_step if ddim else ops.ddpm_step
        if ddim == "pndm" and st.get("state") is None:
            st["state"] = torch.zeros(6 * image.numel(), device=dev)

        def body():
            eng.io["x"], eng.io["t"] = st["x"], st["t_vec"]
            eng.run_forward()
            if ddim == "pndm":
                ops.pndm_step(st["x"], eng.eps_hat, st["x"], st["state"], st["coef"], st["step"])
            else:
                step_fn(st["x"], eng.eps_hat, st["z"], st["x"], st["coef"], st["step"], seed=st["seed"], offset=0)
            ops.sampler_advance(st["step"], st["ts"], st["t_vec"], False)

        # (re)capture when the tables change size or identity
        need_capture = (st["graph"] is None or st["coef"] is None or st["coef"].shape != coef.shape or st["seed"] != seed)
        if need_capture:
            st["coef"], st["ts"], st["seed"] = coef.clone(), torch.cat([ts_dev, ts_dev[-1:]]).clone(), seed
            if self.use_cuda_graph:
                ops.sampler_advance(st["step"], st["ts"], st["t_vec"], True)
                s = torch.cuda.Stream()
                s.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(s):
                    keep = st["x"].clone()
                    body()  # warm-up (lazy module loading, attribute setting) outside capture
                    st["x"].copy_(keep)
                torch.cuda.current_stream().wait_stream(s)
                g = torch.cuda.CUDAGraph()
                ops.sampler_advance(st["step"], st["ts"], st["t_vec"], True)
                with torch.cuda.graph(g):
                    body()
                st["graph"] = g
                st["x"].copy_(image)
        else:
            st["coef"].copy_(coef)
            st["ts"].copy_(torch.cat([ts_dev, ts_dev[-1:]]))
        ops.sampler_advance(st["step"], st["ts"], st["t_vec"], True)
        for i, t in enumerate(self.progress_bar(timesteps)):
            if cpu_gen and noise_steps[i]:
                # reference noise stream: one CPU draw of the whole batch per step (D/utils/torch_utils.py:47-68)
                st["z"].copy_(torch.randn(image.shape, generator=generator, dtype=torch.float32), non_blocking=True)
            if st["graph"] is not None:
                st["graph"].replay()
            else:
                body()
            if save_every_step:
                mov.append(self._post(st["x"]))
        image.copy_(st["x"])
        return image


class DDPMPipeline(_Pipeline):
    model_index_class = "DDPMPipeline"

    def __init__(self, unet, scheduler):
        super().__init__(unet, scheduler)

    @torch.no_grad()
    def __call__(self, batch_size: int = 1, generator: Optional[torch.Generator] = None, num_inference_steps: int = 1000,
                 start_from: int = 0, output_type: Optional[str] = "pil", init: torch.Tensor = None,
                 save_every_step: bool = False, return_dict: bool = True, **kwargs):
        dev = self.device
        shape = self._image_shape(batch_size)
        if init is None:
            if generator is not None and generator.device.type == "cpu":
                image = torch.randn(shape, generator=generator, dtype=torch.float32).to(dev)
            else:
                image = torch.randn(shape, generator=generator, device=dev, dtype=torch.float32)
        else:
            image = init.detach().clone().to(device=dev, dtype=torch.float32)
        if image.shape[0] != batch_size:
            batch_size = image.shape[0]
        self.scheduler.set_timesteps(num_inference_steps)
        self.scheduler._check_supported()
        mov = []
        if save_every_step:
            mov = [self._post(image)]
        timesteps = [int(t) for t in self.scheduler.timesteps[start_from:]]
        if len(timesteps) > 0 and batch_size > 0:
            table = self.scheduler.coef_table(timesteps)
            self._run_loop(image.contiguous(), timesteps, table, generator, False, [t > 0 for t in timesteps],
                           save_every_step, mov)
        image = self._post_u8(image) if output_type == "u8" else self._post(image)
        if output_type == "pil":
            image = self.numpy_to_pil(image)
            if save_every_step:
                mov = list(map(self.numpy_to_pil, mov))
        if not return_dict:
            return (image,)
        return ImagePipelineOutput(images=image, movie=mov)


class DDIMPipeline(_Pipeline):
    model_index_class = "DDIMPipeline"

    def __init__(self, unet, scheduler):
        scheduler = DDIMScheduler.from_config(scheduler.config)  # pipeline_ddim.py:40
        super().__init__(unet, scheduler)

    @torch.no_grad()
    def __call__(self, batch_size: int = 1, generator: Optional[torch.Generator] = None, eta: float = 0.0,
                 num_inference_steps: int = 50, use_clipped_model_output: Optional[bool] = None,
                 output_type: Optional[str] = "pil", init: torch.Tensor = None, save_every_step: bool = False,
                 return_dict: bool = True, **kwargs):
        dev = self.device
        shape = self._image_shape(batch_size)
        if isinstance(generator, list):
            raise NotImplementedError("per-sample generator lists are not used by BadDiffusion")
        if init is None:
            if generator is not None and generator.device.type == "cpu":
                image = torch.randn(shape, generator=generator, dtype=torch.float32).to(dev)
            else:
                image = torch.randn(shape, generator=generator, device=dev, dtype=torch.float32)
        else:
            image = init.detach().clone().to(device=dev, dtype=torch.float32)
        batch_size = image.shape[0]
        self.scheduler.set_timesteps(num_inference_steps)
        mov = []
        if save_every_step:
            mov = [self._post(image)]
        timesteps = [int(t) for t in self.scheduler.timesteps]
        if len(timesteps) > 0 and batch_size > 0:
            table = self.scheduler.coef_table(eta, bool(use_clipped_model_output), timesteps)
            self._run_loop(image.contiguous(), timesteps, table, generator, True, [eta > 0] * len(timesteps),
                           save_every_step, mov)
        image = self._post_u8(image) if output_type == "u8" else self._post(image)
        if output_type == "pil":
            image = self.numpy_to_pil(image)
            if save_every_step:
                mov = list(map(self.numpy_to_pil, mov))
        if not return_dict:
            return (image,)
        return ImagePipelineOutput(images=image, movie=mov)


class PNDMPipeline(_Pipeline):
    """D/pipelines/pndm/pipeline_pndm.py as patched by the reference (init= / save_every_step= / start_from= / clip_sample):
    the pipeline model.py:598-630 pairs with DPM-Solver, UniPC, DEIS, Heun, LMSD and PNDM schedulers.  Its constructor
    rebuilds a PNDMScheduler from whatever scheduler config it is given (:43), so those --sched choices all sample with
    PNDM (12 Runge-Kutta warm-up UNet calls + linear multistep).  One CUDA graph per denoise step: UNet forward ->
    bd_pndm_step -> timestep advance."""
    model_index_class = "PNDMPipeline"

    def __init__(self, unet, scheduler, clip_sample: bool = False, clip_sample_range: float = 1.0):
        scheduler = PNDMScheduler.from_config(scheduler.config)   # pipeline_pndm.py:43
        super().__init__(unet, scheduler)
        self.clip_sample = clip_sample
        self.clip_sample_range = clip_sample_range

    @torch.no_grad()
    def __call__(self, batch_size: int = 1, num_inference_steps: int = 50, start_from: int = 0,
                 generator: Optional[torch.Generator] = None, output_type: Optional[str] = "pil", init: torch.Tensor = None,
                 save_every_step: bool = False, return_dict: bool = True, **kwargs):
        dev = self.device
        shape = self._image_shape(batch_size)
        if init is None:
            if generator is not None and generator.device.type == "cpu":
                image = torch.randn(shape, generator=generator, dtype=torch.float32).to(dev)
            else:
                image = torch.randn(shape, generator=generator, device=dev, dtype=torch.float32)
        else:
            image = init.detach().clone().to(device=dev, dtype=torch.float32)
        batch_size = image.shape[0]
        mov = []
        if save_every_step:
            mov = [self._post(image)]
        self.scheduler.set_timesteps(num_inference_steps)
        timesteps = [int(t) for t in self.scheduler.timesteps[int(start_from):]]
        if len(timesteps) > 0 and batch_size > 0:
            table = self.scheduler.coef_table(timesteps, clip=float(self.clip_sample_range) if self.clip_sample else 0.0)
            self._run_loop(image.contiguous(), timesteps, table, generator, "pndm", [False] * len(timesteps), save_every_step, mov)
        image = self._post_u8(image) if output_type == "u8" else self._post(image)
        if output_type == "pil":
            image = self.numpy_to_pil(image)
            if save_every_step:
                mov = list(map(self.numpy_to_pil, mov))
        if not return_dict:
            return (image,)
        return ImagePipelineOutput(images=image, movie=mov)
