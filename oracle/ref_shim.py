"""Import shim for the *reference* implementation (IBM/BadDiffusion + its vendored diffusers 0.16.0.dev0).

TEST INFRASTRUCTURE ONLY.  Used by `scripts/make_goldens.py` in the build container (where
`/root/reference` is mounted read-only) to generate the committed fixtures under `tests/golden/`
and to cross-check the oracle restatement (`oracle/torch_ref.py`).  `/root/reference` does not
exist on the GPU box, so nothing under `-m gpu`, `smoke()` or `bench.py` imports this module.

Why a shim: the vendored diffusers expects huggingface_hub / transformers APIs that newer
installed versions no longer export.  We register stub parent packages and
import the needed submodules directly from the reference tree -- no reference source is copied.
"""
import os
import sys
import types
import warnings

REF_ROOT = os.environ.get("BADDIFFUSION_REF", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "diffusers", "src", "diffusers"))


_loaded = None


def load():
    """Returns a namespace with the reference classes/functions (imports once)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    warnings.filterwarnings("ignore")
    sys.dont_write_bytecode = True  # reference tree is read-only
    from unittest.mock import MagicMock
    import huggingface_hub as H
    import huggingface_hub.constants as C

    if not hasattr(C, "hf_cache_home"):
        C.hf_cache_home = os.path.expanduser("~/.cache/huggingface")

    class _HfFolder:
        @staticmethod
        def get_token():
            return None

    def _offline(*a, **k):
        raise RuntimeError("offline")

    for n, v in (("HfFolder", _HfFolder), ("cached_download", _offline)):
        if n not in dir(H):
            setattr(H, n, v)

    src = os.path.join(REF_ROOT, "diffusers", "src", "diffusers")
    pkg = types.ModuleType("diffusers")
    pkg.__path__ = [src]
    pkg.__file__ = src + "/__init__.py"
    pkg.__version__ = "0.16.0.dev0"
    sys.modules["diffusers"] = pkg
    import diffusers.utils.import_utils as IU

    IU._transformers_available = False
    pp = types.ModuleType("diffusers.pipelines")
    pp.__path__ = [src + "/pipelines"]
    sys.modules["diffusers.pipelines"] = pp

    from diffusers.models.unet_2d import UNet2DModel
    from diffusers.schedulers.scheduling_ddpm import DDPMScheduler
    from diffusers.schedulers.scheduling_ddim import DDIMScheduler
    from diffusers.models.modeling_utils import ModelMixin
    from diffusers.schedulers.scheduling_utils import SchedulerMixin
    from diffusers.pipelines.pipeline_utils import DiffusionPipeline
    from diffusers.pipelines.ddpm.pipeline_ddpm import DDPMPipeline
    from diffusers.pipelines.ddim.pipeline_ddim import DDIMPipeline
    from diffusers.models.resnet import ResnetBlock2D, Upsample2D, Downsample2D
    from diffusers.models.attention import AttentionBlock
    from diffusers.models.embeddings import get_timestep_embedding, TimestepEmbedding
    from diffusers.optimization import get_cosine_schedule_with_warmup

    exported = dict(
        UNet2DModel=UNet2DModel, DDPMScheduler=DDPMScheduler, DDIMScheduler=DDIMScheduler,
        ModelMixin=ModelMixin, SchedulerMixin=SchedulerMixin, DiffusionPipeline=DiffusionPipeline,
        DDPMPipeline=DDPMPipeline, DDIMPipeline=DDIMPipeline,
    )
    for k, v in exported.items():
        setattr(pkg, k, v)

    def _fallback(name):  # unused scheduler/pipeline names imported at model.py:466
        if name.startswith("__"):
            raise AttributeError(name)
        return type(name, (object,), {})

    pkg.__getattr__ = _fallback
    for n in ("matplotlib", "matplotlib.pyplot", "matplotlib.animation", "matplotlib.dates", "comet_ml"):
        sys.modules.setdefault(n, MagicMock())

    cwd = os.getcwd()
    sys.path.insert(0, REF_ROOT)
    os.chdir(REF_ROOT)  # static/*.png are opened by relative path (dataset.py:384-387)
    try:
        from dataset import Backdoor, DatasetLoader
        from loss import q_sample_diffuser, p_losses_diffuser
        from model import DiffuserModelSched, batch_sampling, batch_sampling_save
        from util import normalize
    finally:
        os.chdir(cwd)

    ns = types.SimpleNamespace(
        REF_ROOT=REF_ROOT,
        Backdoor=Backdoor, DatasetLoader=DatasetLoader, normalize=normalize,
        q_sample_diffuser=q_sample_diffuser, p_losses_diffuser=p_losses_diffuser,
        DiffuserModelSched=DiffuserModelSched, batch_sampling=batch_sampling,
        batch_sampling_save=batch_sampling_save,
        ResnetBlock2D=ResnetBlock2D, Upsample2D=Upsample2D, Downsample2D=Downsample2D,
        AttentionBlock=AttentionBlock, get_timestep_embedding=get_timestep_embedding,
        TimestepEmbedding=TimestepEmbedding, get_cosine_schedule_with_warmup=get_cosine_schedule_with_warmup,
        **exported,
    )
    _loaded = ns
    return ns


class chdir_ref:
    """Context manager: run a block with cwd = reference root (for the relative static/ paths)."""

    def __enter__(self):
        self._cwd = os.getcwd()
        os.chdir(REF_ROOT)

    def __exit__(self, *a):
        os.chdir(self._cwd)
