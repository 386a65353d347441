"""A/B of whole training steps inside ONE process: each variant = a set of environment switches read by the C side
while the step is captured into its CUDA graph.  Usage: python scripts/ab_step.py "BD_GN_V1=1" "BD_GN_VMAX=4,BD_GN_FT=256" ...
(an empty string "" is the default build).  Prints ms/step for every variant, interleaved twice to expose drift."""
import gc, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from baddiffusion_b200 import _lib
from baddiffusion_b200.dataset import SyntheticDataset
from baddiffusion_b200.model import DiffuserModelSched
from baddiffusion_b200.schedulers import DDPMScheduler
from baddiffusion_b200.train import Trainer
from baddiffusion_b200.unet import UNet2DModel

variants = sys.argv[1:] or [""]
ARCH = os.environ.get("AB_ARCH", "DDPM-CIFAR10-32")            # AB_ARCH=DDPM-CELEBA-HQ-256 AB_BATCH=4: BASELINE configs[3]
B, K = int(os.environ.get("AB_BATCH", "128")), 10
_lib.lib()
torch.manual_seed(0)
model = UNet2DModel(**DiffuserModelSched.ARCH[ARCH]).cuda()
sched = DDPMScheduler(variance_type="fixed_large")
ds = (SyntheticDataset(256, 3, poison_rate=0.1, trigger="GLASSES", target="CAT") if "256" in ARCH
      else SyntheticDataset(32, 3, poison_rate=0.1))
hb = ds.batch(B)
KEYS = set()
for v in variants:
    for kv in filter(None, v.split(",")):
        KEYS.add(kv.split("=")[0])


def run(variant):
    for k in KEYS:
        os.environ.pop(k, None)
    for kv in filter(None, variant.split(",")):
        k, val = kv.split("=")
        os.environ[k] = val
    model._engines = {}
    tr = Trainer(model, sched, B, ds.trigger, ds.target)
    tr.load_batch(hb.image, hb.is_poison)
    for _ in range(3):
        tr.step_resident(True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        tr.step_resident(True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    loss = float(tr.loss)
    n = tr.launches_per_step
    del tr
    gc.collect()
    torch.cuda.empty_cache()
    return ms, loss, n


for rnd in range(2):
    for v in variants:
        ms, loss, n = run(v)
        print(f"round {rnd} [{v or 'default'}]: {ms:.3f} ms/step  loss {loss:.4f}  launches {n}", flush=True)
