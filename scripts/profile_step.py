"""One eager (non-graph) training step and one sampling step between cudaProfilerStart/Stop, for
`ncu --profile-from-start off` launch lists (profiles/)."""
import os, sys
os.environ["BD_NO_GRAPH"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from baddiffusion_b200 import _lib
from baddiffusion_b200.dataset import SyntheticDataset
from baddiffusion_b200.model import DiffuserModelSched
from baddiffusion_b200.schedulers import DDPMScheduler
from baddiffusion_b200.train import Trainer
from baddiffusion_b200.unet import UNet2DModel

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
what = sys.argv[2] if len(sys.argv) > 2 else "train"
arch = sys.argv[3] if len(sys.argv) > 3 else "DDPM-CIFAR10-32"
S = 256 if "256" in arch else 32
_lib.lib()
torch.manual_seed(0)
model = UNet2DModel(**DiffuserModelSched.ARCH[arch]).cuda()
sched = DDPMScheduler(variance_type="fixed_large")
ds = SyntheticDataset(S, 3, poison_rate=0.1)
tr = Trainer(model, sched, B, ds.trigger, ds.target, use_graph=False)
hb = ds.batch(B)
tr.load_batch(hb.image, hb.is_poison)
for _ in range(2):
    tr.step_resident(True)
torch.cuda.synchronize()
torch.cuda.profiler.start()
if what == "train":
    tr.step_resident(True)
else:
    eng = model.engine(B, False)
    eng.forward(tr.x_noisy, tr.t)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done", float(tr.loss))
