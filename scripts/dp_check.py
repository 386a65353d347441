"""torchrun check of the data-parallel step: the bucketed, overlapped all-reduce (up-path gradients reduced while the
rest of backward runs) must train exactly like the single all-reduce after backward.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/dp_check.py
Prints, on rank 0, the parameter / loss differences after K steps and the step time of both modes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from baddiffusion_b200 import _lib
from baddiffusion_b200.dataset import SyntheticDataset
from baddiffusion_b200.model import DiffuserModelSched
from baddiffusion_b200.schedulers import DDPMScheduler
from baddiffusion_b200.train import Trainer
from baddiffusion_b200.unet import UNet2DModel

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
_lib.lib()
B, K = 128, 6
ds = SyntheticDataset(32, 3, poison_rate=0.1, seed=1000 * rank)
hb = ds.batch(B, index=0)
ts = [torch.randint(0, 1000, (B,), generator=torch.Generator().manual_seed(100 * rank + i)).cuda() for i in range(K + 3)]
res = {}
for mode in ("single", "overlap"):
    os.environ["BD_NO_AR_OVERLAP"] = "1" if mode == "single" else "0"
    torch.manual_seed(0)
    model = UNet2DModel(**DiffuserModelSched.ARCH["DDPM-CIFAR10-32"]).cuda()
    tr = Trainer(model, DDPMScheduler(variance_type="fixed_large"), B, ds.trigger, ds.target, lr=2e-4, total_steps=1000,
                 warmup_steps=2, process_group=dist.group.WORLD, seed=1234 + rank)
    tr.load_batch(hb.image, hb.is_poison)
    losses = []
    for i in range(K):
        tr.t.copy_(ts[i])
        tr.step_resident(True)
        losses.append(float(tr.loss))
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(10):
        tr.step_resident(True)
    e1.record()
    torch.cuda.synchronize()
    res[mode] = (model.flat_params.clone(), losses, e0.elapsed_time(e1) / 10, tr.grad_norm)
    # replicas must hold identical parameters after identical averaged gradients
    chk = model.flat_params.double().sum().reshape(1)
    allc = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(allc, chk)
    if rank == 0:
        print(f"[{mode}] losses {['%.5f' % l for l in losses]}  ms/step {res[mode][2]:.3f}  grad_norm {res[mode][3]:.4f}  "
              f"replica checksums equal: {all(float(c) == float(allc[0]) for c in allc)}", flush=True)
    del tr, model
    torch.cuda.empty_cache()
if rank == 0:
    pa, pb = res["single"][0], res["overlap"][0]
    d = float((pa - pb).abs().max())
    print(f"max |param(single) - param(overlap)| after {K + 10} steps: {d:.3e} (param scale {float(pa.abs().max()):.3f}); "
          f"loss diff {max(abs(a - b) for a, b in zip(res['single'][1], res['overlap'][1])):.3e}")
    assert d < 5e-3 and max(abs(a - b) for a, b in zip(res["single"][1], res["overlap"][1])) < 5e-3
    print("dp_check ok")
dist.barrier()
dist.destroy_process_group()
