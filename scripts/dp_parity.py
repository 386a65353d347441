"""torchrun worker of tests/test_train_parity_gpu.py::test_two_rank_nccl_gradient_matches_global_batch.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/dp_parity.py out.json

Every rank trains the product `Trainer(process_group=WORLD)` on its shard (B/world samples of one seeded global batch,
supplied noise / t) with the bucketed-overlapped and the single all-reduce; rank 0 also runs the 1-rank Trainer on the
whole global batch and compares: loss (mean of the rank losses), the DP-averaged flat gradient, parameters after two
optimizer steps, and that the replicas stay bitwise identical."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from baddiffusion_b200 import _lib
from baddiffusion_b200.dataset import Backdoor
from baddiffusion_b200.model import DiffuserModelSched, shard_for_rank
from baddiffusion_b200.schedulers import DDPMScheduler
from baddiffusion_b200.train import Trainer
from baddiffusion_b200.unet import UNet2DModel

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
_lib.lib()
GB = int(os.environ.get("BD_DP_GLOBAL_BATCH", "128"))
S, K = 32, 2
bd = Backdoor(root="datasets")
trig = bd.get_trigger(type="BOX_14", channel=3, image_size=S)
targ = bd.get_target(type="HAT", trigger=trig)
g = torch.Generator().manual_seed(0)
batches = []
for _ in range(K):
    batches.append((torch.randn(GB, 3, S, S, generator=g).clamp(-1, 1), torch.randint(0, 1000, (GB,), generator=g),
                    torch.randn(GB, 3, S, S, generator=g)))
isp = torch.tensor([i % 10 == 0 for i in range(GB)])


def run(pg, lo, hi, overlap):
    os.environ["BD_NO_AR_OVERLAP"] = "0" if overlap else "1"
    torch.manual_seed(0)
    model = UNet2DModel(**DiffuserModelSched.ARCH["DDPM-CIFAR10-32"]).cuda()
    p_init = model.flat_params.clone()
    tr = Trainer(model, DDPMScheduler(variance_type="fixed_large"), hi - lo, trig, targ, lr=2e-4, total_steps=100, warmup_steps=1,
                 process_group=pg, seed=1)
    losses, grads = [], []
    for image, t, noise in batches:
        losses.append(float(tr.step(image[lo:hi], isp[lo:hi], noise=noise[lo:hi], t=t[lo:hi])))
        torch.cuda.synchronize()
        grads.append(tr.gflat.clone() / tr.loss_scale)
    assert _lib.lib().bd_umma_error() == 0
    return model.flat_params.clone(), grads, losses, p_init


verdict = {}
lo, hi = shard_for_rank(GB, rank, world)
res = {mode: run(dist.group.WORLD, lo, hi, mode == "overlap") for mode in ("overlap", "single")}
ref = run(None, 0, GB, False) if rank == 0 else None
for mode, (params, grads, losses, _) in res.items():
    lt = torch.tensor(losses, device="cuda", dtype=torch.float64)
    dist.all_reduce(lt)
    lt /= world
    chk = params.double().sum().reshape(1)
    allc = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(allc, chk)
    if rank == 0:
        rp, rg, rl, p0 = ref
        cos = min(float((a @ b) / (a.norm() * b.norm())) for a, b in zip(grads, rg))
        nrel = max(abs(float(a.norm()) - float(b.norm())) / float(b.norm()) for a, b in zip(grads, rg))
        verdict[mode] = {"grad_cos": cos, "grad_norm_rel": nrel,
                         "loss_rel": max(abs(float(a) - b) / abs(b) for a, b in zip(lt, rl)),
                         "param_max_abs": float((params - rp).abs().max()),
                         # Adam moves every element by ~lr whatever its gradient's size: elements whose gradient is within
                         # fp16 noise of zero may flip sign (max-abs up to 2 lr per step), so compare the UPDATE in the 2-norm
                         "param_update_rel": float((params - rp).norm() / (rp - p0).norm()),
                         "replicas_equal": all(float(c) == float(allc[0]) for c in allc)}
if rank == 0:
    print(json.dumps(verdict), flush=True)
    if len(sys.argv) > 1:
        json.dump(verdict, open(sys.argv[1], "w"))
dist.barrier()
dist.destroy_process_group()
