"""Per-kernel timeline of ONE captured training step (both streams), taken with torch.profiler (CUPTI activity records of
the graph replay -- no replays, no serialisation, unlike ncu).  Writes gpurun_out/timeline_step.csv and prints where the
main stream idles and what it waits for.  AB_ARCH / AB_BATCH as in ab_step.py."""
import json, os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile
from baddiffusion_b200 import _lib
from baddiffusion_b200.dataset import SyntheticDataset
from baddiffusion_b200.model import DiffuserModelSched
from baddiffusion_b200.schedulers import DDPMScheduler
from baddiffusion_b200.train import Trainer
from baddiffusion_b200.unet import UNet2DModel

ARCH = os.environ.get("AB_ARCH", "DDPM-CIFAR10-32")
B = int(os.environ.get("AB_BATCH", "128"))
_lib.lib()
torch.manual_seed(0)
model = UNet2DModel(**DiffuserModelSched.ARCH[ARCH]).cuda()
sched = DDPMScheduler(variance_type="fixed_large")
ds = (SyntheticDataset(256, 3, poison_rate=0.1, trigger="GLASSES", target="CAT") if "256" in ARCH
      else SyntheticDataset(32, 3, poison_rate=0.1))
hb = ds.batch(B)
tr = Trainer(model, sched, B, ds.trigger, ds.target)
tr.load_batch(hb.image, hb.is_poison)
for _ in range(5):
    tr.step_resident(True)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(4):
        tr.step_resident(True)
    torch.cuda.synchronize()
path = os.path.join(tempfile.mkdtemp(), "trace.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
ev.sort(key=lambda e: e["ts"])
# split into steps at the batch-prep kernel; keep the third
starts = [i for i, e in enumerate(ev) if "batch_prep" in e["name"]]
assert len(starts) >= 4, len(starts)
step = ev[starts[2]: starts[3]]
t0 = step[0]["ts"]
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/timeline_step.csv", "w") as f:
    f.write("start_us,dur_us,stream,name\n")
    for e in step:
        f.write(f'{e["ts"] - t0:.2f},{e["dur"]:.2f},{e["args"].get("stream")},"{e["name"][:90]}"\n')
streams = {}
for e in step:
    streams.setdefault(e["args"].get("stream"), []).append(e)
main = max(streams, key=lambda s: len(streams[s]))
end = max(e["ts"] + e["dur"] for e in step) - t0
print(f"{ARCH} B={B}: step window {end:.1f} us, {len(step)} device activities, streams "
      + ", ".join(f"{s}: n={len(v)} busy={sum(e['dur'] for e in v):.0f}us" for s, v in streams.items()))
m = streams[main]
gaps = []
for a, b in zip(m, m[1:]):
    g = b["ts"] - (a["ts"] + a["dur"])
    gaps.append((g, a, b))
print(f"main stream: busy {sum(e['dur'] for e in m):.0f} us, idle {sum(g for g, _, _ in gaps):.0f} us in {len(gaps)} gaps "
      f"(median {sorted(g for g, _, _ in gaps)[len(gaps) // 2]:.2f} us); tail after its last kernel "
      f"{end - (m[-1]['ts'] + m[-1]['dur'] - t0):.1f} us")
others = [e for s, v in streams.items() if s != main for e in v]
print("largest main-stream gaps (us)  [after -> before]  side-stream kernels running in the gap:")
for g, a, b in sorted(gaps, key=lambda x: -x[0])[:25]:
    lo, hi = a["ts"] + a["dur"], b["ts"]
    side = [o["name"][:40] for o in others if o["ts"] < hi and o["ts"] + o["dur"] > lo]
    print(f"  {g:7.2f} at {lo - t0:8.1f}  {a['name'][:44]:44s} -> {b['name'][:44]:44s}  side: {side[:3]}")
# time by kernel on each stream
for s, v in streams.items():
    agg = {}
    for e in v:
        k = e["name"].split("(")[0][:60]
        n, d = agg.get(k, (0, 0.0))
        agg[k] = (n + 1, d + e["dur"])
    print(f"stream {s}{' (main)' if s == main else ''}:")
    for k, (n, d) in sorted(agg.items(), key=lambda x: -x[1][1])[:14]:
        print(f"   {d:8.1f} us  n={n:3d}  avg={d / n:6.1f}  {k}")

# Exposed time: walking each lane in launch order, the part of a kernel that runs after everything earlier on the lane
# has ended (PDL-launched kernels start early and wait, so raw durations overlap).  Forward = everything before the
# loss kernel; backward lanes = the stream that carries the GroupNorm backward (the dependent chain) and the others.
def short(n):
    return n.split("(")[0].replace("void ", "").replace("bd::", "").replace("umma::", "")[:52]


def exposed(evs, t_begin):
    agg, prev = {}, t_begin
    for e in evs:
        end = e["ts"] + e["dur"]
        x = max(0.0, end - max(prev, e["ts"] if prev < e["ts"] else prev))
        prev = max(prev, end)
        n, d = agg.get(short(e["name"]), (0, 0.0))
        agg[short(e["name"])] = (n + 1, d + x)
    return agg


loss_i = next(i for i, e in enumerate(step) if "mse_partial" in e["name"])
t_loss = step[loss_i]["ts"]
fwd_lane = step[loss_i]["args"].get("stream")
fwd = [e for e in step if e["ts"] < t_loss and e["args"].get("stream") == fwd_lane]
print(f"forward: {t_loss - t0:.1f} us")
for k, (n, d) in sorted(exposed(fwd, t0).items(), key=lambda x: -x[1][1])[:12]:
    print(f"   {d:8.1f} us exposed  n={n:3d}  avg={d / n:6.1f}  {k}")
bwd = [e for e in step if e["ts"] >= t_loss]
chain = next(e["args"].get("stream") for e in bwd if "gn_bwd" in e["name"])
print(f"backward + optimizer: {end - (t_loss - t0):.1f} us; chain lane busy "
      f"{sum(d for _, d in exposed([e for e in bwd if e['args'].get('stream') == chain], t_loss).values()):.0f} us, other lanes busy "
      f"{sum(d for _, d in exposed([e for e in bwd if e['args'].get('stream') != chain], t_loss).values()):.0f} us")
for k, (n, d) in sorted(exposed([e for e in bwd if e["args"].get("stream") == chain], t_loss).items(), key=lambda x: -x[1][1])[:16]:
    print(f"   {d:8.1f} us exposed  n={n:3d}  avg={d / n:6.1f}  {k}")
