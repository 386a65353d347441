"""Experiment: does running the train step as TWO concurrent half-batch chains (two engines, two streams, one CUDA graph)
beat one full-batch chain?  The idea: latency-bound small-layer kernels of one chain fill the SMs the other chain leaves
idle.  Times forward + MSE + backward only (the optimizer tail is identical)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from baddiffusion_b200 import _lib, ops
from baddiffusion_b200.engine import UNetEngine
from baddiffusion_b200.model import DiffuserModelSched
from baddiffusion_b200.unet import UNet2DModel

_lib.lib()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
K = 10
torch.manual_seed(0)
model = UNet2DModel(**DiffuserModelSched.ARCH["DDPM-CIFAR10-32"]).cuda()
model.flat_half()
gflat = model.flat_grads(attach=True)
scale = torch.tensor([1024.0], device="cuda")


def make(nb):
    eng = UNetEngine(model, nb, True)
    x = torch.randn(nb, 3, 32, 32, device="cuda")
    t = torch.randint(0, 1000, (nb,), device="cuda")
    tgt = torch.randn(nb, 3, 32, 32, device="cuda")
    d_eps = torch.empty_like(x)
    loss = torch.zeros(1, device="cuda")
    part = torch.empty(1024, device="cuda")
    eng.io["x"], eng.io["t"], eng.io["d_eps"] = x, t, d_eps

    dummy = torch.zeros(8, device="cuda")

    def run():
        eng._fork(dummy.zero_)      # the Trainer forks the weight-shadow refresh / gradient memset here; f_temb joins it
        eng.run_forward()
        ops.mse_fwd_bwd(eng.eps_hat, tgt, loss, d_eps, part, scale)
        eng._join()
        eng.run_backward()
    return run


def capture(fn):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    return g


def timeit(g, tag):
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    print(f"{tag}: {e0.elapsed_time(e1) / K:.3f} ms (fwd + mse + bwd), umma_error={_lib.lib().bd_umma_error()}", flush=True)


full = make(B)
timeit(capture(full), f"one chain, B={B}")
h0, h1 = make(B // 2), make(B // 2)


def sequential():
    h0(); h1()


timeit(capture(sequential), f"two half chains back to back, B={B // 2} each")
s0, s1 = torch.cuda.Stream(), torch.cuda.Stream()


def concurrent():
    cur = torch.cuda.current_stream()
    s0.wait_stream(cur); s1.wait_stream(cur)
    with torch.cuda.stream(s0):
        h0()
    with torch.cuda.stream(s1):
        h1()
    cur.wait_stream(s0); cur.wait_stream(s1)


timeit(capture(concurrent), f"two half chains on two streams, B={B // 2} each")
timeit(capture(full), f"one chain again, B={B}")
