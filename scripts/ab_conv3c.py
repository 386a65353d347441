"""A/B inside one process: persistent conv3p kernel (BD_CONV3C=0) vs the cluster kernel conv3c with weight multicast
(BD_CONV3C=2 / 4) at the 32x32 shapes of the CIFAR10 UNet (and one CelebA-HQ shape), forward and dgrad orientation.
Outputs are compared bitwise (same accumulation order).  L2 is flushed before every timed launch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from baddiffusion_b200 import _lib, ops

_lib.lib()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def timed(fn):
    fn(); fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return 1e3 * ts[len(ts) // 2]  # median, us


shapes = [(128, 32, 32, 128, 128, False), (128, 32, 32, 128, 128, True), (128, 32, 32, 256, 128, True), (128, 32, 32, 256, 256, False),
          (4, 256, 256, 128, 128, False), (4, 128, 128, 128, 128, True), (4, 64, 64, 256, 256, False)]
variants = [v for v in os.environ.get("AB_VARIANTS", "0,2,4").split(",")]
for (B, H, W, Cin, Cout, res) in shapes:
    x = torch.randn(B, H, W, Cin, device="cuda").half()
    w = (torch.randn(9, Cout, Cin, device="cuda") / (3 * Cin ** 0.5)).half()
    bias = torch.randn(Cout, device="cuda")
    rowb = torch.randn(B, Cout, device="cuda")
    r = torch.randn(B, H, W, Cout, device="cuda").half() if res else None
    fl = 2.0 * B * H * W * Cin * Cout * 9
    ys, dxs, line = [], [], []
    for v in variants:
        os.environ["BD_CONV3C"] = v
        y = torch.zeros(B, H, W, Cout, dtype=torch.half, device="cuda")
        tf = timed(lambda: ops.conv_fwd(x, w, y, ksize=3, bias=bias, rowbias=rowb, residual=r, scale=0.5 if res else 1.0, impl=_lib.BD_IMPL_UMMA))
        dx = torch.zeros(B, H, W, Cin, dtype=torch.half, device="cuda")
        td = timed(lambda: ops.conv_dgrad(y, w, dx, ksize=3, impl=_lib.BD_IMPL_UMMA))
        torch.cuda.synchronize()
        ys.append(y); dxs.append(dx)
        line.append(f"C={v}: fwd {tf:.1f} us ({fl / tf / 1e6:.0f} TF/s) dgrad {td:.1f} us ({fl / td / 1e6:.0f} TF/s)")
    eq = [bool(torch.equal(ys[0], yy)) and bool(torch.equal(dxs[0], dd)) for yy, dd in zip(ys, dxs)]
    md = [float((ys[0].float() - yy.float()).abs().max()) for yy in ys]
    print(f"conv3 B={B} {H}x{W} {Cin}->{Cout} res={res}: " + " | ".join(line) + f" | equal={eq} maxdiff={md} umma_error={_lib.lib().bd_umma_error()}", flush=True)
os.environ.pop("BD_CONV3C", None)

# ---- back-to-back timing over rotating buffers (no 2 us event quantisation, inputs colder than L2) + per-tile timeline
B, H, W, Cin, Cout = 128, 32, 32, 128, 128
xs = [torch.randn(B, H, W, Cin, device="cuda").half() for _ in range(6)]
ys = [torch.empty(B, H, W, Cout, dtype=torch.half, device="cuda") for _ in range(6)]
w = (torch.randn(9, Cout, Cin, device="cuda") / 34).half()
for v in variants:
    os.environ["BD_CONV3C"] = v
    for i in range(6):
        ops.conv_fwd(xs[i], w, ys[i], ksize=3, impl=_lib.BD_IMPL_UMMA)
    torch.cuda.synchronize()
    n = 48
    e0.record()
    for i in range(n):
        ops.conv_fwd(xs[i % 6], w, ys[i % 6], ksize=3, impl=_lib.BD_IMPL_UMMA)
    e1.record()
    torch.cuda.synchronize()
    us = 1e3 * e0.elapsed_time(e1) / n
    print(f"b2b rotating 128->128@32x32 B=128, BD_CONV3C={v}: {us:.2f} us/launch ({2.0 * B * H * W * Cin * Cout * 9 / us / 1e6:.0f} TF/s)", flush=True)
    dbg = torch.zeros(160, 64, dtype=torch.int64, device="cuda")
    os.environ["BD_CONV3_DBG_PTR"] = str(dbg.data_ptr())
    flush.zero_()
    ops.conv_fwd(xs[0], w, ys[0], ksize=3, impl=_lib.BD_IMPL_UMMA)
    torch.cuda.synchronize()
    os.environ.pop("BD_CONV3_DBG_PTR")
    d = dbg.cpu()
    print(f"  timeline BD_CONV3C={v} (cycles from the CTA's first stamp: start, mma_issued, acc_ready, stored) per tile")
    for cta in (0, 1, 2, 3, 73, 140, 143):
        t0 = int(d[cta][0])
        print(f"    cta {cta:3d}:", [[int(x) - t0 for x in d[cta][4 * i: 4 * i + 4]] for i in range(5) if int(d[cta][4 * i]) != 0])
os.environ.pop("BD_CONV3C", None)
