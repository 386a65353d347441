"""A/B timing inside ONE process (boxes differ by ~10 %, so variants are only comparable within a run):
  * GroupNorm forward / backward: three-kernel path (BD_GN_V1=1) vs the single-launch cluster kernels;
  * 3x3 conv at 32x32: per-tile-launch transposed kernel (BD_NO_CONV3P=1) vs the persistent kernel, plus a bitwise
    comparison of their outputs (same accumulation order, same epilogue arithmetic).
The C side reads the switches with getenv() at every call.  L2 is flushed before every timed launch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from baddiffusion_b200 import _lib, ops

_lib.lib()
what = sys.argv[1] if len(sys.argv) > 1 else "all"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def timed(fn):
    fn(); fn()
    tot = 0.0
    for _ in range(reps):
        flush.zero_()
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return 1e3 * tot / reps  # us


def setenv(k, on):
    if on:
        os.environ[k] = "1"
    else:
        os.environ.pop(k, None)


if what in ("all", "gn"):
    G, eps = 32, 1e-6
    for (B, H, C) in [(128, 32, 128), (128, 32, 256), (128, 32, 384), (128, 16, 256), (128, 16, 512), (128, 8, 256), (128, 4, 256)]:
        x = torch.randn(B, H, H, C, device="cuda").half()
        dy = torch.randn(B, H, H, C, device="cuda").half()
        add = torch.randn(B, H, H, C, device="cuda").half()
        y, dx = torch.empty_like(x), torch.empty_like(x)
        gamma, beta = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
        stats = torch.empty(B, G, 2, device="cuda")
        dg, db = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
        gsum = torch.empty(B, C, device="cuda")
        work = torch.empty(ops.gn_workspace_floats(B, C), device="cuda")
        mb = x.numel() * 2 / 1e6
        out = []
        for v1 in (True, False):
            setenv("BD_GN_V1", v1)
            tf = timed(lambda: ops.groupnorm_fwd(x, y, gamma, beta, stats, work, G, eps, True))
            tb = timed(lambda: ops.groupnorm_bwd(x, dy, dx, gamma, beta, stats, dg, db, work, G, True, add_dx=add, gsum=gsum))
            out.append((tf, tb))
        setenv("BD_GN_V1", False)
        print(f"GN B={B} H={H} C={C} ({mb:.1f} MB): fwd v1 {out[0][0]:.1f} us -> fused {out[1][0]:.1f} us "
              f"({2 * mb / out[1][0]:.2f} TB/s alg);  bwd(+add,+gsum) v1 {out[0][1]:.1f} us -> fused {out[1][1]:.1f} us "
              f"({4 * mb / out[1][1]:.2f} TB/s alg)", flush=True)

if what in ("all", "conv"):
    for (B, H, Cin, Cout, res) in [(128, 32, 128, 128, False), (128, 32, 128, 128, True), (128, 32, 256, 128, True), (128, 32, 256, 256, False),
                                   (128, 16, 256, 256, False), (128, 16, 256, 256, True), (128, 16, 512, 256, False), (128, 16, 128, 256, False)]:
        x = torch.randn(B, H, H, Cin, device="cuda").half()
        w = (torch.randn(9, Cout, Cin, device="cuda") / (3 * Cin ** 0.5)).half()
        bias = torch.randn(Cout, device="cuda")
        rowb = torch.randn(B, Cout, device="cuda")
        r = torch.randn(B, H, H, Cout, device="cuda").half() if res else None
        ys = []
        ts = []
        for old in (True, False):
            setenv("BD_NO_CONV3P", old); setenv("BD_NO_CONV3W", old)
            y = torch.zeros(B, H, H, Cout, dtype=torch.half, device="cuda")
            f = lambda: ops.conv_fwd(x, w, y, ksize=3, bias=bias, rowbias=rowb, residual=r, scale=0.5 if res else 1.0, impl=_lib.BD_IMPL_UMMA)
            ts.append(timed(f))
            ys.append(y)
        # dgrad orientation (MN-major weights)
        td = []
        dxs = []
        for old in (True, False):
            setenv("BD_NO_CONV3P", old); setenv("BD_NO_CONV3W", old)
            dx = torch.zeros(B, H, H, Cin, dtype=torch.half, device="cuda")
            dyy = ys[0]
            f = lambda: ops.conv_dgrad(dyy, w, dx, ksize=3, impl=_lib.BD_IMPL_UMMA)
            td.append(timed(f))
            dxs.append(dx)
        setenv("BD_NO_CONV3P", False); setenv("BD_NO_CONV3W", False)
        fl = 2.0 * B * H * H * Cin * Cout * 9
        err = _lib.lib().bd_umma_error()
        print(f"conv3 B={B} H={H} {Cin}->{Cout} res={res}: fwd per-tile {ts[0]:.1f} us ({fl / ts[0] / 1e6:.0f} TF/s) -> persistent "
              f"{ts[1]:.1f} us ({fl / ts[1] / 1e6:.0f} TF/s), equal={torch.equal(ys[0], ys[1])} maxdiff={float((ys[0].float() - ys[1].float()).abs().max()):.3g}; "
              f"dgrad {td[0]:.1f} -> {td[1]:.1f} us ({fl / td[1] / 1e6:.0f} TF/s), equal={torch.equal(dxs[0], dxs[1])} maxdiff={float((dxs[0].float() - dxs[1].float()).abs().max()):.3g} umma_error={err}", flush=True)

if what == "gnsweep":
    G, eps = 32, 1e-6
    for (B, H, C) in [(128, 32, 128), (128, 32, 256), (128, 16, 256), (128, 8, 256)]:
        x = torch.randn(B, H, H, C, device="cuda").half()
        dy = torch.randn(B, H, H, C, device="cuda").half()
        add = torch.randn(B, H, H, C, device="cuda").half()
        y, dx = torch.empty_like(x), torch.empty_like(x)
        gamma, beta = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
        stats = torch.empty(B, G, 2, device="cuda")
        dg, db = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
        gsum = torch.empty(B, C, device="cuda")
        work = torch.empty(ops.gn_workspace_floats(B, C), device="cuda")
        for vmax in (4, 8):
            for ft, bt in ((512, 256), (256, 128), (384, 192)):
                os.environ["BD_GN_VMAX"], os.environ["BD_GN_FT"], os.environ["BD_GN_BT"] = str(vmax), str(ft), str(bt)
                try:
                    tf = timed(lambda: ops.groupnorm_fwd(x, y, gamma, beta, stats, work, G, eps, True))
                    tb = timed(lambda: ops.groupnorm_bwd(x, dy, dx, gamma, beta, stats, dg, db, work, G, True, add_dx=add, gsum=gsum))
                    print(f"GN H={H} C={C} vmax={vmax} ft={ft} bt={bt}: fwd {tf:.1f} us  bwd {tb:.1f} us", flush=True)
                except Exception as e:
                    print(f"GN H={H} C={C} vmax={vmax} ft={ft} bt={bt}: {e}", flush=True)

if what == "timeline":
    B, H, Cin, Cout = 128, 32, 128, 128
    x = torch.randn(B, H, H, Cin, device="cuda").half()
    w = (torch.randn(9, Cout, Cin, device="cuda") / 34).half()
    y = torch.empty(B, H, H, Cout, dtype=torch.half, device="cuda")
    r = torch.randn(B, H, H, Cout, device="cuda").half()
    for res in (None, r):
        ops.conv_fwd(x, w, y, ksize=3, residual=res, impl=_lib.BD_IMPL_UMMA)
        dbg = torch.zeros(256, 64, dtype=torch.int64, device="cuda")
        os.environ["BD_CONV3_DBG_PTR"] = str(dbg.data_ptr())
        flush.zero_()
        ops.conv_fwd(x, w, y, ksize=3, residual=res, impl=_lib.BD_IMPL_UMMA)
        torch.cuda.synchronize()
        os.environ.pop("BD_CONV3_DBG_PTR")
        d = dbg.cpu()
        print("residual" if res is not None else "plain", "(cycles relative to the CTA's first stamp: start, mma_issued, acc_ready, stored) per tile")
        for cta in (0, 1, 73, 107, 108, 147):
            t0 = int(d[cta][0])
            print(f"  cta {cta:3d}:", [[int(v) - t0 for v in d[cta][4 * i: 4 * i + 4]] for i in range(3) if int(d[cta][4 * i]) != 0])

if what == "memref":
    # reference points for the timing regime: a plain copy of the same bytes, with and without the dirty-L2 flush
    G, eps = 32, 1e-6
    B, H, C = 128, 32, 128
    x = torch.randn(B, H, H, C, device="cuda").half()
    y = torch.empty_like(x)
    gamma, beta = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
    stats = torch.empty(B, G, 2, device="cuda")
    work = torch.empty(ops.gn_workspace_floats(B, C), device="cuda")
    xs = [torch.randn(B, H, H, C, device="cuda").half() for _ in range(8)]   # 8 x 33.5 MB > L2: rotating inputs
    ys = [torch.empty_like(x) for _ in range(8)]

    def rot(fn, n=24):
        for i in range(4):
            fn(xs[i % 8], ys[i % 8])
        torch.cuda.synchronize()
        e0.record()
        for i in range(n):
            fn(xs[i % 8], ys[i % 8])
        e1.record()
        torch.cuda.synchronize()
        return 1e3 * e0.elapsed_time(e1) / n

    print(f"copy 33.5 MB (flush before):   {timed(lambda: y.copy_(x)):.1f} us")
    print(f"copy 33.5 MB (rotating, b2b):  {rot(lambda a, b: b.copy_(a)):.1f} us")
    for v1 in (True, False):
        setenv("BD_GN_V1", v1)
        f = lambda a, b: ops.groupnorm_fwd(a, b, gamma, beta, stats, work, G, eps, True)
        print(f"GN fwd {'v1' if v1 else 'fused'} (flush before):  {timed(lambda: f(x, y)):.1f} us")
        print(f"GN fwd {'v1' if v1 else 'fused'} (rotating, b2b): {rot(f):.1f} us")
    setenv("BD_GN_V1", False)
