"""GroupNorm backward per shape of the CIFAR10 step (B = 128), two settings of one environment switch side by side
(default: BD_GN_BWD_DZC = 1 / 0, the shared-memory dz cache; `python scripts/ab_gn_bwd.py BD_GN_BWD_SMEM 1 0` compares the
shared-memory-resident kernel with the register kernel).  L2 is flushed before every launch (the operands of a real step
were written long before).  Prints us/launch and the algorithmic TB/s (x + dy + dx [+ add] at 2 bytes)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from baddiffusion_b200 import ops, _lib
_lib.lib()
B, G = int(os.environ.get("AB_BATCH", "128")), 32
KEY, VA, VB = (sys.argv[1:4] if len(sys.argv) >= 4 else ("BD_GN_BWD_DZC", "1", "0"))
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for (C, H, ld, nadd) in [(128, 32, 128, 0), (128, 32, 128, 1), (128, 32, 128, 2), (256, 32, 256, 1), (384, 32, 384, 1), (128, 32, 256, 1),
                         (256, 16, 256, 1), (512, 16, 512, 1), (384, 16, 384, 1), (256, 8, 256, 1), (512, 8, 512, 1), (256, 4, 256, 1)]:
    mk = lambda: torch.randn(B, H, H, ld, device="cuda").half()[..., :C]
    x, dy, dx, a1, a2 = mk(), mk(), mk(), mk(), mk()
    gamma, beta = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
    stats = torch.empty(B, G, 2, device="cuda"); stats[..., 0] = 0; stats[..., 1] = 1
    work = torch.empty(ops.gn_workspace_floats(B, C), device="cuda")
    parts = torch.empty(B, 2 * C, device="cuda"); gsum = torch.empty(B, C, device="cuda")
    res = []
    for mode in (VA, VB):
        os.environ[KEY] = mode
        ts = []
        for it in range(12):
            flush.zero_()
            e0.record()
            ops.groupnorm_bwd(x, dy, dx, gamma, beta, stats, None, None, work, G, True, add_dx=a1 if nadd else None,
                              add_dx2=a2 if nadd > 1 else None, gsum=gsum, parts=parts)
            e1.record(); torch.cuda.synchronize()
            ts.append(1e3 * e0.elapsed_time(e1))
        ts = sorted(ts[2:])
        res.append(ts[len(ts) // 2])
    mb = B * H * H * C * 2 * (3 + nadd) / 1e6
    print(f"C={C:4d} H={H:3d} ld={ld:4d} add={nadd}: {KEY}={VA} {res[0]:6.1f} us ({mb / res[0]:5.2f} TB/s)   {KEY}={VB} {res[1]:6.1f} us ({mb / res[1]:5.2f} TB/s)", flush=True)
