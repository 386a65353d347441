"""One GroupNorm backward launch per kernel variant at the dominant CIFAR10 shape (B=128, 32x32, C=128), for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from baddiffusion_b200 import ops, _lib
_lib.lib()
B, G, C, H = 128, 32, int(os.environ.get("PROF_C", "128")), 32
mk = lambda: torch.randn(B, H, H, C, device="cuda").half()
x, dy, dx, a1 = mk(), mk(), mk(), mk()
gamma, beta = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
stats = torch.empty(B, G, 2, device="cuda"); stats[..., 0] = 0; stats[..., 1] = 1
work = torch.empty(ops.gn_workspace_floats(B, C), device="cuda")
parts = torch.empty(B, 2 * C, device="cuda"); gsum = torch.empty(B, C, device="cuda")
for mode in ("1", "0"):
    os.environ["BD_GN_BWD_SMEM"] = mode
    for _ in range(2):
        ops.groupnorm_bwd(x, dy, dx, gamma, beta, stats, None, None, work, G, True, add_dx=a1, gsum=gsum, parts=parts)
torch.cuda.synchronize()
