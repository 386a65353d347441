"""Times the first/last conv kernels (convio.cu) at the CIFAR-10 training shape: B=128, 32x32, 3 <-> 128 channels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from baddiffusion_b200 import _lib, ops
_lib.lib()
B, H, C = 128, 32, 128
x = torch.randn(B, 3, H, H, device="cuda")
w_in = torch.randn(9, C, 3, device="cuda"); b_in = torch.randn(C, device="cuda")
y = torch.empty(B, H, H, C, dtype=torch.half, device="cuda")
dy = torch.randn(B, H, H, C, device="cuda").half()
dw_in, db_in = torch.zeros(9, C, 3, device="cuda"), torch.zeros(C, device="cuda")
h = torch.randn(B, H, H, C, device="cuda").half()
w_out = torch.randn(9, 3, C, device="cuda"); b_out = torch.randn(3, device="cuda")
e = torch.empty(B, 3, H, H, device="cuda"); de = torch.randn(B, 3, H, H, device="cuda")
dh = torch.empty_like(h); dw_out, db_out = torch.zeros(9, 3, C, device="cuda"), torch.zeros(3, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
cases = {
    "conv_in_fwd": lambda: ops.conv_in_fwd(x, w_in, b_in, y),
    "conv_in_wgrad": lambda: ops.conv_in_wgrad(x, dy, dw_in, db_in, accumulate=True),
    "conv_out_fwd": lambda: ops.conv_out_fwd(h, w_out, b_out, e),
    "conv_out_bwd(dgrad+wgrad)": lambda: ops.conv_out_bwd(h, w_out, de, dh, dw_out, db_out, accumulate=True),
}
for name, fn in cases.items():
    for _ in range(3):
        fn()
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    print(f"{name:28s} median {ts[len(ts)//2]:7.1f} us  min {ts[0]:7.1f} us")
