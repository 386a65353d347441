"""Times the 3x3 weight-gradient kernels (wgrad3 vs per-tap tile kernel)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from baddiffusion_b200 import _lib, ops
_lib.lib()
B, H, Cin, Cout = (int(v) for v in (sys.argv[1:5] if len(sys.argv) > 4 else (128, 32, 128, 128)))
reps = 20
x = torch.randn(B, H, H, Cin, device="cuda").half()
dy = torch.randn(B, H, H, Cout, device="cuda").half()
dw = torch.zeros(9, Cout, Cin, device="cuda")
db = torch.zeros(Cout, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name, impl in (("wgrad3", _lib.BD_IMPL_UMMA), ("tile", _lib.BD_IMPL_UMMA_TILE)):
    for _ in range(2):
        ops.conv_wgrad(x, dy, dw, None, ksize=3, accumulate=True, impl=impl)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        ops.conv_wgrad(x, dy, dw, None, ksize=3, accumulate=True, impl=impl)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    fl = 2.0 * B * H * H * Cin * Cout * 9
    print(f"{name}: {ms*1e3:.1f} us  {fl/ms/1e9:.1f} TFLOP/s (B={B} H={H} {Cin}->{Cout})")
