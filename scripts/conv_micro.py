"""Launches the dominant 3x3 conv (128->128 @32x32, B=128) with each tcgen05 kernel; used under ncu and for timing."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from baddiffusion_b200 import _lib, ops
_lib.lib()
B, H, Cin, Cout = (int(v) for v in (sys.argv[1:5] if len(sys.argv) > 4 else (128, 32, 128, 128)))
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
x = torch.randn(B, H, H, Cin, device="cuda").half()
w = (torch.randn(9, Cout, Cin, device="cuda") / 34).half()
y = torch.empty(B, H, H, Cout, dtype=torch.half, device="cuda")
bias = torch.zeros(Cout, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name, impl in (("halo", _lib.BD_IMPL_UMMA), ("tile", _lib.BD_IMPL_UMMA_TILE)):
    for _ in range(2):
        ops.conv_fwd(x, w, y, ksize=3, bias=bias, impl=impl)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        ops.conv_fwd(x, w, y, ksize=3, bias=bias, impl=impl)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    fl = 2.0 * B * H * H * Cin * Cout * 9
    print(f"{name}: {ms*1e3:.1f} us  {fl/ms/1e9:.1f} TFLOP/s (B={B} H={H} {Cin}->{Cout})")
