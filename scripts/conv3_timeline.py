"""Per-CTA timeline of the halo conv kernel (bring-up): cycles at A arrival, each tap's B arrival, end of issue,
accumulator ready, end of epilogue."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from baddiffusion_b200 import _lib, ops
_lib.lib()
B, H, Cin, Cout = (int(v) for v in (sys.argv[1:5] if len(sys.argv) > 4 else (128, 32, 128, 128)))
x = torch.randn(B, H, H, Cin, device="cuda").half()
w = (torch.randn(9, Cout, Cin, device="cuda") / 34).half()
y = torch.empty(B, H, H, Cout, dtype=torch.half, device="cuda")
ops.conv_fwd(x, w, y, ksize=3, impl=_lib.BD_IMPL_UMMA)
nct = 4096
dbg = torch.zeros(nct, 64, dtype=torch.int64, device="cuda")
os.environ["BD_CONV3_DBG_PTR"] = str(dbg.data_ptr())
ops.conv_fwd(x, w, y, ksize=3, impl=_lib.BD_IMPL_UMMA)
torch.cuda.synchronize()
d = dbg.cpu()
for cta in (0, 1, 100, 147, 148, 200, 255):
    r = d[cta]
    t0 = int(r[0])
    ev = [int(v) - t0 for v in r[1:40] if int(v) != 0]
    print(f"cta {cta:3d}: events {ev}  accum_ready {int(r[48]) - t0} epi_done {int(r[49]) - t0}")
