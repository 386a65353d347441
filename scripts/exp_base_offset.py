"""Bring-up experiment: does a UMMA SWIZZLE_128B descriptor whose start is shifted by j 128-byte rows need
base_offset=j (pattern-relative swizzle) or 0 (absolute-address swizzle)?  Run once per (shift, boff) via env."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from baddiffusion_b200 import _lib, ops
_lib.lib()
torch.manual_seed(0)
M, K, N = 512, 128, 128
x = torch.randn(1, 1, M, K, device="cuda").half()
w = (torch.randn(1, N, K, device="cuda") / 11).half()
y = torch.empty(1, 1, M, N, dtype=torch.float32, device="cuda")
ops.conv_fwd(x, w, y, ksize=1, impl=_lib.BD_IMPL_UMMA)
torch.cuda.synchronize()
ref = x.float().view(M, K) @ w.float().view(N, K).t()
got = y.view(M, N)
err = (got - ref).abs()
rows_ok = (err.max(dim=1).values < 1e-2)
per_tile = [int(rows_ok[i * 128:(i + 1) * 128].sum()) for i in range(M // 128)]
first_bad = [int((~rows_ok[i * 128:(i + 1) * 128]).nonzero()[0]) if (~rows_ok[i*128:(i+1)*128]).any() else -1 for i in range(M // 128)]
print(f"shift={os.environ.get('BD_UMMA_DBG_SHIFT','0')} boff={os.environ.get('BD_UMMA_DBG_BOFF','0')} ok_rows_per_tile={per_tile} first_bad={first_bad} umma_err={_lib.lib().bd_umma_error()}")
