"""A/B inside one process: fused attention kernels (umma_attn.cu) vs the three-launch forward / five-launch backward
(BD_NO_ATTN_FUSED=1) at the CIFAR10 UNet's attention shape (B=128, S=256, C=256).  Rotating buffers, back-to-back."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from baddiffusion_b200 import _lib, ops

_lib.lib()
B, S, C = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (128, 256, 256)
scale = 1 / math.sqrt(C)
NB = 4
qkvs = [(torch.randn(B, S, 3 * C, device="cuda") * 0.7).half() for _ in range(NB)]
dos = [torch.randn(B, S, C, device="cuda").half() for _ in range(NB)]
probs = [torch.empty(B, S, S, dtype=torch.half, device="cuda") for _ in range(NB)]
outs = [torch.empty(B, S, C, dtype=torch.half, device="cuda") for _ in range(NB)]
dqkvs = [torch.empty_like(q) for q in qkvs]
work = torch.empty(_lib.load().bd_attention_bwd_workspace_bytes(B, S, C, 1), dtype=torch.uint8, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def run(fn, n=40):
    for i in range(NB):
        fn(i)
    torch.cuda.synchronize()
    e0.record()
    for i in range(n):
        fn(i % NB)
    e1.record()
    torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / n


res = {}
for fused in (False, "v1", True):
    os.environ.pop("BD_NO_ATTN_FUSED", None)
    os.environ.pop("BD_ATTN_V1", None)
    if fused == "v1":
        os.environ["BD_ATTN_V1"] = "1"      # one-tile-per-CTA fused kernels
    elif not fused:
        os.environ["BD_NO_ATTN_FUSED"] = "1"
    l0 = ops.launch_count()
    tf = run(lambda i: ops.attention_fwd(qkvs[i], probs[i], outs[i], work, B, S, C, 1, scale, impl=_lib.BD_IMPL_UMMA))
    l1 = ops.launch_count()
    tb = run(lambda i: ops.attention_bwd(qkvs[i], probs[i], dos[i], dqkvs[i], work, B, S, C, 1, scale, impl=_lib.BD_IMPL_UMMA))
    l2 = ops.launch_count()
    res[fused] = (outs[0].clone(), probs[0].clone(), dqkvs[0].clone())
    print(f"{'fused (per-image)' if fused is True else 'fused v1' if fused else 'unfused'} B={B} S={S} C={C}: fwd {tf:.1f} us ({(l1 - l0) // 44} launches), bwd {tb:.1f} us ({(l2 - l1) // 44} launches), umma_error={_lib.lib().bd_umma_error()}", flush=True)
a, b = res[False], res[True]
print("max |fused - unfused|: out %.3g probs %.3g dqkv %.3g" % tuple(float((x.float() - y.float()).abs().max()) for x, y in zip(a, b)))
