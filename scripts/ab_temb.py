"""Times bd_temb_mlp (Timesteps + TimestepEmbedding, embeddings.py:22-62,155-212) back to back at the bench batch sizes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from baddiffusion_b200 import ops, _lib
_lib.lib()
dim, temb = 128, 512
for B in (128, 256, 4):
    t = torch.randint(0, 1000, (B,), device="cuda")
    w1, b1 = torch.randn(temb, dim, device="cuda") / 11, torch.randn(temb, device="cuda")
    w2, b2 = torch.randn(temb, temb, device="cuda") / 22, torch.randn(temb, device="cuda")
    emb = torch.empty(B, temb, device="cuda"); se = torch.empty(B, temb, dtype=torch.half, device="cuda")
    so, h1 = torch.empty(B, dim, device="cuda"), torch.empty(B, temb, device="cuda")
    fr = ops.temb_freqs(dim, 1.0, "cuda")
    for _ in range(5):
        ops.temb_mlp(t, w1, b1, w2, b2, emb, se, fr, so, h1, flip=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        ops.temb_mlp(t, w1, b1, w2, b2, emb, se, fr, so, h1, flip=False)
    e1.record(); torch.cuda.synchronize()
    print(f"temb_mlp B={B}: {1e3 * e0.elapsed_time(e1) / 50:.1f} us/launch (back to back)")
    # cold: the fp32 master weights were last touched by Adam, ~1 GB earlier -- flush L2 between launches
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for _ in range(20):
        flush.zero_()
        e0.record()
        ops.temb_mlp(t, w1, b1, w2, b2, emb, se, fr, so, h1, flip=False)
        e1.record(); torch.cuda.synchronize()
        ts.append(1e3 * e0.elapsed_time(e1))
    ts.sort()
    print(f"temb_mlp B={B}: median {ts[10]:.1f} us, min {ts[0]:.1f} us (L2 flushed before each launch)")
