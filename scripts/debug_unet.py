"""Layer-by-layer comparison of the CUDA engine with the CPU oracle (debug aid; GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from oracle import torch_ref as O
from baddiffusion_b200 import _lib
from baddiffusion_b200.unet import UNet2DModel
from baddiffusion_b200.engine import UNetEngine

cfgname = sys.argv[1] if len(sys.argv) > 1 else "cifar10"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
impl = {"simt": _lib.BD_IMPL_SIMT, "auto": _lib.BD_IMPL_AUTO}[sys.argv[3] if len(sys.argv) > 3 else "auto"]
cfg = {"tiny": O.TINY_CONFIG, "cifar10": O.CIFAR10_CONFIG}[cfgname]
sd = O.make_state_dict(cfg, 0)
m = UNet2DModel(**cfg); m.load_state_dict(sd); m = m.cuda()
S = cfg["sample_size"]
x = torch.randn(B, 3, S, S, generator=torch.Generator().manual_seed(5))
t = torch.randint(0, 1000, (B,), generator=torch.Generator().manual_seed(6))

# oracle with recording
rec = {}
topo = O.unet_topology(cfg); g, eps, hd = cfg["norm_num_groups"], cfg["norm_eps"], cfg["attention_head_dim"]
with torch.no_grad():
    temb = O.timestep_embedding(t, cfg["block_out_channels"][0], cfg["flip_sin_to_cos"], cfg["freq_shift"])
    emb = F.linear(temb, sd["time_embedding.linear_1.weight"], sd["time_embedding.linear_1.bias"])
    emb = F.linear(F.silu(emb), sd["time_embedding.linear_2.weight"], sd["time_embedding.linear_2.bias"])
    rec["emb"] = emb
    h = F.conv2d(x, sd["conv_in.weight"], sd["conv_in.bias"], padding=1); rec["conv_in."] = h
    skips = [h]
    for i, b in enumerate(topo["down"]):
        for j in range(len(b["resnets"])):
            p = f"down_blocks.{i}.resnets.{j}."; h = O.resnet_block(sd, p, h, emb, g, eps); rec[p] = h
            if b["attn"]:
                p = f"down_blocks.{i}.attentions.{j}."; h = O.attention_block(sd, p, h, g, eps, hd); rec[p] = h
            skips.append(h)
        if b["down"]:
            p = f"down_blocks.{i}.downsamplers.0."; h = O.downsample(sd, p, h, cfg["downsample_padding"]); rec[p + "conv."] = h; skips.append(h)
    p = "mid_block.resnets.0."; h = O.resnet_block(sd, p, h, emb, g, eps); rec[p] = h
    p = "mid_block.attentions.0."; h = O.attention_block(sd, p, h, g, eps, hd); rec[p] = h
    p = "mid_block.resnets.1."; h = O.resnet_block(sd, p, h, emb, g, eps); rec[p] = h
    for i, b in enumerate(topo["up"]):
        for j in range(len(b["resnets"])):
            h = torch.cat([h, skips.pop()], 1)
            p = f"up_blocks.{i}.resnets.{j}."; h = O.resnet_block(sd, p, h, emb, g, eps); rec[p] = h
            if b["attn"]:
                p = f"up_blocks.{i}.attentions.{j}."; h = O.attention_block(sd, p, h, g, eps, hd); rec[p] = h
        if b["up"]:
            p = f"up_blocks.{i}.upsamplers.0."; h = O.upsample(sd, p, h); rec[p + "conv."] = h
    ref = O.unet_forward(sd, cfg, x, t)

eng = UNetEngine(m, B, True, impl=impl)   # train=True: no buffer pooling, every intermediate is kept
out = eng.forward(x.cuda(), t.cuda())
torch.cuda.synchronize()
print("umma_error", _lib.lib().bd_umma_error())
print("emb", float((eng.emb.cpu() - rec["emb"]).abs().max()))
for p, a in eng.named.items():
    if p not in rec:
        print("missing in oracle:", p); continue
    got = a.t.float().permute(0, 3, 1, 2).cpu()
    r = rec[p]
    err = (got - r).abs().max().item()
    print(f"{p:45s} shape {tuple(r.shape)} max|ref| {r.abs().max().item():8.3f} maxerr {err:9.3e} nan {bool(torch.isnan(got).any())}")
print("eps_hat mse", float(((out.cpu() - ref) ** 2).mean()), "ref std", float(ref.std()))
