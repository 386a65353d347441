"""Execution plan of the UNet on libb200bd kernels: buffers + an ordered list of kernel launches for the forward
pass (D/models/unet_2d.py:229-326 and the blocks it drives) and, in training mode, the hand-derived backward
pass (the reference relies on autograd; SURVEY.md section 7 "Backward is not written anywhere").

Layout decisions:
  * activations fp16 NHWC; every `torch.cat([h, skip], 1)` of the up path (unet_2d_blocks.py:1726,1924) is a
    pre-allocated concat buffer whose two channel slices are written directly by their producers;
  * GroupNorm statistics, softmax, the timestep path and all accumulations are fp32;
  * gradients of activations are fp16 scaled by the loss scale, gradients of parameters fp32 in the flat buffer.
The plan is static for a given (model config, batch, train) so it can be replayed from a CUDA graph.
"""
from __future__ import annotations

import math
import os
from typing import Callable, Dict, List, Optional

import torch

from . import _lib as L
from . import ops
from .unet import resnet_prefixes, topology


class Act:
    """An activation view and (training) the matching gradient view; g_filled is build-time bookkeeping: whether
    some consumer's backward has already written the gradient (later contributions must add).  gs is a (B, C) fp32
    view that the GroupNorm backward fills with the per-sample channel sums of the FINAL gradient (gs_valid tracks, in
    backward order, whether the last writer of g was such a GroupNorm): the producer's bias gradient without a pass
    over g."""

    __slots__ = ("t", "g", "g_filled", "gs", "gs_valid", "sums", "sums_ok")

    def __init__(self, t, g=None, gs=None, sums=None):
        self.t, self.g, self.g_filled, self.gs, self.gs_valid = t, g, False, gs, False
        # sums: (B, C, 2) fp32 view that the producing conv's epilogue fills with the per-(sample, channel) sum / sum of
        # squares of t (forward GroupNorm statistics without a reduction pass); sums_ok: plan-time flag, every producer of
        # t does so
        self.sums, self.sums_ok = sums, False

    @property
    def C(self):
        return self.t.shape[3]

    @property
    def H(self):
        return self.t.shape[1]


class UNetEngine:
    def __init__(self, model, batch: int, train: bool, impl: int = L.BD_IMPL_AUTO):
        self.model, self.B, self.train, self.impl = model, batch, train, impl
        cfg = model.config
        self.cfg = cfg
        self.dev = model.device
        self.lay = model.layout
        self.topo = topology(cfg)
        S = cfg.sample_size if isinstance(cfg.sample_size, int) else cfg.sample_size[0]
        self.S = S
        self.G, self.eps = cfg.norm_num_groups, cfg.norm_eps
        self.fwd: List[Callable[[], None]] = []
        self.bwd: List[Callable[[], None]] = []
        self._bwd_emitters: List[Callable[[], None]] = []
        self._pool: Dict[tuple, torch.Tensor] = {}
        self.io: Dict[str, torch.Tensor] = {}
        self._bias_jobs: List[tuple] = []   # (gs view, bias-gradient view): one batched launch at the end of backward
        self._gn_parts: Dict[int, torch.Tensor] = {}
        self.named: Dict[str, Act] = {}  # layer prefix -> output (introspection / parity debugging)
        self.flat16 = model.flat_half()
        self.flat32 = model.flat_params
        self.gflat = model.flat_grads(attach=False) if train else None
        if train and (cfg.mid_block_scale_factor != 1):
            raise NotImplementedError("training with mid_block_scale_factor != 1 is not implemented")
        maxC = max(max(cfg.block_out_channels) * 2, 64)
        self.gn_work = torch.empty(ops.gn_workspace_floats(batch, maxC), device=self.dev)
        # Parameter gradients are off the critical path of backward (nothing downstream reads them before the
        # optimizer): they run on a side stream, forked after the activation gradient they consume is ready and
        # joined before the timestep-MLP backward.  The small layers (8x8, 4x4) fill a fraction of the SMs, so the
        # two streams genuinely overlap; inside a CUDA graph the fork/join events become plain dependency edges.
        self.side = (torch.cuda.Stream(device=self.dev)
                     if train and not os.environ.get("BD_NO_SIDE_STREAM") else None)
        # GroupNorm statistics accumulated by the producing convs (bd_conv_args.gn_sums): one arena, zeroed by ONE memset
        # at the start of every forward
        self._sums_cap = 2 * batch * 49152 if not os.environ.get("BD_NO_GN_SUMS") else 0
        self._sums_arena = torch.zeros(max(self._sums_cap, 1), device=self.dev)
        self._sums_used = 0
        self.gn_sums_layers = 0   # GroupNorms planned on the producer-statistics path (introspection / tests)
        self._build()
        self._sums_live = self._sums_arena[: self._sums_used]
        if self._sums_used:
            self.fwd.insert(0, self._sums_live.zero_)

    # ------------------------------------------------------------------ parameter views
    def W16(self, key):
        return self.lay.packed(self.flat16, key)

    def P32(self, key):
        return self.lay.packed(self.flat32, key)

    def G32(self, key):
        return self.lay.packed(self.gflat, key)

    # ------------------------------------------------------------------ buffers
    def new(self, H, C, dtype=torch.float16, W=None):
        return torch.empty(self.B, H, H if W is None else W, C, dtype=dtype, device=self.dev)

    def tmp(self, H, C, tag, dtype=torch.float16):
        """forward temporary: kept per layer in training (needed by backward), pooled by shape in inference."""
        if self.train:
            return self.new(H, C, dtype)
        key = ("f", H, C, tag, dtype)
        if key not in self._pool:
            self._pool[key] = self.new(H, C, dtype)
        return self._pool[key]

    def scratch(self, H, C, tag, dtype=torch.float16):
        key = ("s", H, C, tag, dtype)
        if key not in self._pool:
            self._pool[key] = self.new(H, C, dtype)
        return self._pool[key]

    def new_sums(self, H, C):
        """(B, C, 2) slice of the statistics arena, or None where no producer could fill it."""
        n = self.B * C * 2
        if (H * H) % 8 or self._sums_used + n > self._sums_cap:
            return None
        v = self._sums_arena[self._sums_used: self._sums_used + n].view(self.B, C, 2)
        self._sums_used += n
        return v

    def act(self, H, C, tag="x") -> Act:
        a = Act(self.tmp(H, C, tag), sums=self.new_sums(H, C))
        if self.train:
            a.g = self.new(H, C)
            a.gs = torch.empty(self.B, C, device=self.dev)
        return a

    def _conv3_sums(self, x_t, w, y_t, sums, residual=None, x2=None, w2=None, ksize=3, mode=L.BD_CONV_S1, pad=0):
        """Plan time: `sums` if the conv x_t -> y_t will run on a kernel whose epilogue accumulates them, else None."""
        if sums is None:
            return None
        ok = ops.conv_fwd_gn_sums_supported(x_t, w, y_t, ksize=ksize, mode=mode, pad=pad, residual=residual, x2=x2, w2=w2,
                                            impl=self.impl)
        return sums if ok else None

    def _gn_fwd(self, x_t, y_t, gamma, beta, stats, silu, sums):
        """GroupNorm (+SiLU) forward: streaming apply over producer-accumulated statistics when `sums` is given, else the
        single-launch cluster kernel that reduces and applies."""
        if sums is not None:
            ops.groupnorm_apply_sums(x_t, y_t, gamma, beta, sums, stats, self.G, self.eps, silu)
        else:
            ops.groupnorm_fwd(x_t, y_t, gamma, beta, stats, self.gn_work, self.G, self.eps, silu)

    def _gn_bwd(self, x, dy, dx, gamma, beta, stats, dgamma, dbeta, silu, add_dx=None, gsum=None, add_dx2=None):
        """GroupNorm backward whose dgamma / dbeta leave as per-sample partials (B, 2C) and are summed over the batch by
        the batched column-sum launch at the end of backward (no atomics: 128 samples x 2C same-address updates per
        layer otherwise)."""
        if os.environ.get("BD_GN_ATOMIC_DGB"):
            ops.groupnorm_bwd(x, dy, dx, gamma, beta, stats, dgamma, dbeta, self.gn_work, self.G, silu, add_dx=add_dx, gsum=gsum,
                              add_dx2=add_dx2)
            return
        ops.groupnorm_bwd(x, dy, dx, gamma, beta, stats, dgamma, dbeta, self.gn_work, self.G, silu, add_dx=add_dx, gsum=gsum,
                          parts=self._gn_parts[dgamma.data_ptr()], add_dx2=add_dx2)

    def _reg_gn(self, dgamma, dbeta):
        """Plan-build time: the per-sample {dbeta | dgamma} buffer of one GroupNorm and its two batch-sum jobs."""
        if os.environ.get("BD_GN_ATOMIC_DGB"):
            return
        Cc = dgamma.numel()
        parts = torch.empty(self.B, 2 * Cc, device=self.dev)
        self._gn_parts[dgamma.data_ptr()] = parts
        self._bias_jobs.append((parts[:, :Cc], dbeta))
        self._bias_jobs.append((parts[:, Cc:], dgamma))

    def _bias_from(self, out: Act, *grads):
        """Bias gradients of the layer that produced `out`: from out.gs when the last writer of out.g left the channel
        sums there (returns True: the caller skips its own column-sum pass), else False."""
        if out.gs is None or not out.gs_valid or os.environ.get("BD_NO_GSUM"):
            return False
        for g in grads:
            if g is not None:
                self._bias_jobs.append((out.gs, g))
        return True

    # ------------------------------------------------------------------ side stream (parameter gradients)
    def _fork(self, fn: Callable[[], None]):
        """Run fn on the side stream once everything queued so far on the current stream is done.  Whatever fn reads
        must not be rewritten on the main stream before `_join` (callers give such tensors per-layer storage)."""
        if self.side is None:
            fn()
            return
        ev = torch.cuda.Event()
        ev.record()
        self.side.wait_event(ev)
        with torch.cuda.stream(self.side):
            fn()

    def _join(self):
        if self.side is None:
            return
        ev = torch.cuda.Event()
        ev.record(self.side)
        torch.cuda.current_stream().wait_event(ev)

    def grad_buf(self, H, C, tag):
        """Activation gradient that a forked wgrad reads: per layer when the side stream is on, pooled otherwise."""
        return self.new(H, C) if self.side is not None else self.scratch(H, C, tag)

    # ------------------------------------------------------------------ plan construction
    def _build(self):
        cfg, topo, B, S = self.cfg, self.topo, self.B, self.S
        boc = list(cfg.block_out_channels)
        temb_dim, dim0 = topo["temb"], boc[0]
        dev = self.dev
        f32 = torch.float32
        # ---- timestep path
        self.freqs = ops.temb_freqs(dim0, cfg.freq_shift, dev)
        self.sin = torch.empty(B, dim0, device=dev)
        self.h1 = torch.empty(B, temb_dim, device=dev)
        self.emb = torch.empty(B, temb_dim, device=dev)
        self.semb16 = torch.empty(B, temb_dim, dtype=torch.float16, device=dev)
        ncol = self.lay.tproj_rows
        self.tproj = torch.empty(B, ncol, device=dev)
        wtp16 = self.flat16[self.lay.tproj_w_offset: self.lay.tproj_w_offset + ncol * temb_dim].view(1, ncol, temb_dim)
        btp = self.flat32[self.lay.tproj_b_offset: self.lay.tproj_b_offset + ncol]
        w1, b1 = self.P32("time_embedding.linear_1.weight"), self.P32("time_embedding.linear_1.bias")
        w2, b2 = self.P32("time_embedding.linear_2.weight"), self.P32("time_embedding.linear_2.bias")
        flip = bool(cfg.flip_sin_to_cos)

        def f_temb():
            ops.temb_mlp(self.io["t"], w1, b1, w2, b2, self.emb, self.semb16, self.freqs, self.sin, self.h1, flip=flip)
            self._join()   # the trainer refreshes the fp16 weight shadow on the side stream while the fp32 MLP above runs
            ops.conv_fwd(self.semb16.view(1, 1, B, temb_dim), wtp16, self.tproj.view(1, 1, B, ncol), ksize=1, bias=btp,
                         impl=self.impl)

        self.fwd.append(f_temb)
        if self.train:
            self.d_tproj = torch.zeros(B, ncol, device=dev)
            self._bwd_emitters.append(self._emit_temb_bwd)

        # ---- skip / concat planning
        skip_specs = [(boc[0], S)]
        H = S
        for b in topo["down"]:
            for (_, co) in b["resnets"]:
                skip_specs.append((co, H))
            if b["down"]:
                H //= 2
                skip_specs.append((b["channels"], H))
        Hmid = H
        cats: List[Act] = []
        x_slots: List[Act] = []
        skip_acts: List[Optional[Act]] = [None] * len(skip_specs)
        n = 0
        Hu = Hmid
        for b in topo["up"]:
            for (ri, sk, co) in b["resnets"]:
                k = len(skip_specs) - 1 - n
                assert skip_specs[k] == (sk, Hu), (skip_specs[k], sk, Hu)
                cat = Act(self.new(Hu, ri + sk), self.new(Hu, ri + sk) if self.train else None,
                          torch.empty(B, ri + sk, device=dev) if self.train else None, sums=self.new_sums(Hu, ri + sk))
                cats.append(cat)
                x_slots.append(Act(cat.t[..., :ri], cat.g[..., :ri] if self.train else None,
                                   cat.gs[:, :ri] if self.train else None,
                                   sums=cat.sums[:, :ri] if cat.sums is not None else None))
                skip_acts[k] = Act(cat.t[..., ri:], cat.g[..., ri:] if self.train else None,
                                   cat.gs[:, ri:] if self.train else None,
                                   sums=cat.sums[:, ri:] if cat.sums is not None else None)
                n += 1
            if b["up"]:
                Hu *= 2
        self._cats, self._x_slots, self._skips = cats, x_slots, skip_acts

        # ---- conv_in
        h0 = skip_acts[0]  # (h is re-bound below: the closures must capture h0)
        w_in, b_in = self.P32("conv_in.weight"), self.P32("conv_in.bias")
        self.named["conv_in."] = h0
        h0_sums = h0.sums if (h0.sums is not None and self.impl != L.BD_IMPL_SIMT
                              and ops.conv_in_fwd_gn_sums_supported(cfg.in_channels, S, S, boc[0])) else None
        h0.sums_ok = h0_sums is not None
        self.fwd.append(lambda: ops.conv_in_fwd(self.io["x"], w_in, b_in, h0.t, gn_sums=h0_sums))
        if self.train:
            gw_in, gb_in = self.G32("conv_in.weight"), self.G32("conv_in.bias")
            self._bwd_emitters.append(lambda: self.bwd.append(
                # a parameter gradient: side stream, next to the timestep-path tail instead of in front of it
                lambda: self._fork(lambda: ops.conv_in_wgrad(self.io["x"], h0.g, gw_in, gb_in, accumulate=True))))
        h = h0

        # ---- down path
        k = 1
        H = S
        self._n_emitters_before_down1 = None
        for i, b in enumerate(topo["down"]):
            if i == 1:
                self._n_emitters_before_down1 = len(self._bwd_emitters)
            for j in range(len(b["resnets"])):
                dest = skip_acts[k]
                if b["attn"]:
                    mid = self.act(H, dest.C, "dr")
                    self._resnet(f"down_blocks.{i}.resnets.{j}.", h, mid)
                    self._attention(f"down_blocks.{i}.attentions.{j}.", mid, dest)
                else:
                    self._resnet(f"down_blocks.{i}.resnets.{j}.", h, dest)
                h = dest
                k += 1
            if b["down"]:
                dest = skip_acts[k]
                self._downsample(f"down_blocks.{i}.downsamplers.0.conv.", h, dest)
                h = dest
                H //= 2
                k += 1
        # ---- mid
        ms = 1.0 / float(cfg.mid_block_scale_factor)
        m1 = self.act(H, topo["mid"], "m1")
        self._resnet("mid_block.resnets.0.", h, m1, scale=ms)
        h = m1
        if cfg.get("add_attention", True):
            m2 = self.act(H, topo["mid"], "m2")
            self._attention("mid_block.attentions.0.", h, m2, scale=ms)
            h = m2
        self._resnet("mid_block.resnets.1.", h, x_slots[0], scale=ms)
        # ---- up path
        self._n_emitters_before_up = len(self._bwd_emitters)
        n = 0
        for i, b in enumerate(topo["up"]):
            nres = len(b["resnets"])
            for j in range(nres):
                last = j == nres - 1
                if not last:
                    dest = x_slots[n + 1]
                elif b["up"]:
                    dest = self.act(H, b["channels"], "ub")
                else:
                    dest = self.act(H, b["channels"], "fin")
                if b["attn"]:
                    mid = self.act(H, b["channels"], "ur")
                    self._resnet(f"up_blocks.{i}.resnets.{j}.", cats[n], mid, cat_children=(x_slots[n], self._skip_of(n)))
                    self._attention(f"up_blocks.{i}.attentions.{j}.", mid, dest)
                else:
                    self._resnet(f"up_blocks.{i}.resnets.{j}.", cats[n], dest, cat_children=(x_slots[n], self._skip_of(n)))
                h = dest
                n += 1
            if b["up"]:
                self._upsample(f"up_blocks.{i}.upsamplers.0.conv.", h, x_slots[n])
                H *= 2
        # ---- out
        self._conv_out(h)
        # ---- backward list (reverse layer order)
        if self.train:
            for emit in reversed(self._bwd_emitters):
                emit()
            # Backward in three parts for the data-parallel trainer: every emitter appended exactly one op, so part
            # boundaries are emitter counts.  After part i the GEMM weight gradients of the layers it covered are final
            # (`bwd_parts[i][2]`, contiguous ranges of the flat buffer) and can be all-reduced while the next part runs:
            #   0: conv_out + up path        1: mid block + down blocks 1..      2: down block 0, conv_in, timestep path
            assert len(self.bwd) == len(self._bwd_emitters)
            ne = len(self._bwd_emitters)
            k0 = ne - self._n_emitters_before_up
            k1 = ne - (self._n_emitters_before_down1 if self._n_emitters_before_down1 is not None else 0)

            def ranges_of(pred):
                segs = sorted((self.lay.offset[k], self.lay.offset[k] + math.prod(self.lay.entries[k])) for k in self.lay.offset
                              if pred(k) and self.lay.offset[k] < self.lay.tproj_w_offset and self.lay.offset[k] + math.prod(self.lay.entries[k]) <= self.lay.tproj_w_offset)
                out = []
                for lo, hi in segs:
                    if out and lo - out[-1][1] < self.lay.ALIGN:   # alignment padding only
                        out[-1][1] = hi
                    else:
                        out.append([lo, hi])
                for lo, hi in out:   # nothing else lives inside a range
                    for k, o in self.lay.offset.items():
                        assert pred(k) or not (lo <= o < hi), (k, lo, hi)
                return [tuple(r) for r in out]

            nd = len(self.topo["down"])
            late = tuple(f"down_blocks.{i}." for i in range(1, nd)) + ("mid_block.",)
            self.bwd_parts = [(0, k0, ranges_of(lambda k: k.startswith("up_blocks."))),
                              (k0, k1, ranges_of(lambda k: k.startswith(late))),
                              (k1, None, [])]
            if self._bias_jobs:
                rows = [[gs.data_ptr(), gs.stride(0), g.data_ptr(), g.numel()] for gs, g in self._bias_jobs]
                self._bias_table = torch.tensor(rows, dtype=torch.int64, device=dev)
                nj, maxc = len(rows), max(r[3] for r in rows)
                # parameter gradients: with the other parameter-gradient kernels, before the final join
                self.bwd.append(lambda: self._fork(lambda: ops.bias_from_gsum(self._bias_table, nj, maxc, B)))

    def _skip_of(self, n):
        return self._skips[len(self._skips) - 1 - n]

    # ------------------------------------------------------------------ layers
    def _resnet(self, p: str, x: Act, out: Act, scale: float = 1.0, cat_children=None):
        """D/models/resnet.py:551-601."""
        Cin, Cout, H = x.C, out.C, x.H
        G, eps, impl = self.G, self.eps, self.impl
        self.named[p] = out
        has_sc = (p + "conv_shortcut.weight") in self.lay.entries
        a1 = self.tmp(H, Cin, "a1")
        h1 = self.tmp(H, Cout, "h1")
        a2 = self.tmp(H, Cout, "a2")
        st1 = torch.empty(self.B, G, 2, device=self.dev)
        st2 = torch.empty(self.B, G, 2, device=self.dev)
        col = self.lay.tproj_col[p]
        rowb = self.tproj[:, col: col + Cout]
        n1w, n1b, n2w, n2b = (self.P32(p + s) for s in ("norm1.weight", "norm1.bias", "norm2.weight", "norm2.bias"))
        w1, b1, w2, b2 = self.W16(p + "conv1.weight"), self.P32(p + "conv1.bias"), self.W16(p + "conv2.weight"), self.P32(p + "conv2.bias")
        ws = self.W16(p + "conv_shortcut.weight") if has_sc else None
        bs = self.P32(p + "conv_shortcut.bias") if has_sc else None
        # forward GroupNorm statistics from the producers' epilogues where every producer can deliver them
        x_ok = all(c.sums_ok for c in cat_children) if cat_children else x.sums_ok
        x_sums = x.sums if (x_ok and x.sums is not None) else None
        h1_sums = self._conv3_sums(a1, w1, h1, self.new_sums(H, Cout))
        if has_sc:
            out_sums = self._conv3_sums(a2, w2, out.t, out.sums, x2=x.t, w2=ws)
        else:
            out_sums = self._conv3_sums(a2, w2, out.t, out.sums, residual=x.t)
        out.sums_ok = out_sums is not None
        self.gn_sums_layers += (x_sums is not None) + (h1_sums is not None)

        def f():
            self._gn_fwd(x.t, a1, n1w, n1b, st1, True, x_sums)
            ops.conv_fwd(a1, w1, h1, ksize=3, bias=b1, rowbias=rowb, impl=impl, gn_sums=h1_sums)
            self._gn_fwd(h1, a2, n2w, n2b, st2, True, h1_sums)
            if has_sc:
                ops.conv_fwd(a2, w2, out.t, ksize=3, bias=b2, bias2=bs, x2=x.t, w2=ws, scale=scale, impl=impl, gn_sums=out_sums)
            else:
                ops.conv_fwd(a2, w2, out.t, ksize=3, bias=b2, residual=x.t, scale=scale, impl=impl, gn_sums=out_sums)

        self.fwd.append(f)
        if not self.train:
            return

        def emit():
            g = {s: self.G32(p + s) for s in ("norm1.weight", "norm1.bias", "conv1.weight", "conv1.bias", "norm2.weight",
                                              "norm2.bias", "conv2.weight", "conv2.bias")}
            self._reg_gn(g["norm1.weight"], g["norm1.bias"])
            self._reg_gn(g["norm2.weight"], g["norm2.bias"])
            gws = self.G32(p + "conv_shortcut.weight") if has_sc else None
            gbs = self.G32(p + "conv_shortcut.bias") if has_sc else None
            d_a2 = self.scratch(H, Cout, "da2")
            d_h1 = self.grad_buf(H, Cout, "dh1")
            d_a1 = self.scratch(H, Cin, "da1")
            dcol = self.d_tproj[:, col: col + Cout]
            x_filled = x.g_filled
            dout = out.g
            # conv2.bias / conv_shortcut.bias: channel sums of dout, already left in out.gs by dout's last writer
            bias_done = self._bias_from(out, g["conv2.bias"], gbs)
            gb2 = None if bias_done else g["conv2.bias"]
            gbs_ = None if bias_done else gbs
            xgs = x.gs

            def wg2():
                # conv2 (+ shortcut) parameter gradients
                ops.conv_wgrad(a2, dout, g["conv2.weight"], gb2, ksize=3, accumulate=True, impl=impl)
                if has_sc:
                    ops.conv_wgrad(x.t, dout, gws, gbs_, ksize=1, accumulate=True, impl=impl)

            def wg1():
                # d(conv1.bias) == d(time_emb_proj.bias) == column sums of d_tproj, added once in _emit_temb_bwd
                ops.conv_wgrad(a1, d_h1, g["conv1.weight"], None, ksize=3, accumulate=True, impl=impl)

            def bw():
                self._fork(wg2)
                ops.conv_dgrad(dout, w2, d_a2, ksize=3, impl=impl)
                # gsum = per-sample channel sums of d_h1 = gradient of the temb projection (resnet.py:577-580 broadcast add)
                self._gn_bwd(h1, d_a2, d_h1, n2w, n2b, st2, g["norm2.weight"], g["norm2.bias"], True, gsum=dcol)
                self._fork(wg1)
                ops.conv_dgrad(d_h1, w1, d_a1, ksize=3, impl=impl)
                # input gradient: residual / shortcut branch + norm1 branch (+ whatever is already there)
                if has_sc:
                    ops.conv_dgrad(dout, ws, x.g, ksize=1, residual=x.g if x_filled else None, impl=impl)
                    self._gn_bwd(x.t, d_a1, x.g, n1w, n1b, st1, g["norm1.weight"], g["norm1.bias"], True, add_dx=x.g, gsum=xgs)
                elif x_filled:
                    # three-way fan-in in the GroupNorm backward itself: what is already in x.g + the residual branch + norm1's
                    self._gn_bwd(x.t, d_a1, x.g, n1w, n1b, st1, g["norm1.weight"], g["norm1.bias"], True, add_dx=x.g, add_dx2=dout,
                                 gsum=xgs)
                else:
                    self._gn_bwd(x.t, d_a1, x.g, n1w, n1b, st1, g["norm1.weight"], g["norm1.bias"], True, add_dx=dout, gsum=xgs)

            self.bwd.append(bw)
            x.g_filled = True
            x.gs_valid = xgs is not None
            if cat_children:
                for c in cat_children:
                    c.g_filled = True
                    c.gs_valid = xgs is not None

        self._bwd_emitters.append(emit)

    def _attention(self, p: str, x: Act, out: Act, scale: float = 1.0):
        """D/models/attention.py:121-174 with fused [q;k;v] projection."""
        C, H, B = x.C, x.H, self.B
        S = H * H
        G, eps, impl = self.G, self.eps, self.impl
        self.named[p] = out
        hd = self.cfg.attention_head_dim
        heads = C // hd if hd is not None else 1
        sm_scale = 1.0 / math.sqrt(C / heads)
        a = self.tmp(H, C, "aa")
        qkv = self.tmp(H, 3 * C, "qkv")
        ao = self.tmp(H, C, "ao")
        probs = (torch.empty(B * heads, S, S, dtype=torch.float16, device=self.dev) if self.train
                 else self._pool.setdefault(("p", B * heads, S), torch.empty(B * heads, S, S, dtype=torch.float16, device=self.dev)))
        wbytes = max(L.load().bd_attention_bwd_workspace_bytes(B, S, C, heads), L.load().bd_attention_fwd_workspace_bytes(B, S, C, heads))
        work = self._pool.setdefault(("aw", wbytes), torch.empty(wbytes, dtype=torch.uint8, device=self.dev))
        st = torch.empty(B, G, 2, device=self.dev)
        gnw, gnb = self.P32(p + "group_norm.weight"), self.P32(p + "group_norm.bias")
        off = self.lay.offset[p + "query.weight"]
        assert self.lay.offset[p + "key.weight"] == off + C * C and self.lay.offset[p + "value.weight"] == off + 2 * C * C
        wqkv = self.flat16[off: off + 3 * C * C].view(1, 3 * C, C)
        boff = self.lay.offset[p + "query.bias"]
        assert self.lay.offset[p + "key.bias"] == boff + C and self.lay.offset[p + "value.bias"] == boff + 2 * C
        bqkv = self.flat32[boff: boff + 3 * C]
        wp, bp = self.W16(p + "proj_attn.weight").view(1, C, C), self.P32(p + "proj_attn.bias")
        gw = self.gn_work

        x_sums = x.sums if (x.sums_ok and x.sums is not None) else None
        self.gn_sums_layers += x_sums is not None
        out_sums = self._conv3_sums(ao, wp, out.t, out.sums, residual=x.t, ksize=1)
        out.sums_ok = out_sums is not None

        def f():
            self._gn_fwd(x.t, a, gnw, gnb, st, False, x_sums)
            ops.conv_fwd(a, wqkv, qkv, ksize=1, bias=bqkv, impl=impl)
            ops.attention_fwd(qkv.view(B, S, 3 * C), probs, ao.view(B, S, C), work, B, S, C, heads, sm_scale, impl=impl)
            ops.conv_fwd(ao, wp, out.t, ksize=1, bias=bp, residual=x.t, scale=scale, impl=impl, gn_sums=out_sums)

        self.fwd.append(f)
        if not self.train:
            return

        def emit():
            g_gnw, g_gnb = self.G32(p + "group_norm.weight"), self.G32(p + "group_norm.bias")
            self._reg_gn(g_gnw, g_gnb)
            g_wqkv = self.gflat[off: off + 3 * C * C].view(1, 3 * C, C)
            g_bqkv = self.gflat[boff: boff + 3 * C]
            g_wp, g_bp = self.G32(p + "proj_attn.weight").view(1, C, C), self.G32(p + "proj_attn.bias")
            d_ao = self.scratch(H, C, "dao")
            d_qkv = self.grad_buf(H, 3 * C, "dqkv")
            d_a = self.scratch(H, C, "daa")
            x_filled = x.g_filled
            dout = out.g
            g_bp_ = None if self._bias_from(out, g_bp) else g_bp
            xgs = x.gs

            def bw():
                self._fork(lambda: ops.conv_wgrad(ao, dout, g_wp, g_bp_, ksize=1, accumulate=True, impl=impl))
                ops.conv_dgrad(dout, wp, d_ao, ksize=1, impl=impl)
                ops.attention_bwd(qkv.view(B, S, 3 * C), probs, d_ao.view(B, S, C), d_qkv.view(B, S, 3 * C), work, B, S, C,
                                  heads, sm_scale, impl=impl)
                self._fork(lambda: ops.conv_wgrad(a, d_qkv, g_wqkv, g_bqkv, ksize=1, accumulate=True, impl=impl))
                ops.conv_dgrad(d_qkv, wqkv, d_a, ksize=1, impl=impl)
                if x_filled:
                    self._gn_bwd(x.t, d_a, x.g, gnw, gnb, st, g_gnw, g_gnb, False, add_dx=x.g, add_dx2=dout, gsum=xgs)
                else:
                    self._gn_bwd(x.t, d_a, x.g, gnw, gnb, st, g_gnw, g_gnb, False, add_dx=dout, gsum=xgs)

            self.bwd.append(bw)
            x.g_filled = True
            x.gs_valid = xgs is not None

        self._bwd_emitters.append(emit)

    def _downsample(self, p: str, x: Act, out: Act):
        """D/models/resnet.py:199-208 (3x3 stride 2; padding=0 -> zero pad right/bottom)."""
        pad = int(self.cfg.downsample_padding)
        if pad not in (0, 1):
            raise NotImplementedError("downsample_padding must be 0 or 1")
        w, b = self.W16(p + "weight"), self.P32(p + "bias")
        self.named[p] = out
        out_sums = self._conv3_sums(x.t, w, out.t, out.sums, mode=L.BD_CONV_S2_PAD01, pad=pad)
        out.sums_ok = out_sums is not None
        self.fwd.append(lambda: ops.conv_fwd(x.t, w, out.t, ksize=3, mode=L.BD_CONV_S2_PAD01, pad=pad, bias=b, gn_sums=out_sums))
        if not self.train:
            return

        def emit():
            gw_, gb_ = self.G32(p + "weight"), self.G32(p + "bias")
            x_filled = x.g_filled
            dout = out.g
            gb__ = None if self._bias_from(out, gb_) else gb_

            def bw():
                self._fork(lambda: ops.conv_wgrad(x.t, dout, gw_, gb__, ksize=3, mode=L.BD_CONV_S2_PAD01, pad=pad,
                                                  accumulate=True))
                ops.conv_dgrad(dout, w, x.g, ksize=3, mode=L.BD_CONV_S2_PAD01, pad=pad, residual=x.g if x_filled else None)

            self.bwd.append(bw)
            x.g_filled = True
            x.gs_valid = False  # last writer is the strided dgrad

        self._bwd_emitters.append(emit)

    def _upsample(self, p: str, x: Act, out: Act):
        """D/models/resnet.py:126-161: nearest x2 then 3x3 conv."""
        H, C, impl = x.H, x.C, self.impl
        u = self.tmp(2 * H, C, "up")
        w, b = self.W16(p + "weight"), self.P32(p + "bias")
        self.named[p] = out

        out_sums = self._conv3_sums(u, w, out.t, out.sums)
        out.sums_ok = out_sums is not None

        def f():
            ops.upsample2x(x.t, u)
            ops.conv_fwd(u, w, out.t, ksize=3, bias=b, impl=impl, gn_sums=out_sums)

        self.fwd.append(f)
        if not self.train:
            return

        def emit():
            gw_, gb_ = self.G32(p + "weight"), self.G32(p + "bias")
            d_u = self.scratch(2 * H, C, "dup")
            dout = out.g
            assert not x.g_filled
            gb__ = None if self._bias_from(out, gb_) else gb_

            def bw():
                self._fork(lambda: ops.conv_wgrad(u, dout, gw_, gb__, ksize=3, accumulate=True, impl=impl))
                ops.conv_dgrad(dout, w, d_u, ksize=3, impl=impl)
                ops.upsample2x_bwd(d_u, x.g)

            self.bwd.append(bw)
            x.g_filled = True
            x.gs_valid = False

        self._bwd_emitters.append(emit)

    def _conv_out(self, x: Act):
        """unet_2d.py:312-314: conv_norm_out -> SiLU -> conv_out (fp32 NCHW eps_hat)."""
        H, C, B = x.H, x.C, self.B
        G, eps = self.G, self.eps
        a = self.tmp(H, C, "ao_")
        st = torch.empty(B, G, 2, device=self.dev)
        nw, nb = self.P32("conv_norm_out.weight"), self.P32("conv_norm_out.bias")
        w, b = self.P32("conv_out.weight"), self.P32("conv_out.bias")
        self.eps_hat = torch.empty(B, self.cfg.out_channels, H, H, device=self.dev)
        gw = self.gn_work

        x_sums = x.sums if (x.sums_ok and x.sums is not None) else None
        self.gn_sums_layers += x_sums is not None

        def f():
            self._gn_fwd(x.t, a, nw, nb, st, True, x_sums)
            ops.conv_out_fwd(a, w, b, self.eps_hat)

        self.fwd.append(f)
        if not self.train:
            return

        def emit():
            g_nw, g_nb = self.G32("conv_norm_out.weight"), self.G32("conv_norm_out.bias")
            self._reg_gn(g_nw, g_nb)
            g_w, g_b = self.G32("conv_out.weight"), self.G32("conv_out.bias")
            d_a = self.scratch(H, C, "dco")
            assert not x.g_filled

            def bw():
                self._fork(lambda: ops.conv_out_bwd(a, w, self.io["d_eps"], None, g_w, g_b, accumulate=True))  # parameter gradients
                ops.conv_out_bwd(a, w, self.io["d_eps"], d_a, None, None, accumulate=True)                        # data gradient
                self._gn_bwd(x.t, d_a, x.g, nw, nb, st, g_nw, g_nb, True, gsum=x.gs)

            self.bwd.append(bw)
            x.g_filled = True
            x.gs_valid = x.gs is not None

        self._bwd_emitters.append(emit)

    def _emit_temb_bwd(self):
        """Backward of time_emb_proj (all resnets at once) and of the TimestepEmbedding MLP, fp32."""
        B, lay = self.B, self.lay
        temb_dim = self.topo["temb"]
        dim0 = self.cfg.block_out_channels[0]
        ncol = lay.tproj_rows
        dev = self.dev
        wtp32 = self.flat32[lay.tproj_w_offset: lay.tproj_w_offset + ncol * temb_dim].view(ncol, temb_dim)
        g_wtp = self.gflat[lay.tproj_w_offset: lay.tproj_w_offset + ncol * temb_dim].view(ncol, temb_dim)
        g_btp = self.gflat[lay.tproj_b_offset: lay.tproj_b_offset + ncol]
        g_c1b = self.gflat[lay.conv1_b_offset: lay.conv1_b_offset + ncol]  # all conv1.bias, same column order
        w2 = self.P32("time_embedding.linear_2.weight")
        g_w1, g_b1 = self.G32("time_embedding.linear_1.weight"), self.G32("time_embedding.linear_1.bias")
        g_w2, g_b2 = self.G32("time_embedding.linear_2.weight"), self.G32("time_embedding.linear_2.bias")
        ones = torch.ones(B, device=dev)
        d_se = torch.empty(B, temb_dim, device=dev)
        d_emb = torch.empty(B, temb_dim, device=dev)
        d_a1 = torch.empty(B, temb_dim, device=dev)
        d_h1 = torch.empty(B, temb_dim, device=dev)
        dt = self.d_tproj

        def bw():
            # d_tproj was filled (overwritten, column block by column block) by the norm2 GroupNorm backwards
            # The parameter gradients of this tail are off the chain d_tproj -> d_emb -> d_h1 that the remaining
            # (data-gradient) GEMMs form: they go to the side stream, so the end of backward is ~half as long.
            def pg_tproj():
                # time_emb_proj: y = silu(emb) @ Wtp^T + b
                ops.sgemm(dt, 1, ncol, self.emb, temb_dim, 1, g_wtp, temb_dim, 1, ncol, temb_dim, B, accumulate=True, act=2)
                ops.sgemm(ones, 0, 1, dt, ncol, 1, g_btp, 0, 1, 1, ncol, B, accumulate=True)
                ops.sgemm(ones, 0, 1, dt, ncol, 1, g_c1b, 0, 1, 1, ncol, B, accumulate=True)

            def pg_lin2():
                # linear_2: emb = silu(h1) @ W2^T + b2
                ops.sgemm(d_emb, 1, temb_dim, self.h1, temb_dim, 1, g_w2, temb_dim, 1, temb_dim, temb_dim, B, accumulate=True, act=2)
                ops.sgemm(ones, 0, 1, d_emb, temb_dim, 1, g_b2, 0, 1, 1, temb_dim, B, accumulate=True)

            self._fork(pg_tproj)
            ops.sgemm(dt, ncol, 1, wtp32, temb_dim, 1, d_se, temb_dim, 1, B, temb_dim, ncol)
            ops.silu_bwd_f32(d_se, self.emb, d_emb)
            self._fork(pg_lin2)
            ops.sgemm(d_emb, temb_dim, 1, w2, temb_dim, 1, d_a1, temb_dim, 1, B, temb_dim, temb_dim)
            ops.silu_bwd_f32(d_a1, self.h1, d_h1)
            # linear_1: h1 = sin @ W1^T + b1
            ops.sgemm(d_h1, 1, temb_dim, self.sin, dim0, 1, g_w1, dim0, 1, temb_dim, dim0, B, accumulate=True)
            ops.sgemm(ones, 0, 1, d_h1, temb_dim, 1, g_b1, 0, 1, 1, temb_dim, B, accumulate=True)

        self.bwd.append(bw)

    # ------------------------------------------------------------------ execution
    def refresh_weights(self):
        self.model.flat_half()

    def run_forward(self):
        for f in self.fwd:
            f()

    def run_backward(self, part: Optional[int] = None):
        """part None: everything; i: the ops of `bwd_parts[i]`.  Each part ends with the side stream joined."""
        if part is None:
            ops_ = self.bwd
        else:
            a, b, _ = self.bwd_parts[part]
            ops_ = self.bwd[a:b]   # b None: through the batched bias / dgamma / dbeta launch appended after the emitters
        for f in ops_:
            f()
        self._join()

    def forward(self, x: torch.Tensor, timesteps: torch.Tensor) -> torch.Tensor:
        """x (B,C,S,S) fp32 NCHW, timesteps (B,) int64 -> eps_hat (B,C,S,S) fp32 (engine-owned buffer)."""
        assert x.shape[0] == self.B and x.dtype == torch.float32 and x.is_contiguous()
        self.refresh_weights()
        self.io["x"], self.io["t"] = x, timesteps.contiguous()
        self.run_forward()
        return self.eps_hat

    def backward(self, d_eps: torch.Tensor, flat_grad: torch.Tensor, loss_scale: float = 1.0):
        """Accumulates parameter gradients of sum(d_eps * eps_hat) into flat_grad (fp32).  Called by the autograd
        bridge: activation gradients travel in fp16, so d_eps is scaled into fp16 range internally."""
        assert self.train
        assert flat_grad.data_ptr() == self.gflat.data_ptr()
        amax = float(d_eps.abs().max())
        if amax == 0.0 or not math.isfinite(amax):
            return
        s = 2.0 ** math.floor(math.log2(1024.0 / amax))
        if s != 1.0:
            # gradients accumulate linearly: scale what is already there, run, scale back
            flat_grad.mul_(s)
            d_eps = d_eps * s
        self.io["d_eps"] = d_eps.contiguous()
        self.run_backward()
        if s != 1.0:
            flat_grad.mul_(1.0 / s)
