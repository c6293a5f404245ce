"""EDSR-baseline with the reference's EDSR constructor contract, executed by libsrk.

The reference imports `dlib.models.network_edsr_liif.EDSR_LIIF` (select_network.py:39-50) but
that file is absent from its tree; its building blocks live in dlib/models/network_nlsn.py:
`default_conv` :38-41, `ResBlock` :72-93, `Upsampler` :96-128, wiring `NLSN` :325-369.  This
module is the EDSR-baseline those primitives assemble to (head conv, n_resblocks x
[conv-ReLU-conv, *res_scale, +x], conv, global skip, PixelShuffle upsampler, output conv) with
the state_dict names that wiring produces: head.0, body.{i}.body.{0,2}, body.{n}, tail.0.{2k},
tail.1.  LIIF's implicit decoder is out of scope.
"""
from __future__ import annotations

import ctypes as C
import math

import torch
import torch.nn as nn

from . import _lib as L
from . import packing as P
from .network_swinir import GraphedForward


class _Holder(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter holder: the network is executed by EDSR.forward")


class EDSR(GraphedForward, nn.Module):
    def __init__(self, in_chans=1, n_resblocks=16, n_feats=64, scale=4, rgb_range=1.0,
                 res_scale=1.0, **kwargs):
        super().__init__()
        if in_chans != 1:
            raise NotImplementedError("sr_caco_2_b200.EDSR: only in_chans == 1 is built")
        if n_feats % 64 != 0:
            raise NotImplementedError("sr_caco_2_b200.EDSR: n_feats must be a multiple of 64")
        if scale & (scale - 1) != 0 or scale > 16:
            raise NotImplementedError("sr_caco_2_b200.EDSR: scale must be 2^n")
        self.in_chans, self.n_resblocks, self.n_feats = in_chans, n_resblocks, n_feats
        self.scale, self.rgb_range, self.res_scale = scale, float(rgb_range), float(res_scale)
        self.upscale = scale
        self.head = nn.Sequential(nn.Conv2d(in_chans, n_feats, 3, padding=1))
        body = []
        for _ in range(n_resblocks):
            rb = _Holder()
            rb.body = nn.Sequential(nn.Conv2d(n_feats, n_feats, 3, padding=1), nn.ReLU(True),
                                    nn.Conv2d(n_feats, n_feats, 3, padding=1))
            body.append(rb)
        body.append(nn.Conv2d(n_feats, n_feats, 3, padding=1))
        self.body = nn.Sequential(*body)
        ups = []
        for _ in range(int(round(math.log2(scale)))):
            ups += [nn.Conv2d(n_feats, 4 * n_feats, 3, padding=1), nn.PixelShuffle(2)]
        self.tail = nn.Sequential(nn.Sequential(*ups), nn.Conv2d(n_feats, in_chans, 3, padding=1))
        self._plan = self._keep = self._ws = None
        self._graphs = {}
        self.options = 0             # srk.h SRK_OPT_* bits; 0 = product path
        self.register_load_state_dict_post_hook(lambda mod, keys: mod._invalidate())

    def _invalidate(self):
        self._plan = self._keep = None
        self._graphs = {}

    def _apply(self, fn, *a, **k):
        self._invalidate()
        return super()._apply(fn, *a, **k)

    def flush(self):
        self._invalidate()
        self._ws = None

    def _build_plan(self, conv_dtype=L.SRK_FP16):
        keep = []

        def k(t):
            keep.append(t)
            return L.ptr(t)

        Fe = self.n_feats
        plan = L.EDSRPlan()
        plan.in_chans, plan.n_resblocks, plan.n_feats, plan.scale = self.in_chans, self.n_resblocks, Fe, self.scale
        plan.res_scale, plan.rgb_range, plan.Fp = self.res_scale, self.rgb_range, Fe
        w, b = P.pack_conv_in(self.head[0].weight.detach(), self.head[0].bias.detach())
        plan.head_w, plan.head_b = k(w), k(b)
        body = (L.ConvParams * (2 * self.n_resblocks + 1))()
        for i in range(self.n_resblocks):
            for j, idx in enumerate((0, 2)):
                m = self.body[i].body[idx]
                w, b = P.pack_conv3x3(m.weight.detach(), m.bias.detach(), Fe, Fe, conv_dtype)
                body[2 * i + j] = L.ConvParams(k(w), k(b), Fe, Fe)
        m = self.body[self.n_resblocks]
        w, b = P.pack_conv3x3(m.weight.detach(), m.bias.detach(), Fe, Fe, conv_dtype)
        body[2 * self.n_resblocks] = L.ConvParams(k(w), k(b), Fe, Fe)
        plan.body = body
        n_up = 0
        for m in self.tail[0]:
            if isinstance(m, nn.Conv2d):
                w, b = P.pack_conv3x3(m.weight.detach(), m.bias.detach(), Fe, 4 * Fe, conv_dtype, pixel_shuffle_r=2)
                plan.tail_up[n_up] = L.ConvParams(k(w), k(b), Fe, 4 * Fe)
                n_up += 1
        plan.n_tail_up = n_up
        plan.tail_w = k(P.pack_conv_out(self.tail[1].weight.detach()))
        plan.tail_b = float(self.tail[1].bias.detach().float().item())
        if Fe == 64 and self.tail[1].weight.shape[0] == 1:
            # the linear tail (Upsampler convs + PixelShuffles + output conv) as ONE 5x5 conv
            ups = [(m.weight, m.bias) for m in self.tail[0] if isinstance(m, nn.Conv2d)]
            fw, fb, bw, bb, wsc = P.fold_tail(ups, self.tail[1].weight, self.tail[1].bias, self.scale)
            plan.tail_fold = L.TailFold(k(fw), k(fb), k(bw), k(bb), wsc)
        plan.conv_dtype = conv_dtype
        keep.append(body)
        self._plan, self._keep = plan, keep

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        lib = L.load()
        if self.training:
            raise NotImplementedError("sr_caco_2_b200.EDSR implements the evaluation path only; call .eval()")
        L.require_device(x)
        if x.dim() != 4 or x.shape[1] != self.in_chans:
            raise ValueError(f"expected (B,{self.in_chans},h,w), got {tuple(x.shape)}")
        x = x.float().contiguous()
        B, _, h, w = x.shape
        ps = list(self.parameters())
        fp = (ps[0].data_ptr(), sum(p._version for p in ps), len(ps))
        if self._plan is None or fp != getattr(self, "_plan_fp", None):   # in-place parameter updates bump ._version
            self._build_plan()
            self._plan_fp = fp
            self._graphs = {}
        self._plan.options = int(self.options)
        with torch.cuda.device(x.device):
            if self._use_graph():
                return self._forward_graph(lib, x, B, h, w, self.scale)
            need = self._ws_bytes(lib, B, h, w)
            if self._ws is None or self._ws.numel() < need or self._ws.device != x.device:
                self._ws = None
                self._ws = torch.empty(need, dtype=torch.uint8, device=x.device)
            y = torch.empty(B, self.in_chans, h * self.scale, w * self.scale, dtype=torch.float32, device=x.device)
            self._launch(lib, x, y, B, h, w, self._ws)
        return y

    def _ws_bytes(self, lib, B, h, w):
        return lib.srk_edsr_workspace_bytes(C.byref(self._plan), B, h, w)

    def _launch(self, lib, x, y, B, h, w, ws):
        L.check(lib.srk_edsr_forward(C.byref(self._plan), L.ptr(x), L.ptr(y), B, h, w, L.ptr(ws), ws.numel(), L.stream_ptr()))


def EDSR_LIIF(in_chans, n_resblocks, n_feats, scale, rgb_range, local_ensemble=True,
              feat_unfold=True, cell_decode=True):
    """Constructor signature of the reference's missing EDSR_LIIF (select_network.py:42-50),
    resolving to the EDSR-baseline (the LIIF decoder arguments are accepted and ignored)."""
    return EDSR(in_chans=in_chans, n_resblocks=n_resblocks, n_feats=n_feats, scale=scale,
                rgb_range=rgb_range)
