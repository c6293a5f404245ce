"""SwinIR with the reference's constructor and state_dict contract, executed by libsrk.

Drop-in for `dlib.models.network_swinir.SwinIR` (reference dlib/models/network_swinir.py:710-970)
on the evaluation path: same constructor arguments (:747-769), same parameter / buffer names
and shapes (so `load_state_dict(strict=True)` of a reference checkpoint works, SURVEY.md
appendix B), same forward I/O (fp32 (B,1,h,w) in [0,1] -> fp32 (B,1,h*s,w*s), un-clamped).

The sub-modules below only HOLD parameters under the reference's names; none of them has a
forward of its own.  `SwinIR.forward` repacks the weights once (padded 16-bit GEMM operands)
and issues one `srk_swinir_forward` call: hand-written sm_100a kernels, no PyTorch ops on the
hot path, no CPU / eager fallback (a CPU tensor raises).

Built: in_chans == 1, window_size == 8, upsampler in {'pixelshuffle', 'pixelshuffledirect',
'nearest_conv' (X4)}, resi_connection '1conv' / '3conv', eval mode.  Anything else raises
NotImplementedError at construction.  'nearest_conv' runs each nearest x2 + Conv3x3 pair as one
3x3 conv on the low-res grid with the PixelShuffle epilogue (packing.pack_conv_nearest2x); '3conv'
adds a C -> C/4 conv, a 1x1 row GEMM and LeakyReLU(0.2) epilogues in front of the RSTB conv.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import List

import torch
import torch.nn as nn

from . import _lib as L
from . import packing as P

US_PIXEL_SHUFFLE = "pixelshuffle"              # dlib/utils/constants.py:91-93
US_PIXEL_SHUFFLE_DIRECT = "pixelshuffledirect"
US_NEAREST_CONV = "nearest_conv"               # dlib/utils/constants.py:93
R_CONNECTION_1CONV = "1conv"
R_CONNECTION_3CONV = "3conv"                   # dlib/utils/constants.py:96


def _to_2tuple(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v)


class _Holder(nn.Module):
    """Parameter container (no forward)."""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter holder: the network is executed by SwinIR.forward")


def _attention_holder(dim, ws, num_heads):
    m = _Holder()
    m.relative_position_bias_table = nn.Parameter(torch.zeros((2 * ws - 1) ** 2, num_heads))
    nn.init.trunc_normal_(m.relative_position_bias_table, std=.02)
    p = torch.arange(ws * ws)
    dy = p[:, None] // ws - p[None, :] // ws + ws - 1
    dx = p[:, None] % ws - p[None, :] % ws + ws - 1
    m.register_buffer("relative_position_index", (dy * (2 * ws - 1) + dx).long())
    m.qkv = nn.Linear(dim, dim * 3, bias=True)
    m.proj = nn.Linear(dim, dim)
    return m


def _shift_mask(res, ws, shift):
    """Values of the reference's `attn_mask` buffer (network_swinir.py:260-285).  The kernels
    derive the mask arithmetically; the buffer only exists for state_dict compatibility."""
    def lab(n):
        t = torch.zeros(n, dtype=torch.long)
        t[n - ws:n - shift] = 1
        t[n - shift:] = 2
        return t
    H, W = res
    g = lab(H)[:, None] * 3 + lab(W)[None, :]
    g = g.view(H // ws, ws, W // ws, ws).permute(0, 2, 1, 3).reshape(-1, ws * ws)
    d = g[:, None, :] - g[:, :, None]
    return torch.where(d != 0, torch.tensor(-100.0), torch.tensor(0.0))


class GraphedForward:
    """Mixin: the launches of a forward replayed as ONE CUDA graph per (B, h, w).  The workspace, the input and the
    output of a graph are static buffers (the workspace is caller-owned by design, include/srk.h), the tensor maps are
    encoded once at capture.  The first call of a shape runs eagerly (lazy kernel attributes), the second captures;
    the output is copied out of the static buffer.  Subclasses provide _ws_bytes(lib, B, h, w), _launch(lib, x, y, B,
    h, w, ws) and the attributes in_chans / out scale."""
    MAX_GRAPHS = 4

    def _forward_graph(self, lib, x, B, h, w, scale):
        key = (B, h, w, x.device.index, int(self.options))
        g = self._graphs.get(key)
        if g is None:
            if len(self._graphs) >= self.MAX_GRAPHS:
                self._graphs.pop(next(iter(self._graphs)))
            need = self._ws_bytes(lib, B, h, w)
            g = {"ws": torch.empty(need, dtype=torch.uint8, device=x.device), "x": torch.empty_like(x),
                 "y": torch.empty(B, self.in_chans, h * scale, w * scale, dtype=torch.float32, device=x.device),
                 "graph": None, "launches": 0}
            self._graphs[key] = g

        def run():
            self._launch(lib, g["x"], g["y"], B, h, w, g["ws"])
        g["x"].copy_(x)
        if g["graph"] is None and g["launches"] == 0:          # first call of this shape: eager
            n0 = L.launch_count()
            run()
            g["launches"] = L.launch_count() - n0
        elif g["graph"] is None:                                # second call: capture, then replay
            graph = torch.cuda.CUDAGraph()
            n0 = int(lib.srk_launch_count(0))
            with torch.cuda.graph(graph):
                run()
            L.add_graph_launches(n0 - int(lib.srk_launch_count(0)))   # the launches recorded during capture did not execute
            g["graph"] = graph
            graph.replay()
            L.add_graph_launches(g["launches"])
        else:
            g["graph"].replay()
            L.add_graph_launches(g["launches"])
        return g["y"].clone()

    def _use_graph(self):
        return not (int(self.options) & L.OPT_NO_GRAPH) and not L.profiling and not torch.cuda.is_current_stream_capturing()


class SwinIR(GraphedForward, nn.Module):
    def __init__(self, img_size=64, patch_size=1, in_chans=3, embed_dim=96,
                 depths=[6, 6, 6, 6], num_heads=[6, 6, 6, 6], window_size=7, mlp_ratio=4.,
                 qkv_bias=True, qk_scale=None, drop_rate=0., attn_drop_rate=0.,
                 drop_path_rate=0.1, norm_layer=nn.LayerNorm, ape=False, patch_norm=True,
                 use_checkpoint=False, upscale=2, img_range=1., upsampler='',
                 resi_connection=R_CONNECTION_1CONV, **kwargs):
        super().__init__()
        if in_chans != 1:
            raise NotImplementedError("sr_caco_2_b200.SwinIR: only in_chans == 1 (grayscale "
                                      "microscopy patches) is built")
        if window_size != 8:
            raise NotImplementedError("sr_caco_2_b200.SwinIR: only window_size == 8 is built")
        if upsampler not in (US_PIXEL_SHUFFLE, US_PIXEL_SHUFFLE_DIRECT, US_NEAREST_CONV):
            raise NotImplementedError(f"sr_caco_2_b200.SwinIR: upsampler {upsampler!r} not built")
        if upsampler == US_NEAREST_CONV:
            assert upscale == 4, 'only support x4 now.'            # network_swinir.py:877
        if resi_connection not in (R_CONNECTION_1CONV, R_CONNECTION_3CONV):
            raise NotImplementedError(f"sr_caco_2_b200.SwinIR: resi_connection {resi_connection!r} not built")
        if resi_connection == R_CONNECTION_3CONV and embed_dim // 4 > 64:
            raise NotImplementedError("sr_caco_2_b200.SwinIR: '3conv' is built for embed_dim <= 256")
        if patch_size != 1 or ape or not patch_norm or not qkv_bias or qk_scale is not None \
                or norm_layer is not nn.LayerNorm:
            raise NotImplementedError("sr_caco_2_b200.SwinIR: non-default patch/ape/norm options")
        if upsampler == US_PIXEL_SHUFFLE and (upscale & (upscale - 1)) != 0:
            raise NotImplementedError("sr_caco_2_b200.SwinIR: pixelshuffle is built for 2^n scales")
        if upsampler == US_PIXEL_SHUFFLE_DIRECT and upscale > 8:
            raise NotImplementedError("sr_caco_2_b200.SwinIR: direct upsampler is built up to X8")
        if len(depths) != len(num_heads):
            raise ValueError("depths and num_heads must have the same length")
        num_feat = 64
        self.img_range = float(img_range)
        self.mean = torch.zeros(1, 1, 1, 1)
        self.upscale, self.upsampler, self.window_size = upscale, upsampler, window_size
        self.resi_connection = resi_connection
        self.in_chans, self.embed_dim, self.mlp_ratio = in_chans, embed_dim, mlp_ratio
        self.depths, self.num_heads = list(depths), list(num_heads)
        self.num_layers = len(depths)
        res = _to_2tuple(img_size)
        self.patches_resolution = list(res)
        hidden = int(embed_dim * mlp_ratio)
        self.hidden_dim = hidden

        self.conv_first = nn.Conv2d(in_chans, embed_dim, 3, 1, 1)
        self.patch_embed = _Holder()
        self.patch_embed.norm = nn.LayerNorm(embed_dim)
        self.layers = nn.ModuleList()
        self._block_geometry = []        # (window_size, shift) per block, decided HERE from
        for li, depth in enumerate(depths):  # img_size like network_swinir.py:232-236
            rstb = _Holder()
            rstb.residual_group = _Holder()
            rstb.residual_group.blocks = nn.ModuleList()
            for bi in range(depth):
                ws, shift = window_size, (0 if bi % 2 == 0 else window_size // 2)
                if min(res) <= ws:
                    shift, ws = 0, min(res)
                if ws != 8:
                    raise NotImplementedError(
                        f"sr_caco_2_b200.SwinIR: img_size {img_size} makes the reference shrink "
                        f"its window to {ws}; only 8x8 windows are built")
                self._block_geometry.append((ws, shift))
                blk = _Holder()
                blk.norm1 = nn.LayerNorm(embed_dim)
                blk.attn = _attention_holder(embed_dim, ws, num_heads[li])
                blk.norm2 = nn.LayerNorm(embed_dim)
                blk.mlp = _Holder()
                blk.mlp.fc1 = nn.Linear(embed_dim, hidden)
                blk.mlp.fc2 = nn.Linear(hidden, embed_dim)
                blk.register_buffer("attn_mask", _shift_mask(res, ws, shift) if shift > 0 else None)
                rstb.residual_group.blocks.append(blk)
            rstb.conv = self._resi_conv(embed_dim, resi_connection)
            self.layers.append(rstb)
        self.norm = nn.LayerNorm(embed_dim)
        self.conv_after_body = self._resi_conv(embed_dim, resi_connection)
        if upsampler == US_NEAREST_CONV:                          # network_swinir.py:875-886
            self.conv_before_upsample = nn.Sequential(nn.Conv2d(embed_dim, num_feat, 3, 1, 1),
                                                      nn.LeakyReLU(inplace=True))
            self.conv_up1 = nn.Conv2d(num_feat, num_feat, 3, 1, 1)
            self.conv_up2 = nn.Conv2d(num_feat, num_feat, 3, 1, 1)
            self.conv_hr = nn.Conv2d(num_feat, num_feat, 3, 1, 1)
            self.conv_last = nn.Conv2d(num_feat, in_chans, 3, 1, 1)
            self.lrelu = nn.LeakyReLU(negative_slope=0.2, inplace=True)
        elif upsampler == US_PIXEL_SHUFFLE:
            self.conv_before_upsample = nn.Sequential(nn.Conv2d(embed_dim, num_feat, 3, 1, 1),
                                                      nn.LeakyReLU(inplace=True))
            ups = []
            for _ in range(int(round(math.log2(upscale)))):
                ups += [nn.Conv2d(num_feat, 4 * num_feat, 3, 1, 1), nn.PixelShuffle(2)]
            self.upsample = nn.Sequential(*ups)
            self.conv_last = nn.Conv2d(num_feat, in_chans, 3, 1, 1)
        else:
            self.upsample = nn.Sequential(nn.Conv2d(embed_dim, upscale ** 2 * in_chans, 3, 1, 1),
                                          nn.PixelShuffle(upscale))
        for m in self.modules():          # reference init, network_swinir.py:891-898
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=.02)
                nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.LayerNorm):
                nn.init.constant_(m.bias, 0)
                nn.init.constant_(m.weight, 1.0)
        self._plan = None
        self._plan_fp = None
        self._keep = None
        self._ws = None
        self._graphs = {}
        self.options = 0             # srk.h SRK_OPT_* bits (tests switch single fusions off); 0 = product path
        self.register_load_state_dict_post_hook(lambda mod, keys: mod._invalidate())

    @staticmethod
    def _resi_conv(dim, resi_connection):                          # network_swinir.py:543-552, 870-884
        if resi_connection == R_CONNECTION_1CONV:
            return nn.Conv2d(dim, dim, 3, 1, 1)
        return nn.Sequential(nn.Conv2d(dim, dim // 4, 3, 1, 1), nn.LeakyReLU(negative_slope=0.2, inplace=True),
                             nn.Conv2d(dim // 4, dim // 4, 1, 1, 0), nn.LeakyReLU(negative_slope=0.2, inplace=True),
                             nn.Conv2d(dim // 4, dim, 3, 1, 1))

    # ---- packed-weight cache --------------------------------------------------------------
    def _invalidate(self):
        self._plan = None
        self._keep = None
        self._graphs = {}

    def _apply(self, fn, *a, **k):
        self._invalidate()
        return super()._apply(fn, *a, **k)

    def _fingerprint(self):
        """Cheap identity of the parameter values the packed plan was built from: storage address of the first
        parameter + the sum of all tensor version counters (every in-place write increments one)."""
        ps = list(self.parameters())
        return (ps[0].data_ptr(), sum(p._version for p in ps), len(ps))

    def flush(self):  # called by the reference's ModelPlain (model_plain.py:55) when present
        self._invalidate()
        self._ws = None

    def _build_plan(self, linear_dtype=L.SRK_BF16, conv_dtype=L.SRK_FP16):
        dev = self.conv_first.weight.device
        keep: List[torch.Tensor] = []

        def k(t):
            keep.append(t)
            return L.ptr(t)

        Cdim, hid = self.embed_dim, self.hidden_dim
        Cp, hid_p = P.up(Cdim, 64), P.up(hid, 64)
        max_d = max(Cdim // nh for nh in self.num_heads)
        dp = P.up(max_d, 16)
        ao_p = P.up(max(self.num_heads) * dp, 64)
        nq_p = P.up(3 * max(self.num_heads) * dp, 64)
        nblk = sum(self.depths)
        stbs = (L.StbParams * max(nblk, 1))()
        convs = (L.ConvParams * max(self.num_layers, 1))()
        three = self.resi_connection == R_CONNECTION_3CONV
        convs0 = (L.ConvParams * max(self.num_layers, 1))()
        convs1 = (L.ConvParams * max(self.num_layers, 1))()
        depths = (C.c_int * max(self.num_layers, 1))(*self.depths)
        bi_flat = 0
        for li, rstb in enumerate(self.layers):
            nh = self.num_heads[li]
            if Cdim % nh != 0:
                raise ValueError("embed_dim must be divisible by num_heads")
            d = Cdim // nh
            for bi, blk in enumerate(rstb.residual_group.blocks):
                s = stbs[bi_flat]
                s.ln1_g, s.ln1_b = k(blk.norm1.weight.detach().float().contiguous()), k(blk.norm1.bias.detach().float().contiguous())
                s.ln2_g, s.ln2_b = k(blk.norm2.weight.detach().float().contiguous()), k(blk.norm2.bias.detach().float().contiguous())
                wq, bq = P.pack_qkv(blk.attn.qkv.weight.detach(), blk.attn.qkv.bias.detach(), nh, d, dp, nq_p, Cp, linear_dtype)
                s.w_qkv, s.b_qkv = k(wq), k(bq)
                if Cp - Cdim >= 2 and Cdim % 2 == 0:
                    wfb = P.fold_qkv_bias(wq, bq, Cdim)
                    s.w_qkv_fb = k(wfb)
                    if dp == 32 and nh * dp == Cp:
                        s.w_qkv_hm = k(P.pack_qkv_heads(wfb, nh, dp))
                s.w_proj = k(P.pack_proj(blk.attn.proj.weight.detach(), nh, d, dp, Cp, ao_p, linear_dtype))
                s.b_proj = k(P.pad_bias(blk.attn.proj.bias.detach(), Cp))
                s.w_fc1 = k(P.pack_linear(blk.mlp.fc1.weight.detach(), hid_p, Cp, linear_dtype))
                s.b_fc1 = k(P.pad_bias(blk.mlp.fc1.bias.detach(), hid_p))
                s.w_fc2 = k(P.pack_linear(blk.mlp.fc2.weight.detach(), Cp, hid_p, linear_dtype))
                s.b_fc2 = k(P.pad_bias(blk.mlp.fc2.bias.detach(), Cp))
                s.rel_table = k(blk.attn.relative_position_bias_table.detach().float().t().contiguous())
                s.shift = self._block_geometry[bi_flat][1]
                s.num_heads = nh
                bi_flat += 1
            if three:
                c0, c1, c2 = rstb.conv[0], rstb.conv[2], rstb.conv[4]
                w, b = P.pack_conv3x3(c0.weight.detach(), c0.bias.detach(), Cp, 64, conv_dtype)
                convs0[li].w, convs0[li].b, convs0[li].cin_p, convs0[li].n_p = k(w), k(b), Cp, 64
                convs1[li].w = k(P.pack_linear(c1.weight.detach()[:, :, 0, 0], 64, 64, conv_dtype))
                convs1[li].b, convs1[li].cin_p, convs1[li].n_p = k(P.pad_bias(c1.bias.detach(), 64)), 64, 64
                w, b = P.pack_conv3x3(c2.weight.detach(), c2.bias.detach(), 64, Cp, conv_dtype)
                convs[li].w, convs[li].b, convs[li].cin_p, convs[li].n_p = k(w), k(b), 64, Cp
            else:
                w, b = P.pack_conv3x3(rstb.conv.weight.detach(), rstb.conv.bias.detach(), Cp, Cp, conv_dtype)
                convs[li].w, convs[li].b, convs[li].cin_p, convs[li].n_p = k(w), k(b), Cp, Cp
        plan = L.SwinIRPlan()
        plan.upscale, plan.in_chans, plan.window_size = self.upscale, self.in_chans, self.window_size
        plan.embed_dim, plan.hidden_dim, plan.n_layers = Cdim, hid, self.num_layers
        plan.upsampler = {US_PIXEL_SHUFFLE: L.UPSAMPLER_PIXELSHUFFLE, US_NEAREST_CONV: L.UPSAMPLER_NEAREST_CONV,
                          US_PIXEL_SHUFFLE_DIRECT: L.UPSAMPLER_PIXELSHUFFLEDIRECT}[self.upsampler]
        plan.img_range = self.img_range
        plan.Cp, plan.hid_p, plan.dp, plan.ao_p = Cp, hid_p, dp, ao_p
        plan.depths, plan.stbs, plan.rstb_convs = depths, stbs, convs
        w, b = P.pack_conv_in(self.conv_first.weight.detach(), self.conv_first.bias.detach())
        plan.conv_first_w, plan.conv_first_b = k(w), k(b)
        plan.pe_norm_g = k(self.patch_embed.norm.weight.detach().float().contiguous())
        plan.pe_norm_b = k(self.patch_embed.norm.bias.detach().float().contiguous())
        plan.norm_g = k(self.norm.weight.detach().float().contiguous())
        plan.norm_b = k(self.norm.bias.detach().float().contiguous())
        if three:
            c0, c1, c2 = self.conv_after_body[0], self.conv_after_body[2], self.conv_after_body[4]
            w, b = P.pack_conv3x3(c0.weight.detach(), c0.bias.detach(), Cp, 64, conv_dtype)
            plan.cab_c0 = L.ConvParams(k(w), k(b), Cp, 64)
            plan.cab_c1 = L.ConvParams(k(P.pack_linear(c1.weight.detach()[:, :, 0, 0], 64, 64, conv_dtype)),
                                       k(P.pad_bias(c1.bias.detach(), 64)), 64, 64)
            w, b = P.pack_conv3x3(c2.weight.detach(), c2.bias.detach(), 64, Cp, conv_dtype)
            plan.conv_after_body = L.ConvParams(k(w), k(b), 64, Cp)
            plan.resi_3conv, plan.rstb_c0, plan.rstb_c1 = 1, convs0, convs1
        else:
            w, b = P.pack_conv3x3(self.conv_after_body.weight.detach(), self.conv_after_body.bias.detach(), Cp, Cp, conv_dtype)
            plan.conv_after_body = L.ConvParams(k(w), k(b), Cp, Cp)
        if self.upsampler == US_NEAREST_CONV:
            c0 = self.conv_before_upsample[0]
            w, b = P.pack_conv3x3(c0.weight.detach(), c0.bias.detach(), Cp, 64, conv_dtype)
            plan.conv_before_upsample = L.ConvParams(k(w), k(b), Cp, 64)
            for i, m in enumerate((self.conv_up1, self.conv_up2)):
                w, b = P.pack_conv_nearest2x(m.weight.detach(), m.bias.detach(), 64, conv_dtype)
                plan.upsample[i] = L.ConvParams(k(w), k(b), 64, 256)
            plan.n_upsample = 2
            w, b = P.pack_conv3x3(self.conv_hr.weight.detach(), self.conv_hr.bias.detach(), 64, 64, conv_dtype)
            plan.conv_hr = L.ConvParams(k(w), k(b), 64, 64)
            plan.conv_last_w = k(P.pack_conv_out(self.conv_last.weight.detach()))
            plan.conv_last_b = float(self.conv_last.bias.detach().float().item())
        elif self.upsampler == US_PIXEL_SHUFFLE:
            c0 = self.conv_before_upsample[0]
            w, b = P.pack_conv3x3(c0.weight.detach(), c0.bias.detach(), Cp, 64, conv_dtype)
            plan.conv_before_upsample = L.ConvParams(k(w), k(b), Cp, 64)
            n_up = 0
            for m in self.upsample:
                if isinstance(m, nn.Conv2d):
                    w, b = P.pack_conv3x3(m.weight.detach(), m.bias.detach(), 64, 256, conv_dtype, pixel_shuffle_r=2)
                    plan.upsample[n_up] = L.ConvParams(k(w), k(b), 64, 256)
                    n_up += 1
            plan.n_upsample = n_up
            plan.conv_last_w = k(P.pack_conv_out(self.conv_last.weight.detach()))
            plan.conv_last_b = float(self.conv_last.bias.detach().float().item())
            # the whole linear tail (upsample convs + PixelShuffles + conv_last) as ONE 5x5 conv
            ups = [(m.weight, m.bias) for m in self.upsample if isinstance(m, nn.Conv2d)]
            fw, fb, bw, bb, wsc = P.fold_tail(ups, self.conv_last.weight, self.conv_last.bias, self.upscale)
            plan.tail_fold = L.TailFold(k(fw), k(fb), k(bw), k(bb), wsc)
        else:
            m = self.upsample[0]
            w, b = P.pack_conv3x3(m.weight.detach(), m.bias.detach(), Cp, 64, conv_dtype)
            plan.upsample[0] = L.ConvParams(k(w), k(b), Cp, 64)
            plan.n_upsample = 1
        plan.linear_dtype, plan.conv_dtype = linear_dtype, conv_dtype
        plan.options = int(self.options)
        keep += [stbs, convs, convs0, convs1, depths]
        self._plan, self._keep = plan, keep
        assert all(t.device == dev for t in keep if isinstance(t, torch.Tensor))

    # ---- forward --------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        lib = L.load()
        if self.training:
            raise NotImplementedError("sr_caco_2_b200.SwinIR implements the evaluation path only; "
                                      "call .eval() (training is out of scope)")
        L.require_device(x)
        if x.dim() != 4 or x.shape[1] != self.in_chans:
            raise ValueError(f"expected (B,{self.in_chans},h,w), got {tuple(x.shape)}")
        if x.device != self.conv_first.weight.device:
            raise L.SrkError("input and parameters are on different devices")
        x = x.float().contiguous()
        B, _, h, w = x.shape
        fp = self._fingerprint()
        if self._plan is None or fp != self._plan_fp:        # in-place parameter updates (EMA, optimizer steps) bump ._version
            self._build_plan()
            self._plan_fp = fp
            self._graphs = {}
        self._plan.options = int(self.options)
        with torch.cuda.device(x.device):
            if self._use_graph():
                return self._forward_graph(lib, x, B, h, w, self.upscale)
            need = self._ws_bytes(lib, B, h, w)
            if self._ws is None or self._ws.numel() < need or self._ws.device != x.device:
                self._ws = None
                self._ws = torch.empty(need, dtype=torch.uint8, device=x.device)
            y = torch.empty(B, self.in_chans, h * self.upscale, w * self.upscale,
                            dtype=torch.float32, device=x.device)
            self._launch(lib, x, y, B, h, w, self._ws)
        return y

    def _ws_bytes(self, lib, B, h, w):
        return lib.srk_swinir_workspace_bytes(C.byref(self._plan), B, h, w)

    def _launch(self, lib, x, y, B, h, w, ws):
        L.check(lib.srk_swinir_forward(C.byref(self._plan), L.ptr(x), L.ptr(y), B, h, w, L.ptr(ws), ws.numel(), L.stream_ptr()))

