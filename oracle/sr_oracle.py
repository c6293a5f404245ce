"""CPU oracle for the SR-CACO-2 evaluation hot path (SwinIR / EDSR forward + PSNR/SSIM/NRMSE).

TEST INFRASTRUCTURE ONLY.  This file is the checker, never the product: only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import
it.  The product path (`sr_caco_2_b200/`) never imports anything under `oracle/` and fails
loudly when its CUDA library is missing.

It is a plain-PyTorch fp32 *restatement* (functional, driven by a reference-layout
state_dict) of the algorithms in the reference, written from the reference's behaviour and
citing the lines it follows (paths relative to /root/reference):

  SwinIR forward ............ dlib/models/network_swinir.py:930-970  (forward)
     reflect pad ............ :908-913   (check_image_size)
     body ................... :915-928   (forward_features), :562-565 (RSTB.forward)
     Swin block ............. :287-337   (SwinTransformerBlock.forward)
     window attention ....... :140-179   (WindowAttention.forward), index :116-129
     shift mask ............. :260-285   (calculate_mask)
     partition / reverse .... :48-80
     MLP .................... :39-45
     upsamplers ............. :662-702   (Upsample / UpsampleOneStep)
  EDSR-baseline ............. dlib/models/network_nlsn.py:38-41 (default_conv), :72-93
                              (ResBlock), :96-128 (Upsampler), wiring :325-369 (NLSN without
                              its attention blocks); defaults utils_init_default_args.py:37-50.
                              NB `dlib/models/network_edsr_liif.py` is ABSENT from the
                              reference tree, so for EDSR the parity is pinned on those
                              in-repo primitives (imported live in tests/golden/make_golden.py).
  metrics ................... dlib/utils/utils_image.py:369-372 (tensor2uint82float),
                              :843-891 (psnr), :894-934 (mse), :937-1007 (nrmse),
                              :1010-1099 + :1102-1117 + :1120-1198 (ssim), :618-653 (ycbcr);
                              caller dlib/utils/utils_trainer.py:961-1035 (_compute_metrics),
                              :874-930 (marginalize_roi_th_perf).

Parity pin: `tests/golden/make_golden.py` ran the unmodified reference modules in the build
container (through `oracle/ref_import.py`) and committed their outputs under `tests/golden/`;
`tests/test_oracle.py` checks this restatement against those vectors (and against the live
reference whenever /root/reference is present).  Parity is therefore PINNED for SwinIR and
the metrics, and pinned-on-primitives for EDSR (see above).

`emulate_bf16=True` additionally applies the operand roundings of the CUDA path's arithmetic
contract: Linear / attention operands rounded to bfloat16, 3x3-conv operands rounded to
float16 (bf16's 8-bit mantissa on the conv operands that feed the skip connections costs
5e-3 max-abs on configs[0]; fp16 keeps it under 1e-3, see DESIGN.md "Numerics"), the 1-channel
input conv in exact fp32; fp32 accumulation,
fp32 residual stream / LayerNorm / softmax / GELU everywhere.  The GPU tests use it to
separate indexing bugs (tight tolerance against the emulation) from the precision budget
(2e-3 against pure fp32).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# ----------------------------------------------------------------------------------------
# configuration records (mirror the reference constructor arguments)
# ----------------------------------------------------------------------------------------
@dataclass
class SwinIRCfg:
    upscale: int = 2
    in_chans: int = 1
    img_size: int = 64
    window_size: int = 8
    img_range: float = 1.0
    depths: List[int] = field(default_factory=lambda: [6, 6, 6, 6, 6, 6])
    embed_dim: int = 180
    num_heads: List[int] = field(default_factory=lambda: [6, 6, 6, 6, 6, 6])
    mlp_ratio: float = 2.0
    upsampler: str = "pixelshuffle"
    resi_connection: str = "1conv"


@dataclass
class EDSRCfg:
    in_chans: int = 1
    n_resblocks: int = 16
    n_feats: int = 64
    scale: int = 4
    rgb_range: float = 1.0
    res_scale: float = 1.0


# ----------------------------------------------------------------------------------------
# index maps (integer work: must be bit exact)
# ----------------------------------------------------------------------------------------
def relative_position_index(ws: int) -> Tensor:
    """(ws*ws, ws*ws) int64, value = (dy + ws-1) * (2ws-1) + (dx + ws-1)
    with (dy, dx) = position(i) - position(j).  network_swinir.py:116-129."""
    p = torch.arange(ws * ws)
    py, px = p // ws, p % ws
    dy = py[:, None] - py[None, :] + ws - 1
    dx = px[:, None] - px[None, :] + ws - 1
    return (dy * (2 * ws - 1) + dx).to(torch.int64)


def _region_label(n: int, ws: int, shift: int) -> Tensor:
    """Label 0/1/2 of every coordinate of the *shifted* frame along one axis:
    [0, n-ws) -> 0, [n-ws, n-shift) -> 1, [n-shift, n) -> 2.  network_swinir.py:264-269."""
    lab = torch.zeros(n, dtype=torch.int64)
    lab[n - ws:n - shift] = 1
    lab[n - shift:] = 2
    return lab


def shift_attention_mask(H: int, W: int, ws: int, shift: int) -> Tensor:
    """(nW, ws*ws, ws*ws) fp32 additive mask {0, -100}.  network_swinir.py:260-285."""
    lab = (_region_label(H, ws, shift)[:, None] * 3 + _region_label(W, ws, shift)[None, :])
    lab = lab.view(H // ws, ws, W // ws, ws).permute(0, 2, 1, 3).reshape(-1, ws * ws)
    diff = lab[:, None, :] - lab[:, :, None]
    return torch.where(diff != 0, torch.tensor(-100.0), torch.tensor(0.0))


def window_gather_map(H: int, W: int, ws: int, shift: int) -> Tensor:
    """int64 (nW*ws*ws,): flat token id (row-major h*W+w of the un-shifted frame) that sits at
    each position of the window-major layout after `roll(-shift)` + `window_partition`.
    network_swinir.py:296-306 and :48-62."""
    hh = (torch.arange(H) + shift) % H
    ww = (torch.arange(W) + shift) % W
    tok = hh[:, None] * W + ww[None, :]
    return tok.view(H // ws, ws, W // ws, ws).permute(0, 2, 1, 3).reshape(-1)


def pixel_shuffle_map(C: int, H: int, W: int, r: int) -> Tensor:
    """For an input of shape (C*r*r, H, W) flattened row-major, the flat input index that
    lands at each flat output index of the (C, H*r, W*r) result.  torch.nn.PixelShuffle as
    used at network_swinir.py:675,701 / network_nlsn.py:108."""
    c = torch.arange(C)[:, None, None]
    oy = torch.arange(H * r)[None, :, None]
    ox = torch.arange(W * r)[None, None, :]
    ch = c * r * r + (oy % r) * r + (ox % r)
    return (ch * H + oy // r) * W + ox // r


# ----------------------------------------------------------------------------------------
# arithmetic helpers
# ----------------------------------------------------------------------------------------
def _r(t: Tensor, emu: bool) -> Tensor:
    """Round a Linear/attention operand to bf16 (and back) when emulating the CUDA path."""
    return t.to(torch.bfloat16).to(torch.float32) if emu else t


def _h(t: Tensor, emu: bool) -> Tensor:
    """Round a conv operand to fp16 (saturating, and back) when emulating the CUDA path."""
    return t.clamp(-65504.0, 65504.0).to(torch.float16).to(torch.float32) if emu else t


def _linear(x, w, b, emu):
    return F.linear(_r(x, emu), _r(w, emu), b)


def _conv3(x, w, b, emu, kind="gemm"):
    """kind: 'gemm' (tensor-core implicit GEMM, fp16 operands), 'in' (1-channel input conv,
    exact fp32 on CUDA cores), 'out' (1-channel output conv: tensor-core direct conv, fp16 operands)."""
    if kind == "in":
        return F.conv2d(x, w, b, stride=1, padding=1)
    if kind == "out":
        return F.conv2d(_h(x, emu), _h(w, emu), b, stride=1, padding=1)
    return F.conv2d(_h(x, emu), _h(w, emu), b, stride=1, padding=1)


def _to_windows(x: Tensor, ws: int) -> Tensor:
    B, H, W, C = x.shape
    x = x.reshape(B, H // ws, ws, W // ws, ws, C).transpose(2, 3)
    return x.reshape(-1, ws * ws, C)


def _from_windows(xw: Tensor, ws: int, B: int, H: int, W: int) -> Tensor:
    C = xw.shape[-1]
    x = xw.reshape(B, H // ws, W // ws, ws, ws, C).transpose(2, 3)
    return x.reshape(B, H, W, C)


def block_geometry(cfg: SwinIRCfg, blk_idx: int):
    """(window_size, shift) a block decides AT CONSTRUCTION from img_size, not from the
    runtime tensor: network_swinir.py:232-236 (and :451 for the alternating shift)."""
    ws = cfg.window_size
    shift = 0 if blk_idx % 2 == 0 else ws // 2
    res = cfg.img_size if isinstance(cfg.img_size, int) else min(cfg.img_size)
    if res <= ws:
        shift, ws = 0, res
    return ws, shift


def _swin_block(x: Tensor, H: int, W: int, p: Dict[str, Tensor], pre: str, nh: int,
                ws: int, shift: int, emu: bool) -> Tensor:
    """One Swin transformer block on (B, H*W, C).  network_swinir.py:287-337."""
    B, L, C = x.shape
    d = C // nh
    y = F.layer_norm(x, (C,), p[pre + "norm1.weight"], p[pre + "norm1.bias"], 1e-5)
    y = y.view(B, H, W, C)
    if shift > 0:
        y = torch.roll(y, (-shift, -shift), (1, 2))
    yw = _to_windows(y, ws)                                            # (B*nW, N, C)
    N = ws * ws
    qkv = _linear(yw, p[pre + "attn.qkv.weight"], p[pre + "attn.qkv.bias"], emu)
    qkv = qkv.view(-1, N, 3, nh, d).permute(2, 0, 3, 1, 4)             # 3, B_, nh, N, d
    q, k, v = qkv[0], qkv[1], qkv[2]
    if emu:
        # CUDA path: q,k,v are stored bf16; the softmax scale is applied to fp32 scores
        att = (_r(q, emu) @ _r(k, emu).transpose(-1, -2)) * (d ** -0.5)
    else:
        att = (q * (d ** -0.5)) @ k.transpose(-1, -2)
    table = p[pre + "attn.relative_position_bias_table"]
    bias = table[relative_position_index(ws).reshape(-1).to(table.device)].view(N, N, nh).permute(2, 0, 1)
    att = att + bias[None]
    if shift > 0:
        m = shift_attention_mask(H, W, ws, shift).to(att.device)       # nW, N, N
        nW = m.shape[0]
        att = (att.view(-1, nW, nh, N, N) + m[None, :, None]).view(-1, nh, N, N)
    att = torch.softmax(att, dim=-1)
    o = (_r(att, emu) @ _r(v, emu)).transpose(1, 2).reshape(-1, N, C)
    o = _linear(o, p[pre + "attn.proj.weight"], p[pre + "attn.proj.bias"], emu)
    o = _from_windows(o, ws, B, H, W)
    if shift > 0:
        o = torch.roll(o, (shift, shift), (1, 2))
    x = x + o.reshape(B, L, C)
    y = F.layer_norm(x, (C,), p[pre + "norm2.weight"], p[pre + "norm2.bias"], 1e-5)
    y = _linear(y, p[pre + "mlp.fc1.weight"], p[pre + "mlp.fc1.bias"], emu)
    y = F.gelu(y)                                                      # exact erf GELU (:30)
    y = _linear(y, p[pre + "mlp.fc2.weight"], p[pre + "mlp.fc2.bias"], emu)
    return x + y


@torch.no_grad()
def swinir_forward(sd: Dict[str, Tensor], cfg: SwinIRCfg, x: Tensor,
                   emulate_bf16: bool = False) -> Tensor:
    """SwinIR forward, (B,C,h,w) fp32 in [0,1] -> (B,C,h*s,w*s).  network_swinir.py:930-970.
    Supported: upsampler in {'pixelshuffle','pixelshuffledirect','nearest_conv'}, resi_connection
    '1conv' / '3conv'."""
    emu = emulate_bf16
    p = {k: v.float() if v.is_floating_point() else v for k, v in sd.items()}
    assert cfg.resi_connection in ("1conv", "3conv")

    def resi_conv(img, name):                                           # :543-552, 870-884
        if cfg.resi_connection == "1conv":
            return _conv3(img, p[name + ".weight"], p[name + ".bias"], emu)
        t1 = F.leaky_relu(_conv3(img, p[name + ".0.weight"], p[name + ".0.bias"], emu), 0.2)
        t2 = F.leaky_relu(F.conv2d(_h(t1, emu), _h(p[name + ".2.weight"], emu), p[name + ".2.bias"]), 0.2)
        return _conv3(t2, p[name + ".4.weight"], p[name + ".4.bias"], emu)
    B, Cin, h0, w0 = x.shape
    wsz = cfg.window_size
    ph, pw = (wsz - h0 % wsz) % wsz, (wsz - w0 % wsz) % wsz
    x = F.pad(x.float(), (0, pw, 0, ph), mode="reflect")               # :908-913
    mean = torch.zeros(1, Cin, 1, 1, device=x.device)
    if Cin == 3:
        mean = torch.tensor([0.4488, 0.4371, 0.4040], device=x.device).view(1, 3, 1, 1)  # :776-779
    x = (x - mean) * cfg.img_range
    H, W = x.shape[2:]
    C = cfg.embed_dim

    f0 = _conv3(x, p["conv_first.weight"], p["conv_first.bias"], emu, "in")   # (B,C,H,W)
    t = f0.flatten(2).transpose(1, 2)                                   # (B,HW,C)  :610
    t = F.layer_norm(t, (C,), p["patch_embed.norm.weight"], p["patch_embed.norm.bias"], 1e-5)
    for li, depth in enumerate(cfg.depths):
        t_in = t
        for bi in range(depth):
            ws, shift = block_geometry(cfg, bi)
            t = _swin_block(t, H, W, p, f"layers.{li}.residual_group.blocks.{bi}.",
                            cfg.num_heads[li], ws, shift, emu)
        img = t.transpose(1, 2).reshape(B, C, H, W)                     # unembed :651-655
        img = resi_conv(img, f"layers.{li}.conv")
        t = img.flatten(2).transpose(1, 2) + t_in                       # :563-565
    t = F.layer_norm(t, (C,), p["norm.weight"], p["norm.bias"], 1e-5)
    img = t.transpose(1, 2).reshape(B, C, H, W)
    y = resi_conv(img, "conv_after_body") + f0

    s = cfg.upscale
    if cfg.upsampler == "pixelshuffle":                                 # :941-942
        y = F.leaky_relu(_conv3(y, p["conv_before_upsample.0.weight"],
                                p["conv_before_upsample.0.bias"], emu), 0.01)
        if s & (s - 1) == 0:
            for i in range(int(round(math.log2(s)))):
                y = F.pixel_shuffle(_conv3(y, p[f"upsample.{2 * i}.weight"],
                                           p[f"upsample.{2 * i}.bias"], emu), 2)
        elif s == 3:
            y = F.pixel_shuffle(_conv3(y, p["upsample.0.weight"], p["upsample.0.bias"], emu), 3)
        else:
            raise ValueError(f"scale {s} is not supported")
        y = _conv3(y, p["conv_last.weight"], p["conv_last.bias"], emu, "out")
    elif cfg.upsampler == "pixelshuffledirect":                         # :947
        y = F.pixel_shuffle(_conv3(y, p["upsample.0.weight"], p["upsample.0.bias"], emu), s)
    elif cfg.upsampler == "nearest_conv":                               # :948-961 (X4)
        assert s == 4
        y = F.leaky_relu(_conv3(y, p["conv_before_upsample.0.weight"],
                                p["conv_before_upsample.0.bias"], emu), 0.01)
        for name in ("conv_up1", "conv_up2"):
            y = F.interpolate(y, scale_factor=2, mode="nearest")
            y = F.leaky_relu(_conv3(y, p[name + ".weight"], p[name + ".bias"], emu), 0.2)
        y = F.leaky_relu(_conv3(y, p["conv_hr.weight"], p["conv_hr.bias"], emu), 0.2)
        y = _conv3(y, p["conv_last.weight"], p["conv_last.bias"], emu, "out")
    else:
        raise NotImplementedError(cfg.upsampler)
    y = y / cfg.img_range + mean
    return y[:, :, :h0 * s, :w0 * s]


@torch.no_grad()
def edsr_forward(sd: Dict[str, Tensor], cfg: EDSRCfg, x: Tensor,
                 emulate_bf16: bool = False) -> Tensor:
    """EDSR-baseline forward assembled from the reference's EDSR primitives:
    head conv -> n x [x + res_scale*conv(relu(conv(x)))] -> conv -> + head -> upsampler -> conv.
    network_nlsn.py:72-93, :96-128, :325-369 (mean shift disabled as at :360,:367)."""
    emu = emulate_bf16
    p = {k: v.float() for k, v in sd.items()}
    x = x.float()
    hfeat = _conv3(x, p["head.0.weight"], p["head.0.bias"], emu, "in")
    r = hfeat
    for i in range(cfg.n_resblocks):
        t = F.relu(_conv3(r, p[f"body.{i}.body.0.weight"], p[f"body.{i}.body.0.bias"], emu))
        t = _conv3(t, p[f"body.{i}.body.2.weight"], p[f"body.{i}.body.2.bias"], emu)
        r = t * cfg.res_scale + r
    n = cfg.n_resblocks
    r = _conv3(r, p[f"body.{n}.weight"], p[f"body.{n}.bias"], emu) + hfeat
    s = cfg.scale
    if s & (s - 1) == 0:
        for i in range(int(round(math.log2(s)))):
            r = F.pixel_shuffle(_conv3(r, p[f"tail.0.{2 * i}.weight"],
                                       p[f"tail.0.{2 * i}.bias"], emu), 2)
    elif s == 3:
        r = F.pixel_shuffle(_conv3(r, p["tail.0.0.weight"], p["tail.0.0.bias"], emu), 3)
    else:
        raise NotImplementedError
    return _conv3(r, p["tail.1.weight"], p["tail.1.bias"], emu, "out")


# ----------------------------------------------------------------------------------------
# metrics
# ----------------------------------------------------------------------------------------
def quantize_u8f(img: Tensor) -> Tensor:
    """[0,1] float -> integer valued float in [0,255]; round-half-even.  utils_image.py:369."""
    return (img.float().clamp(0, 1) * 255.0).round().clamp(0, 255).float()


def _crop(t: Optional[Tensor], b: int) -> Optional[Tensor]:
    if t is None:
        return None
    h, w = t.shape[-2:]
    return t[..., b:h - b, b:w - b]


def _mse64(a: Tensor, b: Tensor, border: int, roi: Optional[Tensor]) -> Tensor:
    a, b, roi = _crop(a, border).double(), _crop(b, border).double(), _crop(roi, border)
    n = a.shape[0]
    if roi is None:
        return ((a - b) ** 2).reshape(n, -1).mean(-1)
    roi = roi.double()
    cnt = roi.reshape(n, -1).sum(-1)
    cnt = torch.where(cnt == 0, torch.ones_like(cnt), cnt)
    return (((a - b) * roi) ** 2).reshape(n, -1).sum(-1) / cnt


def mse(a, b, border=0, roi=None) -> Tensor:
    """utils_image.py:894-934."""
    return _mse64(a, b, border, roi)


def psnr(a, b, border=0, roi=None) -> Tensor:
    """utils_image.py:843-891 (fp64; MSE floor 1e-45)."""
    m = _mse64(a, b, border, roi).clamp_min(1e-45)
    return 20.0 * torch.log10(255.0 / torch.sqrt(m))


def nrmse(a, y, border=0, roi=None) -> Tensor:
    """utils_image.py:937-1007: sqrt(mse) / (max(y) - min(y)) on the cropped target; with a
    ROI the extrema are taken over y*roi (min additionally floored by the global min)."""
    m = _mse64(a, y, border, roi)
    yc = _crop(y, border).double()
    n = yc.shape[0]
    if roi is None:
        flat = yc.reshape(n, -1)
        lo = flat.min(-1)[0]
    else:
        flat = (yc * _crop(roi, border).double()).reshape(n, -1)
        lo = torch.maximum(yc.reshape(n, -1).min(-1)[0], flat.min(-1)[0])
    den = flat.max(-1)[0] - lo
    den = torch.where(den == 0, torch.ones_like(den), den)
    return torch.sqrt(m) / den


def gaussian_window(size: int = 11, sigma: float = 1.5) -> Tensor:
    """2-D normalised Gaussian, fp32.  utils_image.py:1102-1117."""
    c = torch.arange(size, dtype=torch.float32) - (size - 1) / 2.0
    g = torch.exp(-(c[None, :] ** 2 + c[:, None] ** 2) / (2 * sigma ** 2))
    return g / g.sum()


def ssim(a, b, border=0, roi=None) -> Tensor:
    """utils_image.py:1120-1198 + :1010-1099: fp32, /255, 11x11 sigma 1.5 VALID filtering,
    c1=1e-4, c2=9e-4, mean of the map (or ROI weighted with the ROI cropped by 5)."""
    assert a.shape == b.shape and a.dim() == 4
    x, y, roi = _crop(a, border) / 255.0, _crop(b, border) / 255.0, _crop(roi, border)
    ch = x.shape[1]
    k = gaussian_window().repeat(ch, 1, 1, 1).to(x.device)
    if x.shape[-1] < 11 or x.shape[-2] < 11:
        raise ValueError("Kernel size can't be greater than actual input size")
    filt = lambda t: F.conv2d(t, k, groups=ch)
    mx, my = filt(x), filt(y)
    sxx, syy, sxy = filt(x * x) - mx * mx, filt(y * y) - my * my, filt(x * y) - mx * my
    c1, c2 = 0.01 ** 2, 0.03 ** 2
    ss = ((2 * mx * my + c1) / (mx * mx + my * my + c1)) * ((2 * sxy + c2) / (sxx + syy + c2))
    n = ss.shape[0]
    if roi is None:
        per_c = ss.reshape(n, ch, -1).mean(-1)
    else:
        r = _crop(roi, 5)
        cnt = r.reshape(n, -1).sum(-1)
        cnt = torch.where(cnt == 0, torch.ones_like(cnt), cnt)
        per_c = (ss * r).reshape(n, ch, -1).sum(-1) / cnt[:, None]
    return per_c.mean(1)


def luma(v01: Tensor) -> Tensor:
    """Y of a gray value replicated to RGB, utils_image.py:618-653 with float input in [0,1]:
    y = ((65.481+128.553+24.966) * 255v / 255 + 16) / 255 clamped to [0,1]."""
    v = v01.float() * 255.0
    y = (65.481 * v + 128.553 * v + 24.966 * v) / 255.0 + 16.0
    return (y / 255.0).clamp(0.0, 1.0)


def all_metrics(E: Tensor, Hh: Tensor, border: int, roi_th: Optional[int] = None
                ) -> Dict[str, Tensor]:
    """What utils_trainer.py:961-1035 computes per image for 1-channel E, H in [0,1]."""
    assert E.shape[1] == 1
    e8, h8 = quantize_u8f(E), quantize_u8f(Hh)
    roi = None if roi_th is None else (h8 >= roi_th).float()
    ey, hy = luma(e8 / 255.0) * 255.0, luma(h8 / 255.0) * 255.0
    return {
        "psnr": psnr(e8, h8, border, roi),
        "mse": mse(e8, h8, border, roi),
        "nrmse": nrmse(e8, h8, border, roi),
        "ssim": ssim(e8, h8, border, roi),
        "psnr_y": psnr(ey, hy, border, roi),
    }


def roi_marginal_metrics(E: Tensor, Hh: Tensor, border: int,
                         ths=(4, 5, 6, 7, 8, 9, 10)) -> Dict[str, Tensor]:
    """Mean over ROI thresholds, utils_trainer.py:874-930 (thresholds constants.py:817)."""
    acc = None
    for th in ths:
        m = all_metrics(E, Hh, border, th)
        acc = m if acc is None else {k: acc[k] + m[k].to(acc[k].dtype) for k in m}
    return {k: v / float(len(ths)) for k, v in acc.items()}


# ----------------------------------------------------------------------------------------
# analytic work counters used by bench.py for the roofline numerator
# ----------------------------------------------------------------------------------------
def swinir_flops(cfg: SwinIRCfg, h: int, w: int) -> float:
    """Dense matmul/conv FLOPs (2*MAC) per patch at net-input size h x w (already padded)."""
    T, C, hid = h * w, cfg.embed_dim, int(cfg.embed_dim * cfg.mlp_ratio)
    nblk = sum(cfg.depths)
    mac = T * nblk * (3 * C * C + C * C + 2 * C * hid)       # qkv, proj, fc1, fc2
    mac += T * nblk * 2 * 64 * C                              # q k^T and att v
    mac += T * 9 * cfg.in_chans * C                           # conv_first
    mac += T * (len(cfg.depths) + 1) * 9 * C * C              # RSTB convs + conv_after_body
    s = cfg.upscale
    if cfg.upsampler == "pixelshuffle":
        mac += T * 9 * C * 64
        for k in range(int(round(math.log2(s)))):
            mac += (4 ** k) * T * 9 * 64 * 256
        mac += s * s * T * 9 * 64 * cfg.in_chans
    else:
        mac += T * 9 * C * s * s * cfg.in_chans
    return 2.0 * mac


def swinir_gemm_flops(cfg: SwinIRCfg, h: int, w: int) -> float:
    """Part of swinir_flops executed by the tensor-core GEMM kernel family of the CUDA path
    (everything except the window attention products and the 1-channel input / output convs)."""
    T, C = h * w, cfg.embed_dim
    mac = T * sum(cfg.depths) * 2 * 64 * C + T * 9 * cfg.in_chans * C
    if cfg.upsampler == "pixelshuffle":
        mac += cfg.upscale ** 2 * T * 9 * 64 * cfg.in_chans
    return swinir_flops(cfg, h, w) - 2.0 * mac


def edsr_gemm_flops(cfg: EDSRCfg, h: int, w: int) -> float:
    T, Fe = h * w, cfg.n_feats
    return edsr_flops(cfg, h, w) - 2.0 * (T * 9 * cfg.in_chans * Fe + cfg.scale ** 2 * T * 9 * Fe * cfg.in_chans)


def edsr_flops(cfg: EDSRCfg, h: int, w: int) -> float:
    T, Fe = h * w, cfg.n_feats
    mac = T * 9 * cfg.in_chans * Fe + T * (2 * cfg.n_resblocks + 1) * 9 * Fe * Fe
    for k in range(int(round(math.log2(cfg.scale)))):
        mac += (4 ** k) * T * 9 * Fe * 4 * Fe
    mac += cfg.scale ** 2 * T * 9 * Fe * cfg.in_chans
    return 2.0 * mac


# ------------------------------------------------------------------------------------------
# Bicubic baseline: Interpolate.forward, dlib/utils/utils_trainer.py:120-147 --
# F.interpolate(x, scale_factor=s, mode='bicubic', antialias=True) then clamp to [0, 1].
# The anti-aliased bicubic of PyTorch is the separable PIL filter with a = -0.5: per output index o,
# center = (o + 0.5) / s, taps xmin = max(0, int(center - 2 + 0.5)) .. min(n, int(center + 2 + 0.5)) - 1,
# weight = cubic(j - center + 0.5), renormalised over the taps that exist (no edge replication).
# ------------------------------------------------------------------------------------------
def _cubic_aa(x, a=-0.5):
    x = np.abs(x)
    return np.where(x < 1.0, ((a + 2.0) * x - (a + 3.0)) * x * x + 1.0,
                    np.where(x < 2.0, (((x - 5.0) * x + 8.0) * x - 4.0) * a, 0.0))


def bicubic_aa_taps(n_in: int, scale: int):
    """[(first tap, weights)] for every output index of a 1-D upsampling by `scale`."""
    taps = []
    for o in range(n_in * scale):
        center = (o + 0.5) / scale
        lo = max(0, int(center - 2.0 + 0.5))
        hi = min(n_in, int(center + 2.0 + 0.5))
        w = _cubic_aa(np.arange(lo, hi, dtype=np.float64) - center + 0.5)
        taps.append((lo, w / w.sum()))
    return taps


def interpolate_baseline(x, scale: int):
    """(B, C, h, w) in [0, 1] -> (B, C, h*s, w*s) float32, clamped to [0, 1]."""
    x = np.asarray(x, dtype=np.float64)
    B, C, h, w = x.shape
    tx, ty = bicubic_aa_taps(w, scale), bicubic_aa_taps(h, scale)
    tmp = np.empty((B, C, h, w * scale))
    for o, (lo, wt) in enumerate(tx):
        tmp[..., o] = (x[..., lo:lo + len(wt)] * wt).sum(-1)
    out = np.empty((B, C, h * scale, w * scale))
    for o, (lo, wt) in enumerate(ty):
        out[..., o, :] = (tmp[..., lo:lo + len(wt), :] * wt[:, None]).sum(-2)
    return np.clip(out, 0.0, 1.0).astype(np.float32)
