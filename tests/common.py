"""Shared test helpers: deterministic reference-layout state_dicts and synthetic inputs.

The state_dict generators enumerate the reference key layout (SURVEY.md appendix B;
network_swinir.py:747-889, network_nlsn.py:325-357) from the constructor arguments.  They are
deterministic for a given torch build (CPU generator), so the build container (where the live
reference is available) and the GPU box (where it is not) see identical weights.
Values are "random-init like" but deliberately have non-zero biases / non-unit LayerNorm gains
so every term of the path is exercised.
"""
from __future__ import annotations

import math
import os
import sys
from typing import Dict

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import sr_oracle as O  # noqa: E402  (tests are allowed to import the oracle)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def _tn(g, shape, std):
    t = torch.empty(shape)
    torch.nn.init.trunc_normal_(t, std=std, a=-2 * std, b=2 * std, generator=g)
    return t


def _conv(g, sd, name, cout, cin):
    k = 1.0 / math.sqrt(cin * 9)
    sd[name + ".weight"] = (torch.rand(cout, cin, 3, 3, generator=g) * 2 - 1) * k
    sd[name + ".bias"] = (torch.rand(cout, generator=g) * 2 - 1) * k


def _ln(g, sd, name, c):
    sd[name + ".weight"] = 1.0 + 0.1 * torch.randn(c, generator=g)
    sd[name + ".bias"] = 0.05 * torch.randn(c, generator=g)


def _lin(g, sd, name, cout, cin):
    sd[name + ".weight"] = _tn(g, (cout, cin), 0.02) * 2.0
    sd[name + ".bias"] = 0.02 * torch.randn(cout, generator=g)


def _conv1(g, sd, name, cout, cin):
    k = 1.0 / math.sqrt(cin)
    sd[name + ".weight"] = (torch.rand(cout, cin, 1, 1, generator=g) * 2 - 1) * k
    sd[name + ".bias"] = (torch.rand(cout, generator=g) * 2 - 1) * k


def _resi(g, sd, name, C, kind):
    if kind == "1conv":
        _conv(g, sd, name, C, C)
    else:                                       # '3conv', network_swinir.py:545-552
        _conv(g, sd, name + ".0", C // 4, C)
        _conv1(g, sd, name + ".2", C // 4, C // 4)
        _conv(g, sd, name + ".4", C, C // 4)


def swinir_state_dict(cfg: O.SwinIRCfg, seed: int = 0) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    C, sd = cfg.embed_dim, {}
    hid = int(C * cfg.mlp_ratio)
    _conv(g, sd, "conv_first", C, cfg.in_chans)
    _ln(g, sd, "patch_embed.norm", C)
    for li, depth in enumerate(cfg.depths):
        nh = cfg.num_heads[li]
        for bi in range(depth):
            pre = f"layers.{li}.residual_group.blocks.{bi}."
            ws, shift = O.block_geometry(cfg, bi)
            _ln(g, sd, pre + "norm1", C)
            if shift > 0:
                res = cfg.img_size
                sd[pre + "attn_mask"] = O.shift_attention_mask(res, res, ws, shift)
            sd[pre + "attn.relative_position_bias_table"] = 0.2 * torch.randn(
                (2 * ws - 1) ** 2, nh, generator=g)
            sd[pre + "attn.relative_position_index"] = O.relative_position_index(ws)
            _lin(g, sd, pre + "attn.qkv", 3 * C, C)
            _lin(g, sd, pre + "attn.proj", C, C)
            _ln(g, sd, pre + "norm2", C)
            _lin(g, sd, pre + "mlp.fc1", hid, C)
            _lin(g, sd, pre + "mlp.fc2", C, hid)
        _resi(g, sd, f"layers.{li}.conv", C, cfg.resi_connection)
    _ln(g, sd, "norm", C)
    _resi(g, sd, "conv_after_body", C, cfg.resi_connection)
    s = cfg.upscale
    if cfg.upsampler == "pixelshuffle":
        _conv(g, sd, "conv_before_upsample.0", 64, C)
        if s & (s - 1) == 0:
            for i in range(int(round(math.log2(s)))):
                _conv(g, sd, f"upsample.{2 * i}", 256, 64)
        else:
            _conv(g, sd, "upsample.0", 9 * 64, 64)
        _conv(g, sd, "conv_last", cfg.in_chans, 64)
    elif cfg.upsampler == "pixelshuffledirect":
        _conv(g, sd, "upsample.0", s * s * cfg.in_chans, C)
    elif cfg.upsampler == "nearest_conv":
        _conv(g, sd, "conv_before_upsample.0", 64, C)
        for name in ("conv_up1", "conv_up2", "conv_hr"):
            _conv(g, sd, name, 64, 64)
        _conv(g, sd, "conv_last", cfg.in_chans, 64)
    else:
        raise NotImplementedError(cfg.upsampler)
    return sd


def edsr_state_dict(cfg: O.EDSRCfg, seed: int = 0) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    Fe, sd = cfg.n_feats, {}
    _conv(g, sd, "head.0", Fe, cfg.in_chans)
    for i in range(cfg.n_resblocks):
        _conv(g, sd, f"body.{i}.body.0", Fe, Fe)
        _conv(g, sd, f"body.{i}.body.2", Fe, Fe)
    _conv(g, sd, f"body.{cfg.n_resblocks}", Fe, Fe)
    s = cfg.scale
    if s & (s - 1) == 0:
        for i in range(int(round(math.log2(s)))):
            _conv(g, sd, f"tail.0.{2 * i}", 4 * Fe, Fe)
    else:
        _conv(g, sd, "tail.0.0", 9 * Fe, Fe)
    _conv(g, sd, "tail.1", cfg.in_chans, Fe)
    return sd


def synthetic_lr(B, h, w, seed, c=1):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(B, c, h, w, generator=g)


def synthetic_pair(B, H, W, seed, realistic=True):
    """(E, Hr) in [0,1]: HR target stored as a uint8-representable image (what the loader
    delivers), estimate = target + noise (so PSNR/SSIM land in a realistic range)."""
    g = torch.Generator().manual_seed(seed)
    base = torch.rand(B, 1, H // 8 + 2, W // 8 + 2, generator=g)
    Hr = torch.nn.functional.interpolate(base, size=(H, W), mode="bilinear", align_corners=False)
    Hr = (Hr * 255).round() / 255
    if realistic:
        E = Hr + 0.05 * torch.randn(B, 1, H, W, generator=g)
    else:
        E = torch.rand(B, 1, H, W, generator=g)
    return E, Hr


# BASELINE.json configs ------------------------------------------------------------------
def cfg_light_x2():      # configs[0]
    return O.SwinIRCfg(upscale=2, in_chans=1, img_size=32, window_size=8, img_range=1.0,
                       depths=[6, 6, 6, 6], embed_dim=60, num_heads=[6, 6, 6, 6], mlp_ratio=2,
                       upsampler="pixelshuffledirect", resi_connection="1conv")


def cfg_classical(scale, img_size=None):   # configs[2..4]; img_size = h_size//scale (128//s)
    return O.SwinIRCfg(upscale=scale, in_chans=1, img_size=img_size or 128 // scale,
                       window_size=8, img_range=1.0, depths=[6] * 6, embed_dim=180,
                       num_heads=[6] * 6, mlp_ratio=2, upsampler="pixelshuffle",
                       resi_connection="1conv")


def cfg_edsr_x4():       # configs[1]
    return O.EDSRCfg(in_chans=1, n_resblocks=16, n_feats=64, scale=4, rgb_range=1.0)


def cfg_imgsize8():
    """img_size == window_size: the reference constructor disables the cyclic shift of every block
    (network_swinir.py:232-236) -- no attn_mask buffers, all windows unshifted."""
    return O.SwinIRCfg(upscale=2, in_chans=1, img_size=8, window_size=8, img_range=1.0, depths=[2, 2],
                       embed_dim=60, num_heads=[6, 6], mlp_ratio=2, upsampler="pixelshuffledirect",
                       resi_connection="1conv")


def cfg_stress():
    return O.SwinIRCfg(upscale=4, in_chans=1, img_size=16, window_size=8, img_range=1.0, depths=[2, 2],
                       embed_dim=180, num_heads=[6, 6], mlp_ratio=2, upsampler="pixelshuffle",
                       resi_connection="1conv")


STRESS_GAIN = 2048.0


def stress_state_dict(sd, gain=STRESS_GAIN):
    """Large-magnitude residual stream: conv_first scaled by `gain`, so every fp16 conv operand (the
    un-normalised residual stream, conv_after_body's sum with the shallow features, the upsampler features)
    is ~gain times larger than with random-init-like weights -- the range a trained checkpoint could reach,
    still below the fp16 maximum the conv operands saturate at."""
    sd = dict(sd)
    sd["conv_first.weight"] = sd["conv_first.weight"] * gain
    sd["conv_first.bias"] = sd["conv_first.bias"] * gain
    return sd


def synthetic_hr(B, H, W, seed):
    """uint8-representable HR targets in [0,1] (what the loader delivers)."""
    g = torch.Generator().manual_seed(seed)
    base = torch.rand(B, 1, H // 8 + 2, W // 8 + 2, generator=g)
    Hr = torch.nn.functional.interpolate(base, size=(H, W), mode="bilinear", align_corners=False)
    return (Hr * 255).round() / 255
