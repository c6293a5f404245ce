"""BASELINE.json workloads (constructor arguments) and the analytic work counters used for the
roofline numerators (dense matmul/conv FLOPs = 2*MAC, un-padded dims)."""
from __future__ import annotations

import math

SWINIR_CLASSICAL = dict(in_chans=1, window_size=8, img_range=1.0, depths=[6] * 6, embed_dim=180,
                        num_heads=[6] * 6, mlp_ratio=2, upsampler="pixelshuffle",
                        resi_connection="1conv")          # utils_init_default_args.py:18-35
SWINIR_LIGHT = dict(in_chans=1, window_size=8, img_range=1.0, depths=[6] * 4, embed_dim=60,
                    num_heads=[6] * 4, mlp_ratio=2, upsampler="pixelshuffledirect",
                    resi_connection="1conv")              # network_swinir.py:993-998 variant

# name: (kind, ctor kwargs, per-GPU batch, h, w, description)
WORKLOADS = {
    "cfg1": ("swinir", dict(SWINIR_LIGHT, upscale=2, img_size=32), 4, 64, 64,
             "SwinIR-light (embed 60, depths 4x6, window 8) X2, 1-ch 64x64 LR patches, batch 4"),
    "cfg2": ("edsr", dict(in_chans=1, n_resblocks=16, n_feats=64, scale=4, rgb_range=1.0), 64, 64, 64,
             "EDSR-baseline (16 res-blocks, 64 feats) X4, 1-ch 64->256, batch 64"),
    "cfg3": ("swinir", dict(SWINIR_CLASSICAL, upscale=8, img_size=16), 32, 64, 64,
             "SwinIR-classical (embed 180, depths 6x6, 6 heads, window 8) X8, 1-ch 64->512, batch 32 per GPU"),
    "cfg4": ("swinir", dict(SWINIR_CLASSICAL, upscale=4, img_size=32), 32, 64, 64,
             "SwinIR-classical X4, 1-ch 64->256, batch 32 per GPU"),
    "cfg5": ("swinir", dict(SWINIR_CLASSICAL, upscale=2, img_size=64), 128, 128, 128,
             "SwinIR-classical X2, 1-ch 128->256, batch 128 per GPU"),
}


def swinir_flops(kw: dict, h: int, w: int, gemm_only: bool = False, attention_in_gemm: bool = False,
                 folded_tail: bool = False) -> float:
    """FLOPs per patch at net-input size h x w (multiples of 8).  gemm_only: the part executed
    by the tensor-core GEMM kernel family (excludes the window-attention products and the
    1-channel input / output convs, which run in their own kernels).
    folded_tail = False counts the network as the reference executes it; True counts what this
    implementation executes when the linear tail (upsample convs + PixelShuffles + conv_last) is
    composed into one 5x5 conv 64 -> s*s (srk_tail_fold): the roofline uses the EXECUTED count."""
    T, C = h * w, kw["embed_dim"]
    hid, nblk, s, cin = int(C * kw["mlp_ratio"]), sum(kw["depths"]), kw["upscale"], kw["in_chans"]
    gemm = T * nblk * (4 * C * C + 2 * C * hid) + T * (len(kw["depths"]) + 1) * 9 * C * C
    attn = T * nblk * 2 * 64 * C                  # q k^T and P v (fused into the qkv GEMM kernel when possible)
    other = T * 9 * cin * C
    if attention_in_gemm:
        gemm += attn
    else:
        other += attn
    if kw["upsampler"] == "pixelshuffle" and folded_tail:
        gemm += T * 9 * C * 64 + T * 25 * 64 * 64          # conv_before_upsample + the folded 5x5 conv (N tile = 64)
    elif kw["upsampler"] == "pixelshuffle":
        gemm += T * 9 * C * 64 + sum((4 ** k) * T * 9 * 64 * 256 for k in range(int(round(math.log2(s)))))
        other += s * s * T * 9 * 64 * cin
    else:
        gemm += T * 9 * C * s * s * cin
    return 2.0 * (gemm if gemm_only else gemm + other)


def edsr_flops(kw: dict, h: int, w: int, gemm_only: bool = False, attention_in_gemm: bool = False,
               folded_tail: bool = False) -> float:
    T, Fe, s, cin = h * w, kw["n_feats"], kw["scale"], kw["in_chans"]
    gemm = T * (2 * kw["n_resblocks"] + 1) * 9 * Fe * Fe
    other = T * 9 * cin * Fe
    if folded_tail and Fe == 64:
        gemm += T * 25 * 64 * 64
    else:
        gemm += sum((4 ** k) * T * 9 * Fe * 4 * Fe for k in range(int(round(math.log2(s)))))
        other += s * s * T * 9 * Fe * cin
    return 2.0 * (gemm if gemm_only else gemm + other)


def build(kind: str, kw: dict):
    from .network_edsr import EDSR
    from .network_swinir import SwinIR
    return SwinIR(**kw) if kind == "swinir" else EDSR(**kw)
