"""Metric functions with the reference's `dlib.utils.utils_image` signatures, computed by the
single-pass CUDA metrics kernel (libsrk `srk_metrics` / `srk_metrics_roi`).

    tensor2uint82float            dlib/utils/utils_image.py:369-372
    mbatch_gpu_calculate_psnr     :843-891      (fp64, (B,))
    mbatch_gpu_calculate_mse      :894-934      (fp64)
    mbatch_gpu_calculate_nrmse    :937-1007     (fp64)
    mbatch_gpu_calculate_ssim     :1120-1198    (fp32)

plus `compute_metrics` -- what `_compute_metrics` / `marginalize_roi_th_perf`
(dlib/utils/utils_trainer.py:961-1035, :874-930) obtain with ~8 x 40 launches and >= 9 host
syncs per batch -- as ONE kernel pass over E and H and one device-side flag word.
Inputs must be CUDA tensors (1 channel); there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence

import torch

from . import _lib as L

PSNR, MSE, NRMSE, SSIM, PSNR_Y = "psnr", "mse", "nrmse", "ssim", "psnr_y"
ROI_THS_DEFAULT = (4, 5, 6, 7, 8, 9, 10)      # dlib/utils/constants.py:817


def tensor2uint82float(img: torch.Tensor) -> torch.Tensor:
    """Elementwise glue kept in PyTorch for API parity; `compute_metrics` fuses it instead."""
    return (img.float().clamp(0, 1) * 255.0).round().clamp(0, 255).float()


def _check_pair(a, b, roi):
    assert a.ndim == 4, a.ndim
    assert b.ndim == 4, b.ndim
    assert a.shape == b.shape, f"{a.shape} {b.shape}"
    if a.shape[1] != 1:
        raise NotImplementedError("sr_caco_2_b200 metrics are built for 1-channel images")
    if roi is not None:
        assert roi.ndim == 4, roi.ndim
        assert roi.shape[1] == 1, f"dont support c = {roi.shape[1]} > 1."
        assert roi.shape[0] == a.shape[0] and roi.shape[2:] == a.shape[2:], f"{roi.shape} {a.shape}"
    L.require_device(a)
    if b.device != a.device or (roi is not None and roi.device != a.device):
        raise L.SrkError("metric inputs must be on the same CUDA device")


def _run(a, b, border, roi, quantize, ths: Sequence[int] = ()):
    lib = L.load()
    _check_pair(a, b, roi)
    a = a.float().contiguous()
    # uint8 target = the levels the loader holds: no float copy, the kernel reads the bytes
    b_u8 = b.dtype == torch.uint8 and quantize and roi is None and list(ths) == sorted(set(int(t) for t in ths))
    # (a uint8 tensor is levels in [0,255]: scaled to [0,1] only for the quantising path, which maps it back)
    b = b.contiguous() if b_u8 else (b.float().div(255.0) if (b.dtype == torch.uint8 and quantize) else b.float()).contiguous()
    B, _, Hh, Ww = a.shape
    if Hh - 2 * border < 11 or Ww - 2 * border < 11:
        raise ValueError("Kernel size can't be greater than actual input size. "
                         f"Input size: {tuple(a.shape)} border {border}. Kernel size: 11")
    nv = 1 if roi is not None else 1 + len(ths)
    with torch.cuda.device(a.device):
        out = torch.empty(B, nv, L.MET_N, dtype=torch.float64, device=a.device)
        flags = torch.empty(B, dtype=torch.int32, device=a.device)
        scratch = torch.empty(lib.srk_metrics_scratch_bytes(B, len(ths)), dtype=torch.uint8,
                              device=a.device)
        if roi is not None:
            r = roi.float().contiguous()
            L.check(lib.srk_metrics_roi(L.ptr(a), L.ptr(b), L.ptr(r), B, Hh, Ww, border,
                                        int(quantize), L.ptr(out), L.ptr(flags), L.ptr(scratch),
                                        L.stream_ptr()))
        else:
            arr = (C.c_int * max(len(ths), 1))(*[int(t) for t in ths])
            if b_u8:
                L.check(lib.srk_metrics_h8(L.ptr(a), L.ptr(b), B, Hh, Ww, border, arr, len(ths), L.ptr(out),
                                           L.ptr(flags), L.ptr(scratch), L.stream_ptr()))
            else:
                L.check(lib.srk_metrics(L.ptr(a), L.ptr(b), B, Hh, Ww, border, int(quantize), arr,
                                        len(ths), L.ptr(out), L.ptr(flags), L.ptr(scratch),
                                        L.stream_ptr()))
    return out, flags


def mbatch_gpu_calculate_psnr(img1, img2, border: int = 0, roi: Optional[torch.Tensor] = None):
    out, _ = _run(img1, img2, border, roi, False)
    return out[:, 0, L.MET_PSNR].contiguous()


def mbatch_gpu_calculate_mse(img1, img2, border: int = 0, roi: Optional[torch.Tensor] = None):
    out, _ = _run(img1, img2, border, roi, False)
    return out[:, 0, L.MET_MSE].contiguous()


def mbatch_gpu_calculate_nrmse(img, y, border: int = 0, roi: Optional[torch.Tensor] = None):
    out, _ = _run(img, y, border, roi, False)
    return out[:, 0, L.MET_NRMSE].contiguous()


def mbatch_gpu_calculate_ssim(x, y, border: int = 0, roi: Optional[torch.Tensor] = None):
    out, flags = _run(x, y, border, roi, False)
    # the reference asserts 0 <= min/max <= 255 on both inputs (:1164-1172)
    if int((flags & 4).any()):
        raise AssertionError("ssim inputs must lie in [0, 255]")
    return out[:, 0, L.MET_SSIM].float().contiguous()


def compute_metrics(E: torch.Tensor, H: torch.Tensor, border: int,
                    roi_ths: Sequence[int] = (), check: bool = True) -> Dict[str, torch.Tensor]:
    """E in [0,1] (B,1,h,w) on CUDA; H either float in [0,1] or the uint8 levels of the stored
    target (same results, a quarter of the bytes).  Returns per-image fp64 tensors of shape (B,) for the
    five metrics of `_compute_metrics`, and when roi_ths is given, the same five averaged over
    the ROI thresholds under the keys 'roi_<metric>' (marginalize_roi_th_perf).
    One kernel pass; one optional host sync for the NaN/Inf/negative guard
    (check_negative_non_float, utils_trainer.py:933-958 -> raises instead of sys.exit())."""
    if len(roi_ths) > L.MAX_ROI_THS:
        raise ValueError(f"at most {L.MAX_ROI_THS} ROI thresholds")
    out, flags = _run(E, H, border, None, True, roi_ths)
    if check:                                     # (check=False: the caller accumulates res['flags'] and tests them once)
        f = int(flags.max().item()) if flags.numel() else 0
        if f & 1:
            raise FloatingPointError("non-finite metric value (inf/nan)")
        if f & 2:
            raise FloatingPointError("negative metric value")
    names = (PSNR, MSE, NRMSE, SSIM, PSNR_Y)
    res = {n: out[:, 0, i] for i, n in enumerate(names)}
    if len(roi_ths):
        m = out[:, 1:, :].mean(dim=1)
        res.update({"roi_" + n: m[:, i] for i, n in enumerate(names)})
        res["per_threshold"] = out[:, 1:, :]
    res["raw"] = out
    res["flags"] = flags                          # (B,) int32: bit 0 non-finite (metric or input pixel), bit 1 negative metric
    return res
