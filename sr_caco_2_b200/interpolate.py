"""Bicubic baseline: host mirror of `Interpolate` (dlib/utils/utils_trainer.py:89-168), the
model-free sweep `evaluate()` runs right after the network (`:1263-1280`).  forward() is
clamp(F.interpolate(L, scale_factor=s, mode='bicubic', antialias=True), 0, 1), computed by
`srk_bicubic_upsample` (csrc/elementwise.cu); the result feeds the same single-pass metrics."""
from __future__ import annotations

import torch

from . import _lib as L

SUPER_RES = "super-resolution"          # dlib/utils/constants.py:2
RECONSTRUCT = "reconstruct"             # constants.RECONSTRUCT
INTER_BICUBIC = "bicubic"               # constants.INTER_BICUBIC


def bicubic_upsample(x: torch.Tensor, scale: int) -> torch.Tensor:
    """(B,1,h,w) in [0,1] on CUDA -> (B,1,h*s,w*s) fp32, clamped to [0,1]."""
    lib = L.load()
    L.require_device(x)
    if x.dim() != 4 or x.shape[1] != 1:
        raise NotImplementedError("bicubic baseline is built for (B,1,h,w) inputs")
    x = x.float().contiguous()
    B, _, h, w = x.shape
    with torch.cuda.device(x.device):
        y = torch.empty(B, 1, h * scale, w * scale, dtype=torch.float32, device=x.device)
        L.check(lib.srk_bicubic_upsample(L.ptr(x), B, h, w, int(scale), L.ptr(y), L.stream_ptr()))
    return y


class Interpolate(torch.nn.Module):
    def __init__(self, task: str, scale: int, scale_mode: str):
        super().__init__()
        self.device = torch.device(f"cuda:{torch.cuda.current_device()}")
        self.scale = int(scale)
        assert task in (SUPER_RES, RECONSTRUCT), task
        self.task = task
        assert scale_mode in [INTER_BICUBIC], scale_mode
        self.scale_mode = scale_mode
        self.L = self.E = self.H = None

    def feed_data(self, data, need_H=True):
        if self.task == SUPER_RES:
            self.L = data["l_im"].to(self.device)
            if need_H:
                self.H = data["h_im"].to(self.device)
        else:
            self.L = data["in_reconstruct"].to(self.device)
            if need_H:
                self.H = data["trg_reconstruct"].to(self.device)

    def forward(self):
        x = self.L
        assert x.ndim == 4, x.ndim
        self.E = bicubic_upsample(x, self.scale if self.task == SUPER_RES else 1)

    def set_eval_mode(self):
        self.eval()

    def set_train_mode(self):
        pass

    def test(self):
        self.eval()
        with torch.no_grad():
            self.forward()

    def current_visuals(self, need_H=True):
        out = {"L": self.L.detach().float(), "E": self.E.detach().float()}
        if need_H:
            out["H"] = self.H.detach().float()
        return out
