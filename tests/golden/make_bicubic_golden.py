"""Golden vectors for the bicubic baseline.  The reference's Interpolate.forward
(dlib/utils/utils_trainer.py:120-147) is F.interpolate(x, scale_factor=s, mode='bicubic',
antialias=True) + clamp(0, 1); its constructor asks for a CUDA device, so the two calls are run
directly here on the CPU build of the same PyTorch.  Run in the build container:
    python tests/golden/make_bicubic_golden.py
"""
import os
import numpy as np
import torch
import torch.nn.functional as F

out = {}
g = torch.Generator().manual_seed(77)
for s, (B, h, w) in {2: (2, 13, 9), 4: (2, 8, 16), 8: (1, 11, 7)}.items():
    x = torch.rand(B, 1, h, w, generator=g)
    x[0, 0, :3] = torch.tensor([0.0, 1.0, 0.0])[:, None]          # overshoot on both sides of the clamp
    y = torch.clamp(F.interpolate(input=x, scale_factor=s, mode="bicubic", antialias=True), 0.0, 1.0)
    out[f"x{s}"], out[f"y{s}"] = x.numpy(), y.numpy()
np.savez_compressed(os.path.join(os.path.dirname(__file__), "bicubic_baseline.npz"), **out)
print({k: v.shape for k, v in out.items()})
