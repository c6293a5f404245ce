"""Golden fixtures for the SwinIR variants of SURVEY 8f-4, from the UNMODIFIED reference on CPU:
    swinir_tiny_3conv.npz     resi_connection '3conv' (network_swinir.py:545-552, 874-884), pixelshuffle X2
    swinir_tiny_nearest.npz   upsampler 'nearest_conv' X4 (:875-886, 948-961)
Run in the build container:  python tests/golden/make_variant_golden.py
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)

from oracle import sr_oracle as O           # noqa: E402
from tests import common as T               # noqa: E402
from make_golden import build_ref_swinir    # noqa: E402

specs = {
    "swinir_tiny_3conv": (O.SwinIRCfg(upscale=2, in_chans=1, img_size=16, window_size=8, depths=[2, 2], embed_dim=60,
                                      num_heads=[6, 6], mlp_ratio=2, upsampler="pixelshuffle",
                                      resi_connection="3conv"), [(2, 16, 24), (1, 21, 13)]),
    "swinir_tiny_nearest": (O.SwinIRCfg(upscale=4, in_chans=1, img_size=16, window_size=8, depths=[2], embed_dim=60,
                                        num_heads=[6], mlp_ratio=2, upsampler="nearest_conv"), [(2, 16, 24), (1, 19, 9)]),
}
for name, (cfg, shapes) in specs.items():
    sd = T.swinir_state_dict(cfg, seed=9)
    net = build_ref_swinir(cfg, sd)
    arrs = {"sd::" + k: v.numpy() for k, v in sd.items()}
    for i, (B, h, w) in enumerate(shapes):
        x = T.synthetic_lr(B, h, w, 200 + i)
        with torch.no_grad():
            y = net(x)
        arrs[f"x{i}"], arrs[f"y{i}"] = x.numpy(), y.numpy()
        ours = O.swinir_forward(sd, cfg, x)
        print(name, i, tuple(y.shape), "oracle max-abs diff", float((ours - y).abs().max()))
    arrs["cfg"] = np.frombuffer(json.dumps(cfg.__dict__).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrs)
