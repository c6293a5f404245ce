"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference
(/root/reference, imported through oracle/ref_import.py) on CPU.

Run in the build container only (the reference does not exist on the GPU box):
    python tests/golden/make_golden.py

Outputs (committed):
    index_maps.json          sha256 + samples of the reference's integer index maps
    metrics_kat.json         reference metric values on seeded inputs (incl. ROI / border / edge)
    swinir_tiny_direct.npz   tiny SwinIR (pixelshuffledirect): state_dict + inputs + outputs
    swinir_tiny_ps.npz       tiny SwinIR (pixelshuffle X4)
    edsr_tiny.npz            tiny EDSR-baseline assembled from network_nlsn.py primitives
    fullsize_samples.json    strided samples of the reference output for the BASELINE configs
                             on tests/common.py's deterministic state_dicts
"""
import hashlib
import json
import os
import sys

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_import as R          # noqa: E402
from oracle import sr_oracle as O           # noqa: E402
from tests import common as T               # noqa: E402


def sha(t: torch.Tensor) -> str:
    return hashlib.sha256(t.contiguous().numpy().tobytes()).hexdigest()


def build_ref_swinir(cfg: O.SwinIRCfg, sd):
    net = R.swinir_class()(upscale=cfg.upscale, in_chans=cfg.in_chans, img_size=cfg.img_size,
                           window_size=cfg.window_size, img_range=cfg.img_range,
                           depths=cfg.depths, embed_dim=cfg.embed_dim, num_heads=cfg.num_heads,
                           mlp_ratio=cfg.mlp_ratio, upsampler=cfg.upsampler,
                           resi_connection=cfg.resi_connection)
    net.load_state_dict(sd, strict=True)
    return net.eval()


class RefEDSR(nn.Module):
    """EDSR-baseline wired from the reference's own primitives exactly as NLSN wires them
    (network_nlsn.py:325-369) with the NLSA blocks left out."""

    def __init__(self, cfg: O.EDSRCfg):
        super().__init__()
        conv, ResBlock, Upsampler = R.nlsn_primitives()
        act = nn.ReLU(True)
        self.head = nn.Sequential(conv(cfg.in_chans, cfg.n_feats, 3))
        body = [ResBlock(conv, cfg.n_feats, 3, act=act, res_scale=cfg.res_scale)
                for _ in range(cfg.n_resblocks)]
        body.append(conv(cfg.n_feats, cfg.n_feats, 3))
        self.body = nn.Sequential(*body)
        self.tail = nn.Sequential(Upsampler(conv, cfg.scale, cfg.n_feats, act=False),
                                  nn.Conv2d(cfg.n_feats, cfg.in_chans, 3, padding=1))

    def forward(self, x):
        x = self.head(x)
        res = self.body(x) + x
        return self.tail(res)


def index_maps():
    SwinIR = R.swinir_class()
    from dlib.models.network_swinir import SwinTransformerBlock, window_partition
    out = {}
    net = SwinIR(upscale=2, in_chans=1, img_size=16, window_size=8, depths=[2], embed_dim=12,
                 num_heads=[2], mlp_ratio=2, upsampler="pixelshuffledirect")
    blk = net.layers[0].residual_group.blocks[1]
    rpi = net.layers[0].residual_group.blocks[0].attn.relative_position_index
    out["relative_position_index_ws8"] = {"sha256": sha(rpi), "sum": int(rpi.sum()),
                                          "row0": rpi[0, :10].tolist()}
    for (H, W) in [(16, 16), (64, 64), (72, 72), (128, 128), (136, 136), (24, 40)]:
        m = blk.calculate_mask((H, W))
        out[f"mask_{H}x{W}"] = {"sha256": sha(m), "n_neg": int((m != 0).sum()),
                                "n_windows_masked": int((m != 0).flatten(1).any(1).sum())}
        for shift in (0, 4):
            ids = torch.arange(H * W, dtype=torch.int64).view(1, H, W, 1)
            if shift:
                ids = torch.roll(ids, (-shift, -shift), (1, 2))
            gm = window_partition(ids, 8).reshape(-1)
            out[f"gather_{H}x{W}_s{shift}"] = {"sha256": sha(gm), "head": gm[:8].tolist(),
                                               "tail": gm[-8:].tolist()}
    for (C, H, W, r) in [(1, 4, 4, 2), (2, 3, 5, 2), (1, 4, 4, 8), (64, 6, 6, 2), (1, 3, 3, 3)]:
        n = C * r * r * H * W
        ps = torch.nn.functional.pixel_shuffle(
            torch.arange(n, dtype=torch.int64).view(1, C * r * r, H, W).double(), r).long()
        out[f"pixelshuffle_C{C}_H{H}_W{W}_r{r}"] = {"sha256": sha(ps.reshape(-1))}
    return out


def metric_cases():
    ui = R.utils_image()
    cases = []

    def run(name, E, H, border, roi_th=None, prequant=False):
        e8 = E if prequant else ui.tensor2uint82float(E)
        h8 = H if prequant else ui.tensor2uint82float(H)
        roi = None if roi_th is None else (h8 >= roi_th).float()
        rep = lambda t: t.repeat(1, 3, 1, 1)
        ey = ui.mb_gpu_rgb2ycbcr(rep(e8) / 255.0, only_y=True) * 255.0
        hy = ui.mb_gpu_rgb2ycbcr(rep(h8) / 255.0, only_y=True) * 255.0
        cases.append({
            "name": name, "border": border, "roi_th": roi_th,
            "sum_e8": float(e8.double().sum()), "sum_h8": float(h8.double().sum()),
            "psnr": ui.mbatch_gpu_calculate_psnr(e8, h8, border=border, roi=roi).tolist(),
            "mse": ui.mbatch_gpu_calculate_mse(e8, h8, border=border, roi=roi).tolist(),
            "nrmse": ui.mbatch_gpu_calculate_nrmse(e8, h8, border=border, roi=roi).tolist(),
            "ssim": ui.mbatch_gpu_calculate_ssim(e8, h8, border=border, roi=roi).double().tolist(),
            "psnr_y": ui.mbatch_gpu_calculate_psnr(ey, hy, border=border, roi=roi).tolist(),
        })

    # appendix-B KAT: torch.manual_seed(0) rand pair
    torch.manual_seed(0)
    E = torch.rand(2, 1, 512, 512)
    H = torch.rand(2, 1, 512, 512)
    run("seed0_rand_512", E, H, 8)
    run("seed0_rand_512_roi7", E, H, 8, 7)
    run("seed0_identical", H, H, 8)
    # realistic pairs from tests/common.py (regenerated identically on the GPU box)
    for (name, B, Hh, Ww, seed, border) in [("real_256", 3, 256, 256, 11, 4),
                                            ("real_512", 2, 512, 512, 12, 8),
                                            ("real_128", 4, 128, 128, 13, 2),
                                            ("real_ragged", 2, 72, 104, 14, 2),
                                            ("real_noborder", 2, 64, 48, 15, 0),
                                            ("real_min", 1, 11, 11, 16, 0)]:
        E, H = T.synthetic_pair(B, Hh, Ww, seed)
        run(name, E, H, border)
        for th in (4, 7, 10):
            run(f"{name}_roi{th}", E, H, border, th)
    # edge cases: black target (empty ROI, zero dynamic range), saturated estimate
    Z = torch.zeros(2, 1, 64, 64)
    run("black_black", Z, Z, 2)
    run("black_black_roi4", Z, Z, 2, 4)
    run("white_vs_black", torch.ones(2, 1, 64, 64) * 1.7, Z, 2)
    E, H = T.synthetic_pair(2, 64, 64, 17)
    run("out_of_range_est", E * 3 - 1, H, 2)
    return cases


def tiny_nets():
    specs = {
        "swinir_tiny_direct": (O.SwinIRCfg(upscale=2, in_chans=1, img_size=32, window_size=8,
                                           depths=[2, 2], embed_dim=60, num_heads=[6, 6],
                                           mlp_ratio=2, upsampler="pixelshuffledirect"),
                               [(2, 24, 24), (1, 20, 29)]),
        "swinir_tiny_ps": (O.SwinIRCfg(upscale=4, in_chans=1, img_size=16, window_size=8,
                                       depths=[2], embed_dim=60, num_heads=[6], mlp_ratio=2,
                                       upsampler="pixelshuffle"),
                           [(2, 16, 24)]),
    }
    for name, (cfg, shapes) in specs.items():
        sd = T.swinir_state_dict(cfg, seed=7)
        net = build_ref_swinir(cfg, sd)
        arrs = {"sd::" + k: v.numpy() for k, v in sd.items()}
        for i, (B, h, w) in enumerate(shapes):
            x = T.synthetic_lr(B, h, w, 100 + i)
            with torch.no_grad():
                y = net(x)
            arrs[f"x{i}"] = x.numpy()
            arrs[f"y{i}"] = y.numpy()
        arrs["cfg"] = np.frombuffer(json.dumps(cfg.__dict__).encode(), dtype=np.uint8)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrs)
        print(name, {k: v.shape for k, v in arrs.items() if not k.startswith("sd::")})

    cfg = O.EDSRCfg(in_chans=1, n_resblocks=2, n_feats=64, scale=2)
    sd = T.edsr_state_dict(cfg, seed=8)
    net = RefEDSR(cfg).eval()
    net.load_state_dict(sd, strict=True)
    x = T.synthetic_lr(2, 20, 28, 103)
    with torch.no_grad():
        y = net(x)
    arrs = {"sd::" + k: v.numpy() for k, v in sd.items()}
    arrs.update(x0=x.numpy(), y0=y.numpy(),
                cfg=np.frombuffer(json.dumps(cfg.__dict__).encode(), dtype=np.uint8))
    np.savez_compressed(os.path.join(HERE, "edsr_tiny.npz"), **arrs)


def sample(y: torch.Tensor, n=512):
    flat = y.reshape(-1)
    idx = torch.linspace(0, flat.numel() - 1, n).long()
    return {"idx": idx.tolist(), "val": flat[idx].double().tolist(),
            "sum": float(flat.double().sum()), "absmax": float(flat.abs().max()),
            "shape": list(y.shape)}


def fullsize():
    out = {}
    jobs = [("cfg1_light_x2_64", T.cfg_light_x2(), (4, 64, 64), 101),
            ("cfg1_light_x2_72", T.cfg_light_x2(), (2, 72, 72), 101),
            ("cfg3_classical_x8_64", T.cfg_classical(8), (1, 64, 64), 103),
            ("cfg4_classical_x4_72", T.cfg_classical(4), (1, 72, 72), 104)]
    for name, cfg, (B, h, w), seed in jobs:
        sd = T.swinir_state_dict(cfg, seed=seed)
        net = build_ref_swinir(cfg, sd)
        x = T.synthetic_lr(B, h, w, seed)
        with torch.no_grad():
            y = net(x)
        out[name] = sample(y)
        print(name, y.shape, float(y.mean()), float(y.std()))
    cfg = T.cfg_edsr_x4()
    sd = T.edsr_state_dict(cfg, seed=102)
    net = RefEDSR(cfg).eval()
    net.load_state_dict(sd, strict=True)
    x = T.synthetic_lr(2, 64, 64, 102)
    with torch.no_grad():
        y = net(x)
    out["cfg2_edsr_x4_64"] = sample(y)
    print("cfg2", y.shape, float(y.mean()), float(y.std()))
    return out


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    json.dump(index_maps(), open(os.path.join(HERE, "index_maps.json"), "w"), indent=1)
    json.dump(metric_cases(), open(os.path.join(HERE, "metrics_kat.json"), "w"), indent=1)
    tiny_nets()
    json.dump(fullsize(), open(os.path.join(HERE, "fullsize_samples.json"), "w"))
    print("golden fixtures written to", HERE)
