"""CPU tests of the host side: C-ABI surface, weight packing, state_dict contract, sharding and
the world_size-2 metric-sum exchange (gloo).  No kernels are launched here."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from tests import common as T
from oracle import sr_oracle as O

ROOT = T.ROOT


@pytest.fixture(scope="module")
def lib():
    from sr_caco_2_b200 import build
    build.build()
    from sr_caco_2_b200 import _lib
    return _lib.load()


def test_library_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "srk.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(srk_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 18
    from sr_caco_2_b200 import _lib
    assert declared == set(_lib.PROTOTYPES), declared ^ set(_lib.PROTOTYPES)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.srk_version() == 100
    assert isinstance(lib.srk_last_error(), bytes)


def test_ctypes_struct_layouts_match_header_sizes(lib, tmp_path):
    """sizeof() of every struct crossing the boundary, as the C compiler sees it."""
    src = tmp_path / "sz.c"
    src.write_text('#include "srk.h"\n#include <stdio.h>\nint main(){printf("%zu %zu %zu %zu %zu\\n",'
                   'sizeof(srk_gemm_args),sizeof(srk_conv_params),sizeof(srk_stb_params),'
                   'sizeof(srk_swinir_plan),sizeof(srk_edsr_plan));printf("%zu %zu %zu\\n",sizeof(srk_mlp_args),sizeof(srk_tail_fold),sizeof(srk_attn_block_args));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    sizes = list(map(int, subprocess.check_output([str(exe)]).split()))
    from sr_caco_2_b200 import _lib as L
    got = [ctypes.sizeof(c) for c in (L.GemmArgs, L.ConvParams, L.StbParams, L.SwinIRPlan, L.EDSRPlan, L.MlpArgs, L.TailFold, L.AttnBlockArgs)]
    assert got == sizes


def test_cpu_tensor_is_an_error_not_a_fallback(lib):
    from sr_caco_2_b200 import SwinIR, SrkError, utils_image
    net = SwinIR(upscale=2, in_chans=1, img_size=16, window_size=8, depths=[2], embed_dim=60,
                 num_heads=[6], mlp_ratio=2, upsampler="pixelshuffledirect").eval()
    with pytest.raises(SrkError):
        net(torch.rand(1, 1, 16, 16))
    with pytest.raises(SrkError):
        utils_image.mbatch_gpu_calculate_psnr(torch.rand(1, 1, 32, 32), torch.rand(1, 1, 32, 32))


def test_unsupported_variants_raise_at_construction():
    from sr_caco_2_b200 import SwinIR
    with pytest.raises(NotImplementedError):
        SwinIR(in_chans=3, window_size=8, upsampler="pixelshuffle")
    with pytest.raises(NotImplementedError):
        SwinIR(in_chans=1, window_size=8, upsampler="nearest+conv", upscale=4)       # not a reference constant
    with pytest.raises(NotImplementedError):
        SwinIR(in_chans=1, window_size=8, upsampler="", upscale=1)                   # denoising branch: out of scope
    with pytest.raises(AssertionError):                                              # network_swinir.py:877
        SwinIR(in_chans=1, window_size=8, upsampler="nearest_conv", upscale=2)
    with pytest.raises(NotImplementedError):   # img_size 4 would shrink the window to 4
        SwinIR(in_chans=1, window_size=8, img_size=4, upsampler="pixelshuffle", upscale=2)


@pytest.mark.parametrize("cfg", [T.cfg_light_x2(), T.cfg_classical(8), T.cfg_classical(2),
                                 O.SwinIRCfg(upscale=8, in_chans=1, img_size=8, depths=[2], embed_dim=60,
                                             num_heads=[6], mlp_ratio=2, upsampler="pixelshuffledirect"),
                                 O.SwinIRCfg(upscale=4, in_chans=1, img_size=16, depths=[2, 2], embed_dim=60,
                                             num_heads=[6, 6], mlp_ratio=2, upsampler="nearest_conv",
                                             resi_connection="3conv")])
def test_swinir_state_dict_contract(cfg):
    """Key names / shapes / dtypes equal the reference layout (tests/common.py enumerates it and
    test_oracle checks that enumeration against the live reference with strict=True)."""
    from sr_caco_2_b200 import SwinIR
    net = SwinIR(upscale=cfg.upscale, in_chans=cfg.in_chans, img_size=cfg.img_size,
                 window_size=cfg.window_size, img_range=cfg.img_range, depths=cfg.depths,
                 embed_dim=cfg.embed_dim, num_heads=cfg.num_heads, mlp_ratio=cfg.mlp_ratio,
                 upsampler=cfg.upsampler, resi_connection=cfg.resi_connection)
    ref = T.swinir_state_dict(cfg, seed=1)
    own = net.state_dict()
    assert set(own) == set(ref)
    for k in ref:
        assert own[k].shape == ref[k].shape and own[k].dtype == ref[k].dtype, k
        if k.endswith("relative_position_index") or k.endswith("attn_mask"):
            assert torch.equal(own[k], ref[k]), k
    net.load_state_dict(ref, strict=True)
    # shift decisions follow the constructor (img_size), not the runtime tensor
    geo = [O.block_geometry(cfg, bi) for d in cfg.depths for bi in range(d)]
    assert net._block_geometry == geo


def test_edsr_state_dict_contract_and_define_G():
    from sr_caco_2_b200 import EDSR, define_G
    cfg = T.cfg_edsr_x4()
    net = EDSR(in_chans=1, n_resblocks=16, n_feats=64, scale=4)
    ref = T.edsr_state_dict(cfg, seed=1)
    own = net.state_dict()
    assert set(own) == set(ref) and all(own[k].shape == ref[k].shape for k in ref)
    net.load_state_dict(ref, strict=True)

    class Args:
        netG = {"net_type": "swinir", "swinir_upscale": 8, "swinir_in_chans": 1, "swinir_img_size": 16,
                "swinir_window_size": 8, "swinir_img_range": 1.0, "swinir_depths": [6] * 6,
                "swinir_embed_dim": 180, "swinir_num_heads": [6] * 6, "swinir_mlp_ratio": 2,
                "swinir_upsampler": "pixelshuffle", "swinir_resi_connection": "1conv"}
    g = define_G(Args)
    assert sum(p.numel() for p in g.parameters()) == 12043517        # SURVEY 8a1, classical X8
    Args.netG = {"net_type": "EDSR_LIIF", "EDSR_LIIF_in_chans": 1, "EDSR_LIIF_n_resblocks": 16,
                 "EDSR_LIIF_n_feats": 64, "EDSR_LIIF_upscale": 4, "EDSR_LIIF_img_range": 1.0}
    e = define_G(Args)
    assert sum(p.numel() for p in e.parameters()) == 1515265
    Args.netG = {"net_type": "srcnn"}
    with pytest.raises(NotImplementedError):
        define_G(Args)


def test_weight_packing_matches_linear_and_conv_semantics():
    from sr_caco_2_b200 import packing as P, _lib as L
    g = torch.Generator().manual_seed(0)
    C, nh, d, dp = 60, 6, 10, 16
    Cp, ao_p, nq_p = 64, 128, 320
    w = torch.randn(3 * C, C, generator=g)
    b = torch.randn(3 * C, generator=g)
    x = torch.randn(5, C, generator=g)
    wq, bq = P.pack_qkv(w, b, nh, d, dp, nq_p, Cp, L.SRK_FP16)
    xp = torch.zeros(5, Cp)
    xp[:, :C] = x
    got = xp @ wq.float().t() + bq
    ref = (x @ w.half().float().t() + b).view(5, 3, nh, d)
    got3 = got[:, :3 * nh * dp].view(5, 3, nh, dp)
    assert torch.allclose(got3[..., :d], ref, atol=1e-5)
    assert float(got3[..., d:].abs().max()) == 0.0 and float(got[:, 3 * nh * dp:].abs().max()) == 0.0
    # proj consumes the head-padded layout
    wpj = torch.randn(C, C, generator=g)
    ao = torch.randn(5, nh, d, generator=g)
    aop = torch.zeros(5, ao_p)
    aop[:, :nh * dp].view(5, nh, dp)[..., :d] = ao
    pk = P.pack_proj(wpj, nh, d, dp, Cp, ao_p, L.SRK_FP16)
    assert torch.allclose((aop @ pk.float().t())[:, :C], ao.reshape(5, C) @ wpj.half().float().t(), atol=1e-5)
    # conv: k = tap*cin_p + c ; pixel-shuffle row permutation
    cw = torch.randn(256, 64, 3, 3, generator=g)
    cb = torch.randn(256, generator=g)
    img = torch.randn(1, 64, 5, 6, generator=g)
    pk, pb = P.pack_conv3x3(cw, cb, 64, 256, L.SRK_FP16, pixel_shuffle_r=2)
    cols = F.unfold(img, 3, padding=1).view(1, 64, 9, 30).permute(0, 3, 2, 1).reshape(30, 9 * 64)
    out = cols @ pk.float().t() + pb                                   # (pixels, (i,j,c))
    ref = F.pixel_shuffle(F.conv2d(img, cw.half().float(), cb, padding=1), 2)   # (1,64,10,12)
    out = out.view(5, 6, 2, 2, 64).permute(4, 0, 2, 1, 3).reshape(64, 10, 12)
    assert torch.allclose(out, ref[0], atol=1e-3)
    assert P.pack_conv_out(torch.arange(18.).view(1, 2, 3, 3)).float().tolist()[0] == [0.0, 9.0]


def test_shard_range_is_exact_and_contiguous():
    from sr_caco_2_b200.evaluator import shard_range
    for n in (0, 1, 7, 1471, 1472):
        for world in (1, 2, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_pad_for_windows_matches_reference_formula():
    from sr_caco_2_b200.evaluator import pad_for_windows
    x = torch.arange(2 * 1 * 64 * 64, dtype=torch.float32).view(2, 1, 64, 64)
    y = pad_for_windows(x, 8)
    assert y.shape == (2, 1, 72, 72)
    assert torch.equal(y[:, :, :64, :64], x)
    assert torch.equal(y[:, :, 64, :64], x[:, :, 63, :])       # mirrored WITH the edge repeated
    assert torch.equal(y[:, :, 71, :64], x[:, :, 56, :])
    assert pad_for_windows(torch.zeros(1, 1, 60, 70), 8).shape == (1, 1, 64, 72)


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sr_caco_2_b200.evaluator import evaluate_patches
    E, H = T.synthetic_pair(7, 48, 40, 21)

    def step(e, h):    # CPU stand-in for the CUDA step: the oracle's metrics on the pair itself
        a, b = O.all_metrics(e, h, 2), O.roi_marginal_metrics(e, h, 2)
        return torch.stack([a[k].double() for k in ("psnr", "mse", "nrmse", "ssim", "psnr_y")] +
                           [b[k].double() for k in ("psnr", "mse", "nrmse", "ssim", "psnr_y")], 1)
    ids = [f"cell/{i:03d}.tif" for i in range(7)]
    res = evaluate_patches(step, E, H, batch_size=2, rank=rank, world=world, device=torch.device("cpu"), ids=ids)
    q.put((rank, res))
    dist.destroy_process_group()


def test_two_rank_gloo_metric_exchange_equals_single_process():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    results = dict(q.get(timeout=120) for _ in range(2))
    [p.join(60) for p in procs]
    E, H = T.synthetic_pair(7, 48, 40, 21)
    a, b = O.all_metrics(E, H, 2), O.roi_marginal_metrics(E, H, 2)
    for r in (0, 1):
        assert results[r]["n"] == 7
        for k in ("psnr", "mse", "nrmse", "ssim", "psnr_y"):
            assert abs(results[r][k] - float(a[k].double().mean())) < 1e-9 * max(1, abs(results[r][k]))
            assert abs(results[r]["roi_" + k] - float(b[k].double().mean())) < 1e-9 * max(1, abs(results[r][k]))
        # per-image details: ragged shards (4 + 3 images) gathered with ONE all_gather, every rank holds all 7 images in order
        det, roi = results[r]["details"], results[r]["roi_details"]
        assert list(det.keys()) == [f"cell/{i:03d}.tif" for i in range(7)]
        for i, im_id in enumerate(det):
            for k in ("psnr", "mse", "nrmse", "ssim", "psnr_y"):
                assert abs(det[im_id][k] - float(a[k][i])) < 1e-9 * max(1, abs(det[im_id][k]))
                assert abs(roi[im_id][k] - float(b[k][i])) < 1e-9 * max(1, abs(roi[im_id][k]))


def test_details_and_summary_files_have_the_reference_shape(tmp_path):
    """details_<ds>.yml / roi_details_<ds>.yml (utils_trainer.py:1140-1147) and the write_current_perf_eval summary
    (utils_tracker.py:133-165): same file names and keys."""
    import yaml
    from sr_caco_2_b200 import evaluator as EV
    block = torch.arange(30, dtype=torch.float64).view(3, 10)
    det, roi = EV.details_dicts(block, ["a", "b", "c"])
    EV.write_details(det, roi, str(tmp_path), "caco2")
    d = yaml.safe_load(open(tmp_path / "details_caco2.yml"))
    r = yaml.safe_load(open(tmp_path / "roi_details_caco2.yml"))
    assert d["b"] == {"psnr": 10.0, "mse": 11.0, "nrmse": 12.0, "ssim": 13.0, "psnr_y": 14.0}
    assert r["c"]["psnr"] == 25.0
    means = {m: 1.0 + i for i, m in enumerate(EV.METRICS)}
    means.update({"roi_" + m: 2.0 + i for i, m in enumerate(EV.METRICS)})
    out = EV.write_current_perf_eval(means, "test", "caco2", str(tmp_path), "perf.yml", current_step=7)
    y = yaml.safe_load(open(tmp_path / "perf.yml"))
    assert y == out and y["last_ssim"] == 4.0 and y["best_psnr"] == 1.0 and y["dataset"] == "caco2" and y["split"] == "test"


def test_fold_tail_equals_layer_chain_on_cpu():
    """packing.fold_tail: the composed 5x5 kernels (interior + 8 border variants, fragment-ordered for the
    ring pass) reproduce [Conv3x3 + PixelShuffle(2)] x n + Conv3x3 at every pixel (fp64 evaluation)."""
    import torch.nn.functional as F
    from sr_caco_2_b200 import packing as P
    g = torch.Generator().manual_seed(3)
    for s, n in ((2, 1), (4, 2), (8, 3)):
        ups = [(torch.randn(256, 64, 3, 3, generator=g) * 0.05, torch.randn(256, generator=g) * 0.1) for _ in range(n)]
        lw, lb = torch.randn(1, 64, 3, 3, generator=g) * 0.05, torch.randn(1, generator=g) * 0.1
        fw, fb, bw, bb, wsc = P.fold_tail(ups, lw, lb, s)
        assert fw.shape == (64, 1600) and bw.shape == (9, 64, 1600) and fw.dtype == torch.float16
        # undo the ring-pass fragment order: [t][ks][w][h] -> k = ks*16 + 8w + 2t + h
        bn = bw.view(9, 64, 25, 4, 4, 2, 2).permute(0, 1, 2, 4, 5, 3, 6).reshape(9, 64, 1600)
        assert torch.equal(bn[4], fw)
        H, W = 6, 7
        x = torch.randn(2, 64, H, W, generator=g).double()
        y = x
        for w_, b_ in ups:
            y = F.pixel_shuffle(F.conv2d(y, w_.double(), b_.double(), padding=1), 2)
        y = F.conv2d(y, lw.double(), lb.double(), padding=1)[:, 0]
        xp = F.pad(x, (2, 2, 2, 2))
        out = torch.zeros_like(y)
        for yy in range(H):
            for xx in range(W):
                v = (0 if yy == 0 else 2 if yy == H - 1 else 1) * 3 + (0 if xx == 0 else 2 if xx == W - 1 else 1)
                patch = xp[:, :, yy:yy + 5, xx:xx + 5].permute(0, 2, 3, 1).reshape(2, 1600)
                o = (patch @ bn[v].double().t() + bb[v].double()) / wsc
                out[:, yy * s:(yy + 1) * s, xx * s:(xx + 1) * s] = o[:, :s * s].view(2, s, s)
        assert float((out - y).abs().max()) < 3e-3 * max(1.0, float(y.abs().max()))     # fp16 weight rounding only


def test_pack_conv_nearest2x_equals_interpolate_then_conv_on_cpu():
    """packing.pack_conv_nearest2x: nearest x2 + Conv3x3 == 3x3 conv on the low-res grid + PixelShuffle(2)."""
    import torch.nn.functional as F
    from sr_caco_2_b200 import _lib as L, packing as P
    g = torch.Generator().manual_seed(4)
    w, b = torch.randn(64, 64, 3, 3, generator=g) * 0.05, torch.randn(64, generator=g) * 0.1
    x = torch.randn(2, 64, 5, 7, generator=g)
    ref = F.conv2d(F.interpolate(x, scale_factor=2, mode="nearest"), w, b, padding=1)
    wk, bk = P.pack_conv_nearest2x(w, b, 64, L.SRK_FP16)
    assert wk.shape == (256, 576) and bk.shape == (256,)
    wc = wk.float().view(2, 2, 64, 3, 3, 64).permute(0, 1, 2, 5, 3, 4).reshape(256, 64, 3, 3)   # rows (i, j, c)
    low = F.conv2d(x, wc, bk, padding=1).view(2, 2, 2, 64, 5, 7)                                # [b][i][j][c][y][x]
    got = low.permute(0, 3, 4, 1, 5, 2).reshape(2, 64, 10, 14)
    assert float((got - ref).abs().max()) < 2e-3
