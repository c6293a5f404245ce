"""CPU tests: pin the oracle port (oracle/sr_oracle.py) against golden vectors produced by the
unmodified reference (tests/golden/make_golden.py), and against the live reference when
/root/reference is present."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from tests import common as T
from oracle import sr_oracle as O
from oracle import ref_import as R

G = T.GOLDEN


def sha(t):
    return hashlib.sha256(t.contiguous().numpy().tobytes()).hexdigest()


def load_npz(name):
    z = np.load(os.path.join(G, name + ".npz"))
    sd = {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd::")}
    cfg = json.loads(bytes(z["cfg"]).decode())
    return z, sd, cfg


def test_index_maps_bit_exact():
    gold = json.load(open(os.path.join(G, "index_maps.json")))
    rpi = O.relative_position_index(8)
    assert sha(rpi) == gold["relative_position_index_ws8"]["sha256"]
    assert sha(rpi)[:16] == "4a65ba32bf59f64d"
    assert int(rpi.sum()) == 458752 and rpi[0, :3].tolist() == [112, 111, 110]
    for key, val in gold.items():
        if key.startswith("mask_"):
            H, W = map(int, key[5:].split("x"))
            m = O.shift_attention_mask(H, W, 8, 4)
            assert sha(m) == val["sha256"], key
            assert int((m != 0).sum()) == val["n_neg"]
        elif key.startswith("gather_"):
            hw, s = key[7:].split("_s")
            H, W = map(int, hw.split("x"))
            gm = O.window_gather_map(H, W, 8, int(s))
            assert sha(gm) == val["sha256"], key
            assert gm[:8].tolist() == val["head"] and gm[-8:].tolist() == val["tail"]
        elif key.startswith("pixelshuffle_"):
            C, H, W, r = [int(p[1:]) for p in key.split("_")[1:]]
            assert sha(O.pixel_shuffle_map(C, H, W, r).reshape(-1)) == val["sha256"], key
    assert sha(O.shift_attention_mask(64, 64, 8, 4))[:16] == "ab76e9ac4c9729a2"
    assert sha(O.shift_attention_mask(72, 72, 8, 4))[:16] == "dd17dbf752066f4a"
    assert sha(O.window_gather_map(64, 64, 8, 4))[:16] == "1b83697a39487f45"
    assert sha(O.window_gather_map(64, 64, 8, 0))[:16] == "88b0f471f7708407"


def _inputs_for(case):
    name = case["name"]
    base = name.split("_roi")[0]
    if base.startswith("seed0"):
        torch.manual_seed(0)
        E = torch.rand(2, 1, 512, 512)
        H = torch.rand(2, 1, 512, 512)
        return (H, H) if base == "seed0_identical" else (E, H)
    table = {"real_256": (3, 256, 256, 11), "real_512": (2, 512, 512, 12),
             "real_128": (4, 128, 128, 13), "real_ragged": (2, 72, 104, 14),
             "real_noborder": (2, 64, 48, 15), "real_min": (1, 11, 11, 16)}
    if base in table:
        B, Hh, Ww, seed = table[base]
        return T.synthetic_pair(B, Hh, Ww, seed)
    Z = torch.zeros(2, 1, 64, 64)
    if base == "black_black":
        return Z, Z
    if base == "white_vs_black":
        return torch.ones(2, 1, 64, 64) * 1.7, Z
    if base == "out_of_range_est":
        E, H = T.synthetic_pair(2, 64, 64, 17)
        return E * 3 - 1, H
    raise KeyError(name)


def test_metrics_against_reference_golden():
    cases = json.load(open(os.path.join(G, "metrics_kat.json")))
    assert len(cases) > 30
    for c in cases:
        E, H = _inputs_for(c)
        m = O.all_metrics(E, H, c["border"], c["roi_th"])
        assert float(O.quantize_u8f(E).double().sum()) == c["sum_e8"], c["name"]
        for k in ("psnr", "mse", "nrmse", "psnr_y"):
            np.testing.assert_allclose(m[k].numpy(), np.array(c[k]), rtol=1e-12, atol=1e-12,
                                       err_msg=f"{c['name']} {k}")
        np.testing.assert_allclose(m["ssim"].double().numpy(), np.array(c["ssim"]),
                                   rtol=0, atol=2e-6, err_msg=c["name"])


def test_metric_known_answers_appendix_b():
    torch.manual_seed(0)
    E = torch.rand(2, 1, 512, 512)
    H = torch.rand(2, 1, 512, 512)
    m = O.all_metrics(E, H, 8)
    np.testing.assert_allclose(m["psnr"].numpy(), [7.7798634222, 7.7677034170], atol=1e-9)
    np.testing.assert_allclose(m["mse"].numpy(), [10841.6159436785, 10872.0144055671], atol=1e-8)
    np.testing.assert_allclose(m["nrmse"].numpy(), [0.4083258068, 0.4088978520], atol=1e-9)
    np.testing.assert_allclose(m["ssim"].numpy(), [0.0053607342, 0.0029760525], atol=1e-6)
    np.testing.assert_allclose((m["psnr_y"] - m["psnr"]).numpy(), 1.3219215, atol=1e-6)
    same = O.all_metrics(H, H, 8)
    np.testing.assert_allclose(same["psnr"].numpy(), 498.1308036087, atol=1e-6)
    np.testing.assert_allclose(same["ssim"].numpy(), 1.0, atol=1e-6)
    g = O.gaussian_window()
    assert abs(float(g[5, 5]) - 0.07076223939657211) < 1e-8
    assert abs(float(g[0, 0]) - 1.057566123563447e-06) < 1e-11


@pytest.mark.parametrize("name", ["swinir_tiny_direct", "swinir_tiny_ps", "swinir_tiny_3conv", "swinir_tiny_nearest"])
def test_swinir_tiny_against_reference_golden(name):
    z, sd, cfgd = load_npz(name)
    cfg = O.SwinIRCfg(**cfgd)
    i = 0
    while f"x{i}" in z.files:
        y = O.swinir_forward(sd, cfg, torch.from_numpy(z[f"x{i}"]))
        ref = torch.from_numpy(z[f"y{i}"])
        assert y.shape == ref.shape
        assert float((y - ref).abs().max()) < 2e-5
        # arithmetic contract of the CUDA path (bf16 operands) stays inside the 2e-3 budget
        yb = O.swinir_forward(sd, cfg, torch.from_numpy(z[f"x{i}"]), emulate_bf16=True)
        assert float((yb - ref).abs().max()) < 2e-3
        i += 1
    assert i >= 1


def test_edsr_tiny_against_reference_golden():
    z, sd, cfgd = load_npz("edsr_tiny")
    cfg = O.EDSRCfg(**cfgd)
    y = O.edsr_forward(sd, cfg, torch.from_numpy(z["x0"]))
    assert float((y - torch.from_numpy(z["y0"])).abs().max()) < 2e-5


def test_fullsize_samples_against_reference_golden():
    gold = json.load(open(os.path.join(G, "fullsize_samples.json")))
    jobs = [("cfg1_light_x2_64", T.cfg_light_x2(), (4, 64, 64), 101),
            ("cfg1_light_x2_72", T.cfg_light_x2(), (2, 72, 72), 101)]
    for name, cfg, (B, h, w), seed in jobs:
        sd = T.swinir_state_dict(cfg, seed=seed)
        y = O.swinir_forward(sd, cfg, T.synthetic_lr(B, h, w, seed))
        g = gold[name]
        assert list(y.shape) == g["shape"]
        got = y.reshape(-1)[torch.tensor(g["idx"])].double().numpy()
        np.testing.assert_allclose(got, np.array(g["val"]), atol=3e-5, rtol=0)
    cfg = T.cfg_edsr_x4()
    sd = T.edsr_state_dict(cfg, seed=102)
    y = O.edsr_forward(sd, cfg, T.synthetic_lr(2, 64, 64, 102))
    g = gold["cfg2_edsr_x4_64"]
    got = y.reshape(-1)[torch.tensor(g["idx"])].double().numpy()
    np.testing.assert_allclose(got, np.array(g["val"]), atol=3e-5, rtol=0)


def test_flop_counters_match_survey():
    assert abs(O.swinir_flops(T.cfg_classical(8), 64, 64) / 1e9 - 126.489) < 1e-3
    assert abs(O.swinir_flops(T.cfg_classical(8), 72, 72) / 1e9 - 160.088) < 1e-3
    assert abs(O.swinir_flops(T.cfg_light_x2(), 64, 64) / 1e9 - 8.521) < 1e-3
    assert abs(O.swinir_flops(T.cfg_classical(4), 64, 64) / 1e9 - 106.935) < 1e-3
    assert abs(O.swinir_flops(T.cfg_classical(2), 128, 128) / 1e9 - 408.188) < 1e-3
    assert abs(O.edsr_flops(T.cfg_edsr_x4(), 64, 64) / 1e9 - 16.086) < 1e-3


@pytest.mark.skipif(not R.available(), reason="live reference not present on this machine")
def test_state_dict_layout_and_forward_match_live_reference():
    from tests.golden.make_golden import build_ref_swinir
    cfg = O.SwinIRCfg(upscale=8, in_chans=1, img_size=16, depths=[2, 2], embed_dim=180,
                      num_heads=[6, 6], mlp_ratio=2, upsampler="pixelshuffle")
    sd = T.swinir_state_dict(cfg, seed=3)
    net = build_ref_swinir(cfg, sd)           # strict=True load: key names + shapes agree
    ref_sd = net.state_dict()
    assert set(ref_sd) == set(sd)
    for k in sd:
        assert ref_sd[k].shape == sd[k].shape and ref_sd[k].dtype == sd[k].dtype, k
    x = T.synthetic_lr(1, 24, 32, 5)
    with torch.no_grad():
        ref = net(x)
    assert float((O.swinir_forward(sd, cfg, x) - ref).abs().max()) < 2e-5
    # ctor trap (network_swinir.py:232-236): img_size <= window_size disables the shift
    cfg2 = O.SwinIRCfg(upscale=8, in_chans=1, img_size=8, depths=[2], embed_dim=60,
                       num_heads=[6], mlp_ratio=2, upsampler="pixelshuffledirect")
    sd2 = T.swinir_state_dict(cfg2, seed=4)
    assert not any(k.endswith("attn_mask") for k in sd2)
    net2 = build_ref_swinir(cfg2, sd2)
    with torch.no_grad():
        ref2 = net2(x)
    assert float((O.swinir_forward(sd2, cfg2, x) - ref2).abs().max()) < 2e-5


def test_bicubic_baseline_oracle_matches_reference_golden():
    """oracle.interpolate_baseline vs F.interpolate(bicubic, antialias=True) + clamp, the body of the
    reference's Interpolate.forward (tests/golden/make_bicubic_golden.py)."""
    import numpy as np
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "bicubic_baseline.npz"))
    for s in (2, 4, 8):
        got = O.interpolate_baseline(z[f"x{s}"], s)
        assert got.shape == z[f"y{s}"].shape
        assert float(np.abs(got - z[f"y{s}"]).max()) < 1e-6
        assert got.min() >= 0.0 and got.max() <= 1.0


def test_oracle_matches_round2_reference_goldens():
    """The oracle restatement against the round-2 fixtures of the unmodified reference that are small enough for the
    CPU suite: img_size == window_size (shift disabled at construction) and the large-magnitude stress network."""
    gold = json.load(open(os.path.join(T.GOLDEN, "fullsize_r2.json")))
    for name, cfg, shape, seed, gain in [("swinir_imgsize8", T.cfg_imgsize8(), (2, 16, 24), 106, 1.0),
                                         ("swinir_stress_large_weights", T.cfg_stress(), (2, 24, 32), 107, T.STRESS_GAIN)]:
        sd = T.swinir_state_dict(cfg, seed=seed)
        if gain != 1.0:
            sd = T.stress_state_dict(sd, gain)
        y = O.swinir_forward(sd, cfg, T.synthetic_lr(*shape, seed))
        gd = gold[name]
        assert list(y.shape) == gd["shape"]
        got = y.reshape(-1)[torch.tensor(gd["idx"])].double().numpy()
        assert np.abs(got - np.array(gd["val"])).max() <= 2e-6 * max(1.0, gd["absmax"]), name
