"""GPU parity tests (run on the B200 box: `pytest -m gpu`).  Everything goes through the C ABI
of libsrk.so; the checker is the CPU oracle (oracle/sr_oracle.py) and the golden vectors the
unmodified reference produced (tests/golden/)."""
import ctypes as C
import json
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from tests import common as T
from oracle import sr_oracle as O

pytestmark = pytest.mark.gpu

ENGINES = os.environ.get("SRK_TEST_ENGINES", "tcgen05,mma_sync").split(",")
DEV = "cuda:0"


@pytest.fixture(scope="module")
def L():
    from sr_caco_2_b200 import _lib
    _lib.load()
    _lib.check(_lib.load().srk_check_device(0))
    return _lib


@pytest.fixture(autouse=True)
def _sync_and_restore_engine():
    yield
    torch.cuda.synchronize()
    from sr_caco_2_b200 import _lib
    _lib.set_engine(ENGINES[0])


def load_npz(name):
    z = np.load(os.path.join(T.GOLDEN, name + ".npz"))
    sd = {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd::")}
    return z, sd, json.loads(bytes(z["cfg"]).decode())


def make_swinir(cfg, sd):
    from sr_caco_2_b200 import SwinIR
    net = SwinIR(upscale=cfg.upscale, in_chans=cfg.in_chans, img_size=cfg.img_size,
                 window_size=cfg.window_size, img_range=cfg.img_range, depths=cfg.depths,
                 embed_dim=cfg.embed_dim, num_heads=cfg.num_heads, mlp_ratio=cfg.mlp_ratio,
                 upsampler=cfg.upsampler, resi_connection=cfg.resi_connection)
    net.load_state_dict(sd, strict=True)
    return net.to(DEV).eval()


# ------------------------------------------------------------------------------------------
# index maps: bit exact
# ------------------------------------------------------------------------------------------
def _index_map(L, kind, H, W, shift, a, n):
    out = torch.empty(n, dtype=torch.int32, device=DEV)
    L.check(L.load().srk_index_map(kind, H, W, shift, a, L.ptr(out), L.stream_ptr()))
    return out.cpu().long()


@pytest.mark.parametrize("HW", [(16, 16), (64, 64), (72, 72), (128, 136), (24, 40)])
def test_index_maps_bit_exact(L, HW):
    H, W = HW
    for shift in (0, 4):
        got = _index_map(L, 0, H, W, shift, 0, H * W)
        assert torch.equal(got, O.window_gather_map(H, W, 8, shift))
        got = _index_map(L, 1, H, W, shift, 0, H * W * 64).view(-1, 64, 64)
        ref = (O.shift_attention_mask(H, W, 8, 4) != 0).long() if shift else torch.zeros_like(got)
        assert torch.equal(got, ref)
    assert torch.equal(_index_map(L, 2, 0, 0, 0, 0, 4096).view(64, 64), O.relative_position_index(8))
    for (Cc, h, w, r) in [(1, 4, 4, 2), (2, 3, 5, 2), (1, 4, 4, 8), (64, 6, 6, 2)]:
        got = _index_map(L, 3, h, w, r, Cc, Cc * h * r * w * r)
        assert torch.equal(got, O.pixel_shuffle_map(Cc, h, w, r).reshape(-1))


# ------------------------------------------------------------------------------------------
# metrics
# ------------------------------------------------------------------------------------------
def _inputs_for(name):
    from tests.test_oracle import _inputs_for as f
    return f({"name": name})


def test_metrics_against_reference_golden(L):
    from sr_caco_2_b200 import utils_image as UI
    cases = json.load(open(os.path.join(T.GOLDEN, "metrics_kat.json")))
    by_base = {}
    for c in cases:
        by_base.setdefault((c["name"].split("_roi")[0], c["border"]), []).append(c)
    for (base, border), group in by_base.items():
        E, H = _inputs_for(base)
        ths = sorted({c["roi_th"] for c in group if c["roi_th"] is not None})
        m = UI.compute_metrics(E.to(DEV), H.to(DEV), border, ths, check=False)
        raw = m["raw"].cpu()
        for c in group:
            v = 0 if c["roi_th"] is None else 1 + ths.index(c["roi_th"])
            for i, k in enumerate(("psnr", "mse", "nrmse", "ssim", "psnr_y")):
                tol = dict(psnr=1e-9, mse=1e-9, nrmse=1e-12, ssim=2e-5, psnr_y=1e-5)[k]
                np.testing.assert_allclose(raw[:, v, i].numpy(), np.array(c[k]), rtol=tol, atol=tol,
                                           err_msg=f"{c['name']} {k}")


def test_metric_function_shims_match_oracle(L):
    from sr_caco_2_b200 import utils_image as UI
    E, H = T.synthetic_pair(3, 96, 120, 31)
    e8, h8 = O.quantize_u8f(E), O.quantize_u8f(H)
    roi = (h8 >= 6).float()
    for r in (None, roi):
        rg = None if r is None else r.to(DEV)
        for border in (0, 3):
            np.testing.assert_allclose(UI.mbatch_gpu_calculate_psnr(e8.to(DEV), h8.to(DEV), border, rg).cpu(),
                                       O.psnr(e8, h8, border, r), rtol=1e-10)
            np.testing.assert_allclose(UI.mbatch_gpu_calculate_mse(e8.to(DEV), h8.to(DEV), border, rg).cpu(),
                                       O.mse(e8, h8, border, r), rtol=1e-10)
            np.testing.assert_allclose(UI.mbatch_gpu_calculate_nrmse(e8.to(DEV), h8.to(DEV), border, rg).cpu(),
                                       O.nrmse(e8, h8, border, r), rtol=1e-10)
            got = UI.mbatch_gpu_calculate_ssim(e8.to(DEV), h8.to(DEV), border, rg)
            assert got.dtype == torch.float32
            np.testing.assert_allclose(got.cpu(), O.ssim(e8, h8, border, r), atol=2e-5)
    # non-integer inputs (the PSNR_Y operands of the reference) go through the same fp64 path
    a, b = e8 * 0.859 + 16, h8 * 0.859 + 16
    np.testing.assert_allclose(UI.mbatch_gpu_calculate_psnr(a.to(DEV), b.to(DEV), 2).cpu(), O.psnr(a, b, 2), rtol=1e-10)
    assert torch.equal(UI.tensor2uint82float(E.to(DEV)).cpu(), e8)
    with pytest.raises(ValueError):       # 11x11 window does not fit (utils_image.py:1044)
        UI.mbatch_gpu_calculate_ssim(e8[..., :12, :12].to(DEV), h8[..., :12, :12].to(DEV), border=1)
    with pytest.raises(AssertionError):   # range assert of the reference (:1164-1172)
        UI.mbatch_gpu_calculate_ssim((e8 + 300).to(DEV), h8.to(DEV))
    with pytest.raises(AssertionError):
        UI.mbatch_gpu_calculate_psnr(e8.to(DEV), h8[:2].to(DEV))


def test_metrics_full_size_properties(L):
    """BASELINE full size (B=32, 512x512): identities that do not need the CPU oracle."""
    from sr_caco_2_b200 import utils_image as UI
    g = torch.Generator(device=DEV).manual_seed(5)
    Hh = (torch.rand(32, 1, 512, 512, device=DEV, generator=g) * 255).round() / 255
    m = UI.compute_metrics(Hh, Hh, 8, (4, 7, 10))
    assert torch.allclose(m["psnr"], torch.full_like(m["psnr"], 498.1308036087), atol=1e-6)
    assert torch.allclose(m["ssim"], torch.ones_like(m["ssim"]), atol=1e-6)
    assert float(m["mse"].abs().max()) == 0.0 and float(m["nrmse"].abs().max()) == 0.0
    E = (Hh + 0.1 * torch.randn(Hh.shape, device=DEV, generator=g))
    m1 = UI.compute_metrics(E, Hh, 8, (4, 7, 10))
    m2 = UI.compute_metrics(E, Hh, 8, (4, 7, 10))
    assert torch.equal(m1["raw"][..., :3], m2["raw"][..., :3])           # integer sums: deterministic
    assert torch.allclose(m1["raw"], m2["raw"], rtol=1e-12, atol=1e-12)
    # PSNR_Y - PSNR = 20 log10(255/219) for gray images
    assert torch.allclose(m1["psnr_y"] - m1["psnr"], torch.full_like(m1["psnr"], 1.3219215), atol=1e-5)
    # batch-permutation equivariance + per-image independence
    perm = torch.randperm(32, device=DEV)
    m3 = UI.compute_metrics(E[perm], Hh[perm], 8, (4, 7, 10))
    assert torch.allclose(m3["raw"], m1["raw"][perm], rtol=1e-12, atol=1e-12)
    sub = O.all_metrics(E[:2].cpu(), Hh[:2].cpu(), 8)
    np.testing.assert_allclose(m1["psnr"][:2].cpu(), sub["psnr"], rtol=1e-10)
    np.testing.assert_allclose(m1["ssim"][:2].cpu(), sub["ssim"].double(), atol=2e-6)


@pytest.mark.parametrize("shape", [(1, 11, 11, 0), (2, 12, 139, 0), (3, 140, 128, 0), (2, 129, 131, 1), (1, 300, 250, 3),
                                   (2, 534, 260, 8), (5, 64, 64, 8), (1, 1040, 47, 2)])
def test_metrics_streaming_kernel_geometries(L, shape):
    """The row-streaming hot-path kernel against (a) the round-1 tile kernel: integer sums (PSNR / MSE / NRMSE of
    every ROI variant) bit-identical, SSIM within fp32 summation noise; (b) the oracle.  Geometries: a single 11 x 11
    window, widths that leave a ragged / one-column last strip, heights that need several row segments (> 244 map
    rows), 8 ROI thresholds (9 buckets: the packed counters of the kernel)."""
    from sr_caco_2_b200 import utils_image as UI
    B, Hh, Ww, border = shape
    E, H = T.synthetic_pair(B, Hh, Ww, 900 + Hh)
    ths = (1, 3, 5, 8, 13, 40, 90, 200)
    lib = L.load()
    try:
        lib.srk_metrics_use_tile_kernel(1)
        tile = UI.compute_metrics(E.to(DEV), H.to(DEV), border, ths, check=False)["raw"].cpu()
    finally:
        lib.srk_metrics_use_tile_kernel(0)
    got = UI.compute_metrics(E.to(DEV), H.to(DEV), border, ths, check=False)["raw"].cpu()
    h8 = UI.compute_metrics(E.to(DEV), (H * 255).round().clamp(0, 255).to(torch.uint8).to(DEV), border, ths, check=False)["raw"].cpu()
    for i in (0, 1, 2, 4):                                  # psnr, mse, nrmse, psnr_y: exact integer sums underneath
        assert torch.equal(got[..., i], tile[..., i]), i
        assert torch.equal(h8[..., i], tile[..., i]), i
    assert float((got[..., 3] - tile[..., 3]).abs().max()) < 1e-5          # a single-window map has no averaging
    assert float((h8[..., 3] - got[..., 3]).abs().max()) == 0.0
    om = O.all_metrics(E, H, border)
    np.testing.assert_allclose(got[:, 0, 0], om["psnr"], rtol=1e-10)
    np.testing.assert_allclose(got[:, 0, 2], om["nrmse"], rtol=1e-10)
    np.testing.assert_allclose(got[:, 0, 3], om["ssim"].double(), atol=2e-5)
    for v, th in ((2, 3), (6, 40)):
        orm = O.all_metrics(E, H, border, th)
        np.testing.assert_allclose(got[:, v, 0], orm["psnr"], rtol=1e-10)
        np.testing.assert_allclose(got[:, v, 2], orm["nrmse"], rtol=1e-10)
        np.testing.assert_allclose(got[:, v, 3], orm["ssim"].double(), atol=2e-5)


@pytest.mark.parametrize("case", [(2, 97, 133, 2, ()), (1, 300, 41, 0, (7,)), (1, 23, 501, 5, (250,)), (2, 50, 50, 4, (0, 255))])
def test_metrics_streaming_kernel_threshold_edge_cases(L, case):
    """No ROI threshold at all, a single one, thresholds at the ends of the level range (a ROI that is everything /
    almost nothing): streaming kernel == tile kernel on the integer sums, SSIM to fp32 noise, PSNR / SSIM == oracle."""
    from sr_caco_2_b200 import utils_image as UI
    B, Hh, Ww, border, ths = case
    g = torch.Generator().manual_seed(3 + Hh)
    H = (torch.rand(B, 1, Hh, Ww, generator=g) * 255).round() / 255
    E = H + 0.03 * torch.randn(B, 1, Hh, Ww, generator=g)
    lib = L.load()
    try:
        lib.srk_metrics_use_tile_kernel(1)
        tile = UI.compute_metrics(E.to(DEV), H.to(DEV), border, ths, check=False)["raw"].cpu()
    finally:
        lib.srk_metrics_use_tile_kernel(0)
    got = UI.compute_metrics(E.to(DEV), H.to(DEV), border, ths, check=False)["raw"].cpu()
    for i in (0, 1, 2, 4):
        assert torch.equal(got[..., i], tile[..., i]), i
    assert float((got[..., 3] - tile[..., 3]).abs().max()) < 2e-6
    om = O.all_metrics(E, H, border)
    np.testing.assert_allclose(got[:, 0, 0], om["psnr"], rtol=1e-10)
    np.testing.assert_allclose(got[:, 0, 3], om["ssim"].double(), atol=2e-6)
    for v, th in enumerate(ths):
        orm = O.all_metrics(E, H, border, th)
        np.testing.assert_allclose(got[:, 1 + v, 0], orm["psnr"], rtol=1e-10)
        np.testing.assert_allclose(got[:, 1 + v, 3], orm["ssim"].double(), atol=2e-6)


# ------------------------------------------------------------------------------------------
# building blocks
# ------------------------------------------------------------------------------------------
def test_layernorm_and_window_gather(L):
    lib = L.load()
    g = torch.Generator().manual_seed(1)
    for (C_, ld, H, W) in [(180, 192, 16, 24), (60, 64, 8, 8)]:
        B = 2
        M = B * H * W
        x = torch.zeros(M, ld)
        x[:, :C_] = torch.randn(M, C_, generator=g) * 3 + 1
        gam, bet = torch.randn(C_, generator=g), torch.randn(C_, generator=g)
        ref = F.layer_norm(x[:, :C_], (C_,), gam, bet, 1e-5)
        xd, gd, bd = x.to(DEV), gam.to(DEV), bet.to(DEV)
        for shift in (-1, 0, 4):
            out = torch.full((M, ld), 7.0, dtype=torch.bfloat16, device=DEV)
            x32 = torch.full((M, ld), 7.0, device=DEV)
            L.check(lib.srk_layernorm(L.ptr(xd), ld, M, C_, L.ptr(gd), L.ptr(bd), 1e-5, L.ptr(out), ld,
                                      L.SRK_BF16, L.ptr(x32), H, W, shift, L.stream_ptr()))
            exp = ref
            if shift >= 0:
                gm = O.window_gather_map(H, W, 8, shift)
                idx = (torch.arange(B)[:, None] * H * W + gm[None, :]).reshape(-1)
                exp = ref[idx]
            assert torch.equal(out[:, :C_].float().cpu(), exp.bfloat16().float()) or \
                float((out[:, :C_].float().cpu() - exp).abs().max()) < 0.04
            assert float((out[:, :C_].float().cpu() - exp).abs().max()) < 0.04
            assert float(out[:, C_:].float().abs().max()) == 0.0
            assert float((x32[:, :C_].cpu() - ref).abs().max()) < 1e-5
        out = torch.empty(M, ld, dtype=torch.float16, device=DEV)       # cast mode
        L.check(lib.srk_layernorm(L.ptr(xd), ld, M, C_, None, None, 1e-5, L.ptr(out), ld, L.SRK_FP16,
                                  None, H, W, -1, L.stream_ptr()))
        assert torch.equal(out.cpu(), x.half())


def test_conv_in_and_conv_out(L):
    lib = L.load()
    g = torch.Generator().manual_seed(2)
    B, h, w, C_ = 2, 13, 18, 60
    H, W, ld = 16, 24, 64
    x = torch.rand(B, 1, h, w, generator=g)
    wt, bs = torch.randn(C_, 1, 3, 3, generator=g), torch.randn(C_, generator=g)
    xp = F.pad(x, (0, W - w, 0, H - h), mode="reflect") * 0.5
    ref = F.conv2d(xp, wt, bs, padding=1).permute(0, 2, 3, 1).reshape(-1, C_)
    o32 = torch.full((B * H * W, ld), 3.0, device=DEV)
    o16 = torch.full((B * H * W, ld), 3.0, dtype=torch.float16, device=DEV)
    xd, wd, bd = x.to(DEV), wt.reshape(C_, 9).contiguous().to(DEV), bs.to(DEV)
    L.check(lib.srk_conv_in(L.ptr(xd), B, h, w, H, W, 0.5, L.ptr(wd), L.ptr(bd), C_, L.ptr(o32), ld,
                            L.ptr(o16), ld, L.SRK_FP16, L.stream_ptr()))
    assert float((o32[:, :C_].cpu() - ref).abs().max()) < 1e-5
    assert float(o32[:, C_:].abs().max()) == 0.0
    assert float((o16[:, :C_].float().cpu() - ref).abs().max()) < 5e-3
    # conv_out with crop
    Cin, lda, Hc, Wc = 64, 64, 29, 41
    a = torch.randn(B, 32, 48, lda, generator=g).half()
    wo, bo = (torch.randn(1, Cin, 3, 3, generator=g) * 0.1).half(), 0.3
    refo = (F.conv2d(a.float().permute(0, 3, 1, 2), wo.float(), torch.tensor([bo]), padding=1) * 2.0)[:, :, :Hc, :Wc]
    y = torch.zeros(B, 1, Hc, Wc, device=DEV)
    wk = wo[0].permute(1, 2, 0).reshape(9, Cin).contiguous().to(DEV)
    ad = a.to(DEV)
    L.check(lib.srk_conv_out(L.ptr(ad), lda, B, 32, 48, Cin, L.ptr(wk), bo, 2.0, L.ptr(y), Hc, Wc,
                             L.stream_ptr()))
    assert float((y.cpu() - refo).abs().max()) < 2e-4


def _gemm(L, **kw):
    g = L.GemmArgs()
    g.res_scale, g.win_shift, g.ln_win_shift = 1.0, -1, -1
    for k, v in kw.items():
        setattr(g, k, L.ptr(v) if isinstance(v, torch.Tensor) else v)
    L.check(L.load().srk_gemm(C.byref(g), L.stream_ptr()))


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("dtype", ["bf16", "fp16"])
def test_gemm_rows_epilogues(L, engine, dtype):
    L.set_engine(engine)
    td, code = (torch.bfloat16, L.SRK_BF16) if dtype == "bf16" else (torch.float16, L.SRK_FP16)
    g = torch.Generator().manual_seed(3)
    H, W, B = 16, 24, 2
    M = B * H * W
    for (N, K) in [(192, 192), (576, 192), (384, 192), (192, 384), (64, 64), (320, 64), (128, 128)]:
        A = (torch.randn(M, K, generator=g)).to(td)
        Wt = (torch.randn(N, K, generator=g) * 0.05).to(td)
        bias = torch.randn(N, generator=g)
        ref = A.float() @ Wt.float().t() + bias
        Ad, Wd, bd = A.to(DEV), Wt.to(DEV), bias.to(DEV)
        # (1) bias + GELU -> 16-bit rows
        o16 = torch.zeros(M, N, dtype=td, device=DEV)
        _gemm(L, A=Ad, a_mode=L.A_ROWS, lda=K, nB=B, H=H, W=W, Wt=Wd, M=M, N=N, K=K, dtype=code, bias=bd,
              act=L.ACT_GELU, out16=o16, ld16=N, out16_dtype=code, out16_mode=L.O16_ROWS)
        exp = F.gelu(ref)
        err = (o16.float().cpu() - exp).abs().max()
        assert float(err) < 0.02 * max(1.0, float(exp.abs().max())), (N, K, float(err))
        # (2) bias + residual with window reverse + roll back -> fp32, plus 16-bit cast
        for shift in (-1, 4):
            res = torch.randn(M, N, generator=g)
            o32 = torch.zeros(M, N, device=DEV)
            o16 = torch.zeros(M, N, dtype=torch.float16, device=DEV)
            _gemm(L, A=Ad, a_mode=L.A_ROWS, lda=K, nB=B, H=H, W=W, Wt=Wd, M=M, N=N, K=K, dtype=code, bias=bd,
                  act=L.ACT_NONE, res=res.to(DEV), res_scale=0.5, out32=o32, ld32=N, win_shift=shift,
                  out16=o16, ld16=N, out16_dtype=L.SRK_FP16, out16_mode=L.O16_ROWS)
            if shift >= 0:
                gm = O.window_gather_map(H, W, 8, shift)
                idx = (torch.arange(B)[:, None] * H * W + gm[None, :]).reshape(-1)
                exp32 = torch.empty(M, N)
                exp32[idx] = ref * 0.5 + res[idx]
                exp16 = ref * 0.5 + res[idx]           # 16-bit copy stays in GEMM-row order
            else:
                exp32 = ref * 0.5 + res
                exp16 = exp32
            assert float((o32.cpu() - exp32).abs().max()) < 2e-3 * max(1.0, float(ref.abs().max()))
            assert float((o16.float().cpu() - exp16).abs().max()) < 0.02 * max(1.0, float(ref.abs().max()))
    # ragged M (not a multiple of the 128-row tile)
    M2, N, K = 200, 64, 128
    A = torch.randn(M2, K, generator=g).to(td)
    Wt = (torch.randn(N, K, generator=g) * 0.05).to(td)
    bias = torch.randn(N, generator=g)
    o32 = torch.zeros(M2 + 56, N, device=DEV)
    _gemm(L, A=A.to(DEV), a_mode=L.A_ROWS, lda=K, Wt=Wt.to(DEV), M=M2, N=N, K=K, dtype=code, bias=bias.to(DEV),
          out32=o32, ld32=N)
    assert float((o32[:M2].cpu() - (A.float() @ Wt.float().t() + bias)).abs().max()) < 2e-3
    assert float(o32[M2:].abs().max()) == 0.0


@pytest.mark.parametrize("engine", ENGINES)
def test_gemm_conv3x3_pixelshuffle_and_image_epilogues(L, engine):
    L.set_engine(engine)
    g = torch.Generator().manual_seed(4)
    B, H, W = 2, 10, 13
    for (Cin, Cout) in [(64, 64), (192, 192), (64, 256), (192, 64)]:
        x = torch.randn(B, Cin, H, W, generator=g).half()
        wt = (torch.randn(Cout, Cin, 3, 3, generator=g) * 0.05).half()
        bias = torch.randn(Cout, generator=g)
        ref = F.conv2d(x.float(), wt.float(), bias, padding=1)                      # (B,Cout,H,W)
        a = x.permute(0, 2, 3, 1).contiguous().to(DEV)                               # NHWC
        from sr_caco_2_b200 import packing as P
        ps = 2 if Cout == 256 else 0
        wk, bk = P.pack_conv3x3(wt.float(), bias, Cin, Cout, L.SRK_FP16, pixel_shuffle_r=ps)
        M = B * H * W
        if ps:
            o16 = torch.zeros(B, 2 * H, 2 * W, 64, dtype=torch.float16, device=DEV)
            _gemm(L, A=a, a_mode=L.A_CONV3X3, lda=Cin, nB=B, H=H, W=W, Wt=wk.to(DEV), M=M, N=Cout, K=9 * Cin,
                  dtype=L.SRK_FP16, bias=bk.to(DEV), act=L.ACT_NONE, out16=o16, ld16=64,
                  out16_dtype=L.SRK_FP16, out16_mode=L.O16_PIXSHUF2)
            exp = F.pixel_shuffle(ref, 2).permute(0, 2, 3, 1)
            assert float((o16.float().cpu() - exp).abs().max()) < 0.02 * max(1.0, float(exp.abs().max()))
        else:
            res = torch.randn(M, Cout, generator=g)
            o32 = torch.zeros(M, Cout, device=DEV)
            o16 = torch.zeros(M, Cout, dtype=torch.float16, device=DEV)
            _gemm(L, A=a, a_mode=L.A_CONV3X3, lda=Cin, nB=B, H=H, W=W, Wt=wk.to(DEV), M=M, N=Cout, K=9 * Cin,
                  dtype=L.SRK_FP16, bias=bk.to(DEV), act=L.ACT_LRELU, res=res.to(DEV), out32=o32, ld32=Cout,
                  out16=o16, ld16=Cout, out16_dtype=L.SRK_FP16, out16_mode=L.O16_ROWS)
            exp = F.leaky_relu(ref, 0.01).permute(0, 2, 3, 1).reshape(M, Cout) + res
            assert float((o32.cpu() - exp).abs().max()) < 3e-3 * max(1.0, float(exp.abs().max()))
            assert float((o16.float().cpu() - exp).abs().max()) < 0.02 * max(1.0, float(exp.abs().max()))
    # pixelshuffle-direct image epilogue with crop (C -> s*s, s = 4)
    Cin, s = 64, 4
    x = torch.randn(B, Cin, H, W, generator=g).half()
    wt = (torch.randn(s * s, Cin, 3, 3, generator=g) * 0.05).half()
    bias = torch.randn(s * s, generator=g)
    wk, bk = P.pack_conv3x3(wt.float(), bias, Cin, 64, L.SRK_FP16)
    hc, wc = H * s - 5, W * s - 3
    img = torch.zeros(B, 1, hc, wc, device=DEV)
    _gemm(L, A=x.permute(0, 2, 3, 1).contiguous().to(DEV), a_mode=L.A_CONV3X3, lda=Cin, nB=B, H=H, W=W,
          Wt=wk.to(DEV), M=B * H * W, N=64, K=9 * Cin, dtype=L.SRK_FP16, bias=bk.to(DEV), img=img, img_s=s,
          img_scale=0.5, img_hc=hc, img_wc=wc)
    exp = F.pixel_shuffle(F.conv2d(x.float(), wt.float(), bias, padding=1), s)[:, :, :hc, :wc] * 0.5
    assert float((img.cpu() - exp).abs().max()) < 3e-3 * max(1.0, float(exp.abs().max()))


@pytest.mark.parametrize("geom", [(3, 40, 24, 64, 3), (2, 33, 17, 64, 3), (40, 64, 64, 64, 3), (2, 21, 30, 64, 5), (1, 16, 8, 64, 3), (2, 24, 24, 192, 3)])
def test_conv_halo_tiles_equal_per_tap_boxes(L, geom):
    """64-output-channel convs of the tcgen05 engine: ONE halo TMA box per channel block with the taps read as
    shifted-window UMMA descriptors (default) against one TMA box per tap (srk_gemm_conv_halo(0)) and against the
    fp32 convolution.  The two operand paths add the same products in a different order (channel-block major vs
    tap major): equal to fp32 accumulation noise.  The halo path is taken by the 64 -> 64 channel convs (the EDSR body,
    the folded 5x5 tail).  Geometries: ragged tiles in both directions (few tiles: streamed weights), enough tiles for
    resident weights, a 5x5 window (halo radius 2), a single partial tile; 192 input channels stay on per-tap boxes
    (both runs identical)."""
    if "tcgen05" not in ENGINES:
        pytest.skip("tcgen05 engine not under test")
    L.set_engine("tcgen05")
    B, H, W, Cin, kk = geom
    g = torch.Generator().manual_seed(77 + H)
    x = torch.randn(B, Cin, H, W, generator=g).half()
    wt = (torch.randn(64, Cin, kk, kk, generator=g) * 0.05).half()
    bias = torch.randn(64, generator=g)
    ref = F.leaky_relu(F.conv2d(x.float(), wt.float(), bias, padding=kk // 2), 0.01).permute(0, 2, 3, 1).reshape(-1, 64)
    wk = wt.float().permute(0, 2, 3, 1).reshape(64, kk * kk * Cin).half().contiguous().to(DEV)          # k = tap * Cin + c
    a = x.permute(0, 2, 3, 1).contiguous().to(DEV)
    M = B * H * W
    res = torch.randn(M, 64, generator=g)
    outs = []
    lib = L.load()
    try:
        for halo in (1, 0):
            lib.srk_gemm_conv_halo(halo)
            o32 = torch.zeros(M, 64, device=DEV)
            o16 = torch.zeros(M, 64, dtype=torch.float16, device=DEV)
            _gemm(L, A=a, a_mode=L.A_CONV3X3, conv_k=kk, lda=Cin, nB=B, H=H, W=W, Wt=wk, M=M, N=64, K=kk * kk * Cin,
                  dtype=L.SRK_FP16, bias=bias.to(DEV), act=L.ACT_LRELU, res=res.to(DEV), out32=o32, ld32=64,
                  out16=o16, ld16=64, out16_dtype=L.SRK_FP16, out16_mode=L.O16_ROWS)
            torch.cuda.synchronize()
            outs.append((o32.cpu(), o16.float().cpu()))
    finally:
        lib.srk_gemm_conv_halo(1)
    exp = ref + res
    scale = max(1.0, float(exp.abs().max()))
    assert float((outs[0][0] - exp).abs().max()) < 3e-3 * scale
    assert float((outs[0][0] - outs[1][0]).abs().max()) < 2e-5 * scale
    assert float((outs[0][1] - outs[1][1]).abs().max()) < 0.01 * scale
    # 16-bit-only output (the E_O16 epilogue the EDSR body uses)
    try:
        for halo in (1, 0):
            lib.srk_gemm_conv_halo(halo)
            o16 = torch.zeros(M, 64, dtype=torch.float16, device=DEV)
            _gemm(L, A=a, a_mode=L.A_CONV3X3, conv_k=kk, lda=Cin, nB=B, H=H, W=W, Wt=wk, M=M, N=64, K=kk * kk * Cin,
                  dtype=L.SRK_FP16, bias=bias.to(DEV), act=L.ACT_RELU, out16=o16, ld16=64, out16_dtype=L.SRK_FP16,
                  out16_mode=L.O16_ROWS)
            torch.cuda.synchronize()
            outs.append(o16.float().cpu())
    finally:
        lib.srk_gemm_conv_halo(1)
    exp16 = F.relu(F.conv2d(x.float(), wt.float(), bias, padding=kk // 2)).permute(0, 2, 3, 1).reshape(-1, 64)
    assert float((outs[2] - exp16).abs().max()) < 0.02 * max(1.0, float(exp16.abs().max()))
    assert float((outs[2] - outs[3]).abs().max()) < 0.01 * max(1.0, float(exp16.abs().max()))


@pytest.mark.parametrize("engine", ENGINES)
def test_gemm_conv5x5_image_epilogue(L, engine):
    """conv_k = 5 (implicit GEMM over a 5x5 window, zero padding 2) with the image epilogue, s = 8."""
    L.set_engine(engine)
    g = torch.Generator().manual_seed(14)
    B, H, W, Cin, s = 2, 11, 9, 64, 8
    x = torch.randn(B, Cin, H, W, generator=g).half()
    wt = (torch.randn(s * s, Cin, 5, 5, generator=g) * 0.03).half()
    bias = torch.randn(s * s, generator=g)
    wk = wt.float().permute(0, 2, 3, 1).reshape(s * s, 25 * Cin).half().contiguous()     # k = tap*Cin + c
    hc, wc = H * s - 3, W * s - 6
    img = torch.zeros(B, 1, hc, wc, device=DEV)
    _gemm(L, A=x.permute(0, 2, 3, 1).contiguous().to(DEV), a_mode=L.A_CONV3X3, conv_k=5, lda=Cin, nB=B, H=H, W=W,
          Wt=wk.to(DEV), M=B * H * W, N=64, K=25 * Cin, dtype=L.SRK_FP16, bias=bias.to(DEV), img=img, img_s=s,
          img_scale=0.25, img_hc=hc, img_wc=wc)
    exp = F.pixel_shuffle(F.conv2d(x.float(), wt.float(), bias, padding=2), s)[:, :, :hc, :wc] * 0.25
    assert float((img.cpu() - exp).abs().max()) < 3e-3 * max(1.0, float(exp.abs().max()))


@pytest.mark.parametrize("scale", [2, 4, 8])
def test_folded_tail_equals_conv_chain(L, scale):
    """packing.fold_tail + 5x5 conv GEMM + srk_tail_border vs the upsampler as the reference runs it:
    [Conv3x3(64->256) + PixelShuffle(2)] x log2(s), Conv3x3(64->1), fp32 torch on the same fp16 features
    (network_swinir.py:661-680, 868).  Every pixel, including the border ring, with a crop."""
    if "tcgen05" not in ENGINES:
        pytest.skip("tcgen05 engine not under test")
    L.set_engine("tcgen05")
    from sr_caco_2_b200 import packing as P
    g = torch.Generator().manual_seed(15 + scale)
    n = {2: 1, 4: 2, 8: 3}[scale]
    ups = [((torch.randn(256, 64, 3, 3, generator=g) * 0.04).to(DEV), (torch.randn(256, generator=g) * 0.1).to(DEV))
           for _ in range(n)]
    lw, lb = (torch.randn(1, 64, 3, 3, generator=g) * 0.05).to(DEV), (torch.randn(1, generator=g) * 0.1).to(DEV)
    for (B, H, W, ch, cw) in [(2, 24, 40, 0, 0), (3, 19, 17, 5, 3), (1, 3, 3, 0, 1)]:
        x = torch.randn(B, 64, H, W, generator=g).half().to(DEV)
        y = x.float()
        for w_, b_ in ups:
            y = F.pixel_shuffle(F.conv2d(y, w_, b_, padding=1), 2)
        hc, wc = H * scale - ch, W * scale - cw
        exp = F.conv2d(y, lw, lb, padding=1)[:, :, :hc, :wc] * 0.5
        fw, fb, bw, bb, wsc = P.fold_tail(ups, lw, lb, scale)
        feat = x.permute(0, 2, 3, 1).contiguous()
        img = torch.full((B, 1, hc, wc), 7.0, device=DEV)
        _gemm(L, A=feat, a_mode=L.A_CONV3X3, conv_k=5, lda=64, nB=B, H=H, W=W, Wt=fw, M=B * H * W, N=64, K=1600,
              dtype=L.SRK_FP16, bias=fb, img=img, img_s=scale, img_scale=0.5 / wsc, img_hc=hc, img_wc=wc)
        tf = L.TailFold(L.ptr(fw), L.ptr(fb), L.ptr(bw), L.ptr(bb), wsc)
        L.check(L.load().srk_tail_border(L.ptr(feat), B, H, W, scale, C.byref(tf), 0.5, L.ptr(img), hc, wc,
                                         L.stream_ptr()))
        err = float((img - exp).abs().max())
        assert err < 2e-3 * max(1.0, float(exp.abs().max())), (scale, B, H, W, err)


def test_folded_tail_network_equals_unfolded(L):
    """options = SRK_OPT_NO_FOLD_TAIL (upsampler convs one by one, fp16 intermediates) vs the folded tail, whole network."""
    if "tcgen05" not in ENGINES:
        pytest.skip("tcgen05 engine not under test")
    L.set_engine("tcgen05")
    cfg = O.SwinIRCfg(upscale=4, img_size=16, embed_dim=60, depths=[2, 2], num_heads=[6, 6], mlp_ratio=2.0,
                      upsampler="pixelshuffle")
    net = make_swinir(cfg, T.swinir_state_dict(cfg, 5))
    x = T.synthetic_lr(2, 21, 27, 3).to(DEV)
    y1 = net(x).clone()
    net.options = L.OPT_NO_FOLD_TAIL
    y0 = net(x).clone()
    net.options = 0
    assert y1.shape == y0.shape == (2, 1, 84, 108)
    assert float((y1 - y0).abs().max()) < 1.5e-3


@pytest.mark.parametrize("geom", [(180, 6, 24, 16, 3), (128, 4, 16, 16, 1), (180, 6, 64, 64, 9), (180, 6, 40, 24, 5)])
def test_fused_qkv_attention_equals_unfused(L, geom):
    """E_ATTN epilogue (qkv GEMM + window attention in one tcgen05 kernel) vs srk_gemm + srk_window_attention:
    the same attention unit, bit for bit."""
    if "tcgen05" not in ENGINES:
        pytest.skip("tcgen05 engine not under test")
    L.set_engine("tcgen05")
    Cc, nh, H, W, B = geom
    Cp = (Cc + 63) // 64 * 64
    d = Cc // nh
    assert d <= 32
    g = torch.Generator().manual_seed(12)
    M = B * H * W
    from sr_caco_2_b200 import packing as P
    wq = torch.randn(3 * Cc, Cc, generator=g) * 0.08
    bq = torch.randn(3 * Cc, generator=g) * 0.2
    nq = 3 * nh * 32
    wpk, bpk = P.pack_qkv(wq, bq, nh, d, 32, nq, Cp, L.SRK_BF16)
    A = torch.zeros(M, Cp); A[:, :Cc] = torch.randn(M, Cc, generator=g)
    table = (torch.randn(225, nh, generator=g) * 0.5).t().contiguous()
    Ad, wd, bd, td = A.bfloat16().to(DEV), wpk.to(DEV), bpk.to(DEV), table.to(DEV)
    for shift in (0, 4):
        qkv = torch.empty(M, nq, dtype=torch.bfloat16, device=DEV)
        ref = torch.empty(M, nh * 32, dtype=torch.bfloat16, device=DEV)
        _gemm(L, A=Ad, a_mode=L.A_ROWS, lda=Cp, nB=B, H=H, W=W, Wt=wd, M=M, N=nq, K=Cp, dtype=L.SRK_BF16, bias=bd,
              out16=qkv, ld16=nq, out16_dtype=L.SRK_BF16)
        L.check(L.load().srk_window_attention(L.ptr(qkv), nq, L.ptr(ref), nh * 32, L.ptr(td), B, H, W, nh, 32,
                                              d ** -0.5, shift, L.stream_ptr()))
        got = torch.full((M, nh * 32), 3.0, dtype=torch.bfloat16, device=DEV)
        _gemm(L, A=Ad, a_mode=L.A_ROWS, lda=Cp, nB=B, H=H, W=W, Wt=wd, M=M, N=nq, K=Cp, dtype=L.SRK_BF16,
              bias=bd, out16=got, ld16=nh * 32, out16_dtype=L.SRK_BF16, attn_table=td, attn_heads=nh,
              attn_scale=d ** -0.5, attn_shift=shift)
        assert torch.equal(got, ref), (shift, float((got.float() - ref.float()).abs().max()))


@pytest.mark.parametrize("geom", [(24, 16, 3), (64, 64, 5), (40, 24, 1), (8, 8, 3), (16, 8, 1)])
def test_attn_block_kernel_equals_unfused_sequence(L, geom):
    """srk_attn_block (LN1 rows -> qkv -> window attention -> proj + residual -> LN2 in ONE kernel) against the
    launch sequence it replaces, through the C ABI: fused qkv + attention GEMM, then the proj GEMM with the
    window_reverse + roll residual map and the fused LayerNorm.  The attention unit and the MMA operand order
    are the same, so x' agrees to fp32 rounding of the residual add and LN2 to one bf16 ulp.
    (8, 8, 3): 64-token images -> odd window counts, the last tile has one window; M % 128 != 0.)"""
    if "tcgen05" not in ENGINES:
        pytest.skip("tcgen05 engine not under test")
    L.set_engine("tcgen05")
    H, W, B = geom
    Cc, nh, Cp, d = 180, 6, 192, 30
    g = torch.Generator().manual_seed(21)
    M = B * H * W
    from sr_caco_2_b200 import packing as P
    wq = torch.randn(3 * Cc, Cc, generator=g) * 0.08
    bq = torch.randn(3 * Cc, generator=g) * 0.2
    wp = torch.randn(Cc, Cc, generator=g) * 0.08
    bp = torch.randn(Cc, generator=g) * 0.1
    nq = 3 * nh * 32
    wpk, bpk = P.pack_qkv(wq, bq, nh, d, 32, nq, Cp, L.SRK_BF16)
    wfb = P.fold_qkv_bias(wpk, bpk, Cc)
    whm = P.pack_qkv_heads(wfb, nh, 32)
    wpr = P.pack_proj(wp, nh, d, 32, Cp, nh * 32, L.SRK_BF16)
    bpr = P.pad_bias(bp, Cp)
    A = torch.zeros(M, Cp); A[:, :Cc] = torch.randn(M, Cc, generator=g); A[:, Cc:Cc + 2] = 1.0
    res = torch.zeros(M, Cp); res[:, :Cc] = torch.randn(M, Cc, generator=g)
    gam, bet = 1.0 + 0.1 * torch.randn(Cc, generator=g), 0.1 * torch.randn(Cc, generator=g)
    table = (torch.randn(225, nh, generator=g) * 0.5).t().contiguous()
    dv = lambda t: t.to(DEV)
    Ad, td = dv(A.bfloat16()), dv(table)
    wfbd, whmd, wprd, bprd, gd, bd = dv(wfb), dv(whm), dv(wpr), dv(bpr), dv(gam), dv(bet)
    for shift in (0, 4):
        if shift and (H <= 8 or W <= 8):
            continue                                          # the reference never shifts a one-window axis
        ao = torch.empty(M, nh * 32, dtype=torch.bfloat16, device=DEV)
        _gemm(L, A=Ad, a_mode=L.A_ROWS, lda=Cp, nB=B, H=H, W=W, Wt=wfbd, M=M, N=nq, K=Cp, dtype=L.SRK_BF16,
              out16=ao, ld16=nh * 32, out16_dtype=L.SRK_BF16, attn_table=td, attn_heads=nh, attn_scale=d ** -0.5,
              attn_shift=shift)
        x_ref = dv(res.clone())
        ln_ref = torch.full((M, Cp), 5.0, dtype=torch.bfloat16, device=DEV)
        _gemm(L, A=ao, a_mode=L.A_ROWS, lda=nh * 32, nB=B, H=H, W=W, Wt=wprd, M=M, N=Cp, K=nh * 32, dtype=L.SRK_BF16,
              bias=bprd, res=x_ref, out32=x_ref, ld32=Cp, win_shift=shift, ln_g=gd, ln_b=bd, ln_C=Cc,
              out16=ln_ref, ld16=Cp, out16_dtype=L.SRK_BF16)
        x_got = dv(res.clone())
        ln_got = torch.full((M, Cp), 7.0, dtype=torch.bfloat16, device=DEV)
        a = L.AttnBlockArgs()
        a.A, a.lda, a.M, a.C, a.Cp, a.H, a.W, a.shift, a.num_heads = L.ptr(Ad), Cp, M, Cc, Cp, H, W, shift, nh
        a.Wqkv, a.Wproj, a.b_proj, a.rel_table, a.scale = L.ptr(whmd), L.ptr(wprd), L.ptr(bprd), L.ptr(td), d ** -0.5
        a.res, a.out32, a.ld32, a.out16, a.ld16, a.out16_dtype = L.ptr(x_got), L.ptr(x_got), Cp, L.ptr(ln_got), Cp, L.SRK_BF16
        a.ln_g, a.ln_b, a.ln_C = L.ptr(gd), L.ptr(bd), Cc
        L.check(L.load().srk_attn_block(C.byref(a), L.stream_ptr()))
        torch.cuda.synchronize()
        scale = max(1.0, float(x_ref.abs().max()))
        assert float((x_got - x_ref).abs().max()) <= 2e-6 * scale, (shift, float((x_got - x_ref).abs().max()))
        assert float(x_got[:, Cc:].abs().max()) == 0.0
        dl = (ln_got.float() - ln_ref.float()).abs()
        assert float(dl.max()) <= 0.04 and float((dl > 0).float().mean()) < 0.02, (shift, float(dl.max()))
        assert float(ln_got.float()[:, Cc:].abs().max()) == 0.0


@pytest.mark.parametrize("dims", [(180, 192, 360, 384), (60, 64, 120, 128), (128, 128, 256, 256)])
def test_fused_mlp_kernel(L, dims):
    """srk_mlp (tcgen05): x + fc2(GELU(fc1(A)+b1))+b2 with fused LayerNorm / cast, vs fp32 torch on
    the same bf16-rounded operands."""
    if "tcgen05" not in ENGINES:
        pytest.skip("tcgen05 engine not under test")
    L.set_engine("tcgen05")
    Cc, Cp, hid, hid_p = dims
    g = torch.Generator().manual_seed(8)
    B, H, W = 3, 16, 24
    M = B * H * W
    A = torch.zeros(M, Cp); A[:, :Cc] = torch.randn(M, Cc, generator=g)
    W1 = torch.zeros(hid_p, Cp); W1[:hid, :Cc] = torch.randn(hid, Cc, generator=g) * 0.08
    W2 = torch.zeros(Cp, hid_p); W2[:Cc, :hid] = torch.randn(Cc, hid, generator=g) * 0.08
    b1 = torch.zeros(hid_p); b1[:hid] = torch.randn(hid, generator=g) * 0.1
    b2 = torch.zeros(Cp); b2[:Cc] = torch.randn(Cc, generator=g) * 0.1
    res = torch.zeros(M, Cp); res[:, :Cc] = torch.randn(M, Cc, generator=g)
    gam, bet = torch.randn(Cc, generator=g), torch.randn(Cc, generator=g)
    Ab, W1b, W2b = A.bfloat16(), W1.bfloat16(), W2.bfloat16()
    hidv = F.gelu(Ab.float() @ W1b.float().t() + b1).bfloat16().float()
    ref = res + hidv @ W2b.float().t() + b2
    refln = F.layer_norm(ref[:, :Cc], (Cc,), gam, bet, 1e-5)
    dv = lambda t: t.to(DEV)
    Ad, W1d, W2d, b1d, b2d, gd, bd = dv(Ab), dv(W1b), dv(W2b), dv(b1), dv(b2), dv(gam), dv(bet)
    for shift in (-1, 0, 4, None):
        x32 = dv(res.clone())
        o16 = torch.full((M, Cp), 5.0, dtype=torch.bfloat16 if shift is not None else torch.float16, device=DEV)
        m = L.MlpArgs()
        m.A, m.lda, m.M, m.C, m.Cp, m.hid_p = L.ptr(Ad), Cp, M, Cc, Cp, hid_p
        m.W1, m.b1, m.W2, m.b2 = L.ptr(W1d), L.ptr(b1d), L.ptr(W2d), L.ptr(b2d)
        m.res, m.out32, m.ld32, m.out16, m.ld16 = L.ptr(x32), L.ptr(x32), Cp, L.ptr(o16), Cp
        m.H, m.W, m.ln_win_shift = H, W, -1
        if shift is None:
            m.out16_dtype = L.SRK_FP16
        else:
            m.out16_dtype, m.ln_g, m.ln_b, m.ln_C, m.ln_win_shift = L.SRK_BF16, L.ptr(gd), L.ptr(bd), Cc, shift
        L.check(L.load().srk_mlp(C.byref(m), L.stream_ptr()))
        got32 = x32.cpu()
        assert float((got32[:, :Cc] - ref[:, :Cc]).abs().max()) < 4e-3 * max(1.0, float(ref.abs().max()))
        if Cp > Cc:
            assert float(got32[:, Cc:].abs().max()) == 0.0
        if shift is None:
            assert float((o16.float().cpu()[:, :Cc] - ref[:, :Cc]).abs().max()) < 0.02 * max(1.0, float(ref.abs().max()))
        else:
            exp = refln
            if shift >= 0:
                gm = O.window_gather_map(H, W, 8, shift)
                idx = (torch.arange(B)[:, None] * H * W + gm[None, :]).reshape(-1)
                exp = refln[idx]
            assert float((o16.float().cpu()[:, :Cc] - exp).abs().max()) < 0.05
            if Cp > Cc:
                assert float(o16.float().cpu()[:, Cc:].abs().max()) == 0.0


@pytest.mark.parametrize("geom", [(60, 6, 16, 16, 24), (180, 6, 32, 24, 16), (64, 2, 32, 8, 8)])
def test_window_attention(L, geom):
    Cdim, nh, dp, H, W = geom
    d = Cdim // nh
    B = 2
    g = torch.Generator().manual_seed(6)
    M = B * H * W
    nq = 3 * nh * dp
    ldq, ldo = (nq + 63) // 64 * 64, (nh * dp + 63) // 64 * 64
    table = torch.randn(225, nh, generator=g) * 0.5
    for shift in (0, 4):
        qkv = torch.zeros(M, 3, nh, dp)
        qkv[..., :d] = torch.randn(M, 3, nh, d, generator=g)
        qkv = qkv.bfloat16()
        buf = torch.zeros(M, ldq, dtype=torch.bfloat16)
        buf[:, :nq] = qkv.view(M, nq)
        out = torch.full((M, ldo), 9.0, dtype=torch.bfloat16, device=DEV)
        bufd, tabd = buf.to(DEV), table.t().contiguous().to(DEV)
        L.check(L.load().srk_window_attention(L.ptr(bufd), ldq, L.ptr(out), ldo, L.ptr(tabd), B, H, W, nh,
                                              dp, d ** -0.5, shift, L.stream_ptr()))
        q, k, v = [qkv[:, i, :, :d].float().view(-1, 64, nh, d).transpose(1, 2) for i in range(3)]
        att = (q @ k.transpose(-1, -2)) * d ** -0.5
        att = att + table[O.relative_position_index(8).reshape(-1)].view(64, 64, nh).permute(2, 0, 1)[None]
        if shift:
            m = O.shift_attention_mask(H, W, 8, shift)
            att = (att.view(B, -1, nh, 64, 64) + m[None, :, None]).view(-1, nh, 64, 64)
        p = torch.softmax(att, -1).bfloat16().float()
        o = (p @ v).transpose(1, 2).reshape(M, nh, d)
        got = out.float().cpu()[:, :nh * dp].view(M, nh, dp)
        assert float((got[..., :d] - o).abs().max()) < 0.03
        if dp > d:
            assert float(got[..., d:].abs().max()) == 0.0
        if ldo > nh * dp:
            assert float(out.float().cpu()[:, nh * dp:].abs().max()) == 0.0


def test_metrics_uint8_targets_equal_float_targets(L):
    """srk_metrics_h8 (targets shipped as the stored uint8 levels) == srk_metrics on H8 / 255."""
    from sr_caco_2_b200 import utils_image as UI
    g = torch.Generator().manual_seed(21)
    H8 = torch.randint(0, 256, (3, 1, 96, 80), generator=g, dtype=torch.uint8)
    H8[0, 0, :40] = torch.randint(0, 12, (40, 80), generator=g, dtype=torch.uint8)      # exercise the ROI thresholds
    E = (H8.float() / 255.0 + 0.05 * torch.randn(3, 1, 96, 80, generator=g)).clamp(0, 1)
    ths = (4, 5, 6, 7, 8, 9, 10)
    a = UI.compute_metrics(E.to(DEV), (H8.float() / 255.0).to(DEV), 4, ths)
    b = UI.compute_metrics(E.to(DEV), H8.to(DEV), 4, ths)
    assert torch.equal(a["raw"], b["raw"])
    # the same through the evaluator: uint8 low-res input and uint8 target, converted on the device
    from sr_caco_2_b200 import evaluator as EV
    L8 = torch.randint(0, 256, (3, 1, 24, 20), generator=g, dtype=torch.uint8)
    step = EV.make_bicubic_step(4, roi_ths=ths)
    v8 = step(L8, H8)
    vf = step(L8.float() / 255.0, H8.float() / 255.0)
    assert v8.shape == (3, 10) and torch.equal(v8, vf)


@pytest.mark.parametrize("scale", [2, 4, 8])
def test_bicubic_baseline_against_reference_golden(L, scale):
    """srk_bicubic_upsample vs the golden of Interpolate.forward (F.interpolate bicubic antialias + clamp)
    and vs the oracle on a larger random batch; then its metrics through the shared kernel."""
    import sr_caco_2_b200 as S
    z = np.load(os.path.join(T.GOLDEN, "bicubic_baseline.npz"))
    y = S.bicubic_upsample(torch.from_numpy(z[f"x{scale}"]).to(DEV), scale).cpu().numpy()
    assert y.shape == z[f"y{scale}"].shape
    assert float(np.abs(y - z[f"y{scale}"]).max()) < 2e-6
    x = T.synthetic_lr(3, 37, 29, 5)
    y = S.bicubic_upsample(x.to(DEV), scale)
    ref = torch.from_numpy(O.interpolate_baseline(x.numpy(), scale))
    assert float((y.cpu() - ref).abs().max()) < 2e-6
    m = S.Interpolate("super-resolution", scale, "bicubic")
    hr = ((ref + 0.04 * torch.randn(ref.shape, generator=torch.Generator().manual_seed(2))).clamp(0, 1) * 255).round() / 255
    m.feed_data({"l_im": x, "h_im": hr})
    m.test()
    vis = m.current_visuals()
    assert torch.equal(vis["E"], y)
    from sr_caco_2_b200 import utils_image as UI
    got = UI.compute_metrics(vis["E"], vis["H"], scale, (4, 6, 8))
    exp = O.all_metrics(ref, hr, scale)
    assert abs(float(got["psnr"][0]) - float(exp["psnr"][0])) < 0.01
    assert abs(float(got["ssim"][0]) - float(exp["ssim"][0])) < 1e-4


# ------------------------------------------------------------------------------------------
# networks
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("name", ["swinir_tiny_direct", "swinir_tiny_ps", "swinir_tiny_3conv", "swinir_tiny_nearest"])
def test_swinir_tiny_against_reference_golden(L, engine, name):
    L.set_engine(engine)
    z, sd, cfgd = load_npz(name)
    cfg = O.SwinIRCfg(**cfgd)
    net = make_swinir(cfg, sd)
    i = 0
    while f"x{i}" in z.files:
        x = torch.from_numpy(z[f"x{i}"])
        y = net(x.to(DEV)).cpu()
        ref = torch.from_numpy(z[f"y{i}"])
        emu = O.swinir_forward(sd, cfg, x, emulate_bf16=True)
        assert y.shape == ref.shape
        assert float((y - emu).abs().max()) < 1e-3, "indexing / arithmetic-contract mismatch"
        assert float((y - ref).abs().max()) < 2e-3, "SR output tolerance (north_star)"
        i += 1


@pytest.mark.parametrize("engine", ENGINES)
def test_edsr_tiny_against_reference_golden(L, engine):
    L.set_engine(engine)
    from sr_caco_2_b200 import EDSR
    z, sd, cfgd = load_npz("edsr_tiny")
    net = EDSR(in_chans=cfgd["in_chans"], n_resblocks=cfgd["n_resblocks"], n_feats=cfgd["n_feats"],
               scale=cfgd["scale"])
    net.load_state_dict(sd, strict=True)
    net = net.to(DEV).eval()
    y = net(torch.from_numpy(z["x0"]).to(DEV)).cpu()
    assert float((y - torch.from_numpy(z["y0"])).abs().max()) < 2e-3


@pytest.mark.parametrize("engine", ENGINES)
def test_baseline_configs_against_reference_samples(L, engine):
    """BASELINE.json configs on tests/common.py weights: strided samples of the unmodified
    reference's output (tests/golden/fullsize_samples.json), 2e-3 max-abs."""
    L.set_engine(engine)
    gold = json.load(open(os.path.join(T.GOLDEN, "fullsize_samples.json")))
    jobs = [("cfg1_light_x2_64", T.cfg_light_x2(), (4, 64, 64), 101),
            ("cfg1_light_x2_72", T.cfg_light_x2(), (2, 72, 72), 101),
            ("cfg3_classical_x8_64", T.cfg_classical(8), (1, 64, 64), 103),
            ("cfg4_classical_x4_72", T.cfg_classical(4), (1, 72, 72), 104)]
    for name, cfg, (B, h, w), seed in jobs:
        net = make_swinir(cfg, T.swinir_state_dict(cfg, seed=seed))
        y = net(T.synthetic_lr(B, h, w, seed).to(DEV)).cpu()
        gd = gold[name]
        assert list(y.shape) == gd["shape"]
        got = y.reshape(-1)[torch.tensor(gd["idx"])].double().numpy()
        err = np.abs(got - np.array(gd["val"])).max()
        assert err < 2e-3, (name, err)
        del net
    from sr_caco_2_b200 import EDSR
    cfg = T.cfg_edsr_x4()
    net = EDSR(in_chans=1, n_resblocks=16, n_feats=64, scale=4)
    net.load_state_dict(T.edsr_state_dict(cfg, seed=102), strict=True)
    y = net.to(DEV).eval()(T.synthetic_lr(2, 64, 64, 102).to(DEV)).cpu()
    gd = gold["cfg2_edsr_x4_64"]
    got = y.reshape(-1)[torch.tensor(gd["idx"])].double().numpy()
    assert np.abs(got - np.array(gd["val"])).max() < 2e-3


def test_cfg1_full_output_and_metrics_against_oracle(L):
    """configs[0] end to end: whole SR output vs the CPU oracle (2e-3) and the scores the
    evaluator reports vs the oracle's scores of the oracle's output (0.01 dB / 1e-4)."""
    from sr_caco_2_b200 import utils_image as UI
    cfg = T.cfg_light_x2()
    sd = T.swinir_state_dict(cfg, seed=101)
    net = make_swinir(cfg, sd)
    x = T.synthetic_lr(4, 64, 64, 101)
    ref = O.swinir_forward(sd, cfg, x)
    y = net(x.to(DEV))
    assert float((y.cpu() - ref).abs().max()) < 2e-3
    Hr = (ref + 0.03 * torch.randn(ref.shape, generator=torch.Generator().manual_seed(1))).clamp(0, 1)
    Hr = (Hr * 255).round() / 255
    m = UI.compute_metrics(y, Hr.to(DEV), cfg.upscale, (4, 5, 6, 7, 8, 9, 10))
    om, orm = O.all_metrics(ref, Hr, cfg.upscale), O.roi_marginal_metrics(ref, Hr, cfg.upscale)
    assert float((m["psnr"].cpu() - om["psnr"]).abs().max()) < 0.01
    assert float((m["ssim"].cpu() - om["ssim"].double()).abs().max()) < 1e-4
    assert float((m["psnr_y"].cpu() - om["psnr_y"]).abs().max()) < 0.01
    assert float((m["roi_psnr"].cpu() - orm["psnr"]).abs().max()) < 0.01
    assert float((m["roi_ssim"].cpu() - orm["ssim"].double()).abs().max()) < 1e-4
    assert float(((m["nrmse"].cpu() - om["nrmse"]) / om["nrmse"]).abs().max()) < 5e-3


def test_headline_config_properties_at_full_size(L):
    """configs[2] (classical X8, B=32, 64x64 -> 512x512) at BASELINE size: batch independence,
    determinism, reflect-pad equivalence and finiteness -- properties that need no CPU oracle."""
    cfg = T.cfg_classical(8)
    net = make_swinir(cfg, T.swinir_state_dict(cfg, seed=103))
    x = T.synthetic_lr(32, 64, 64, 103).to(DEV)
    y = net(x)
    assert y.shape == (32, 1, 512, 512) and bool(torch.isfinite(y).all())
    assert torch.equal(y, net(x))                                  # deterministic
    y1 = net(x[5:6])
    assert float((y[5:6] - y1).abs().max()) == 0.0                  # patches are independent
    perm = torch.randperm(32, device=DEV)
    assert torch.equal(net(x[perm]), y[perm])
    # a 61x59 crop is reflect-padded inside the net exactly like F.pad(..., 'reflect')
    xc = x[:2, :, :61, :59].contiguous()
    yc = net(xc)
    yp = net(F.pad(xc, (0, 5, 0, 3), mode="reflect"))[:, :, :61 * 8, :59 * 8]
    assert yc.shape == (2, 1, 488, 472) and float((yc - yp).abs().max()) == 0.0


def test_evaluator_end_to_end_matches_oracle(L):
    from sr_caco_2_b200.evaluator import evaluate_patches, make_cuda_step, pad_for_windows
    cfg = O.SwinIRCfg(upscale=2, in_chans=1, img_size=32, depths=[2, 2], embed_dim=60, num_heads=[6, 6],
                      mlp_ratio=2, upsampler="pixelshuffledirect")
    sd = T.swinir_state_dict(cfg, seed=9)
    net = make_swinir(cfg, sd)
    lr = T.synthetic_lr(5, 24, 24, 9)
    ref = O.swinir_forward(sd, cfg, pad_for_windows(lr))[..., :48, :48]
    hr = ((ref + 0.02 * torch.randn(ref.shape, generator=torch.Generator().manual_seed(2))).clamp(0, 1) * 255).round() / 255
    res = evaluate_patches(make_cuda_step(net, 2, True), lr, hr, batch_size=2, device=torch.device(DEV))
    om, orm = O.all_metrics(ref, hr, 2), O.roi_marginal_metrics(ref, hr, 2)
    assert res["n"] == 5
    assert abs(res["psnr"] - float(om["psnr"].mean())) < 0.01
    assert abs(res["ssim"] - float(om["ssim"].double().mean())) < 1e-4
    assert abs(res["roi_psnr"] - float(orm["psnr"].mean())) < 0.01
    assert abs(res["roi_ssim"] - float(orm["ssim"].double().mean())) < 1e-4


# ------------------------------------------------------------------------------------------
# round 2: the geometries / batch sizes the headline numbers are quoted on
# ------------------------------------------------------------------------------------------
def _r2_gold():
    return json.load(open(os.path.join(T.GOLDEN, "fullsize_r2.json")))


@pytest.mark.parametrize("name", ["cfg5_classical_x2_128", "cfg5_classical_x2_136", "cfg4_classical_x4_64"])
def test_baseline_cfg4_cfg5_against_reference_samples(L, name):
    """BASELINE configs[3] at 64x64 and configs[4] at 128x128 and at 136x136 (eval.py's padded geometry: 17 x 17
    windows, odd window counts): 2048 strided samples of the unmodified reference's output
    (tests/golden/make_golden_r2.py), 2e-3 max-abs.  Product path: fused attention block + fused MLP kernels."""
    L.set_engine("tcgen05")
    cfg, shape, seed = {"cfg5_classical_x2_128": (T.cfg_classical(2), (1, 128, 128), 105),
                        "cfg5_classical_x2_136": (T.cfg_classical(2), (1, 136, 136), 105),
                        "cfg4_classical_x4_64": (T.cfg_classical(4), (1, 64, 64), 104)}[name]
    net = make_swinir(cfg, T.swinir_state_dict(cfg, seed=seed))
    y = net(T.synthetic_lr(*shape, seed).to(DEV)).cpu()
    gd = _r2_gold()[name]
    assert list(y.shape) == gd["shape"]
    got = y.reshape(-1)[torch.tensor(gd["idx"])].double().numpy()
    assert np.abs(got - np.array(gd["val"])).max() < 2e-3
    assert abs(float(y.double().sum()) - gd["sum"]) < 2e-3 * y.numel() * 0.05


def test_img_size_8_disables_the_shift(L):
    """img_size == window_size: the reference constructor sets shift_size = 0 for every block
    (network_swinir.py:232-236); golden from the unmodified reference on a 16 x 24 input."""
    for engine in ENGINES:
        L.set_engine(engine)
        cfg = T.cfg_imgsize8()
        net = make_swinir(cfg, T.swinir_state_dict(cfg, seed=106))
        assert all(s == 0 for _, s in net._block_geometry)
        y = net(T.synthetic_lr(2, 16, 24, 106).to(DEV)).cpu()
        gd = _r2_gold()["swinir_imgsize8"]
        assert list(y.shape) == gd["shape"]
        got = y.reshape(-1)[torch.tensor(gd["idx"])].double().numpy()
        assert np.abs(got - np.array(gd["val"])).max() < 2e-3


def test_large_magnitude_activations_stress(L):
    """conv_first scaled by 2048: every fp16 conv operand (residual stream, conv_after_body's sum with the shallow
    features, the upsampler features) is three orders of magnitude larger than with random-init-like weights, the
    range a trained checkpoint could reach (none is available offline), still below the fp16 maximum the conv
    operands saturate at.  Tolerance relative to the output magnitude (|y| up to ~200)."""
    L.set_engine("tcgen05")
    cfg = T.cfg_stress()
    sd = T.stress_state_dict(T.swinir_state_dict(cfg, seed=107))
    net = make_swinir(cfg, sd)
    y = net(T.synthetic_lr(2, 24, 32, 107).to(DEV)).cpu()
    gd = _r2_gold()["swinir_stress_large_weights"]
    assert list(y.shape) == gd["shape"] and bool(torch.isfinite(y).all())
    got = y.reshape(-1)[torch.tensor(gd["idx"])].double().numpy()
    assert np.abs(got - np.array(gd["val"])).max() < 2e-3 * gd["absmax"]


def test_cfg3_batch32_patches_and_scores_against_reference(L):
    """The headline configuration AT its batch size: cfg3 (classical X8, 64 -> 512) on the seeded batch of 32; the
    full 512 x 512 output of patches 0, 13 and 31 against the unmodified reference (2e-3), and PSNR / SSIM / NRMSE
    of those patches against the reference's own metric functions on the reference's output (0.01 dB / 1e-4)."""
    from sr_caco_2_b200 import utils_image as UI
    L.set_engine("tcgen05")
    z = np.load(os.path.join(T.GOLDEN, "cfg3_b32_patches.npz"))
    cfg = T.cfg_classical(8)
    net = make_swinir(cfg, T.swinir_state_dict(cfg, seed=103))
    x = T.synthetic_lr(32, 64, 64, 203)
    hr = T.synthetic_hr(32, 512, 512, 204)
    y = net(x.to(DEV))
    m = UI.compute_metrics(y, hr.to(DEV), 8, (4, 7, 10))
    for i in [int(v) for v in z["patches"]]:
        ref = torch.from_numpy(z[f"y{i}"])
        assert float((y[i, 0].cpu() - ref).abs().max()) < 2e-3, i
        assert abs(float(m["psnr"][i]) - float(z[f"psnr{i}"][0])) < 0.01, i
        assert abs(float(m["ssim"][i]) - float(z[f"ssim{i}"][0])) < 1e-4, i
        assert abs(float(m["nrmse"][i]) - float(z[f"nrmse{i}"][0])) < 1e-4 * max(1.0, float(z[f"nrmse{i}"][0])), i
    # the batch is processed as independent patches: a patch alone gives the same pixels
    y13 = net(x[13:14].to(DEV))
    assert float((y13[0] - y[13]).abs().max()) < 1e-5


def test_fused_block_kernels_equal_unfused_network(L):
    """Whole network with the fused attention block / fused MLP kernels switched off one at a time (plan options):
    same output to the fp32 rounding of the reordered residual adds and a bf16 ulp of the LayerNorm rows."""
    L.set_engine("tcgen05")
    cfg = O.SwinIRCfg(upscale=4, img_size=16, embed_dim=180, depths=[3, 2], num_heads=[6, 6], mlp_ratio=2.0,
                      upsampler="pixelshuffle")
    net = make_swinir(cfg, T.swinir_state_dict(cfg, 9))
    x = T.synthetic_lr(3, 24, 40, 4).to(DEV)
    y = net(x).clone()
    for opt in (L.OPT_NO_FUSED_BLOCK, L.OPT_NO_FUSED_MLP, L.OPT_NO_FUSED_BLOCK | L.OPT_NO_FUSED_MLP,
                L.OPT_NO_FUSED_ATTN, L.OPT_NO_FOLD_QKV_BIAS):
        net.options = opt
        y2 = net(x)
        assert float((y2 - y).abs().max()) < 1e-3, opt
    net.options = 0


def test_nan_and_inf_in_the_sr_output_are_reported(L):
    """the clamp of the uint8 quantisation hides NaN / Inf pixels; the reference propagates them into every
    metric and aborts (check_negative_non_float, utils_trainer.py:933-958).  flags bit 0 must be raised by a
    non-finite INPUT pixel, compute_metrics(check=True) and the evaluator sweep must raise."""
    from sr_caco_2_b200 import utils_image as UI
    from sr_caco_2_b200 import evaluator as EV
    E, Hr = T.synthetic_pair(3, 64, 72, 5)
    for bad in (float("nan"), float("inf"), float("-inf")):
        Eb = E.clone(); Eb[1, 0, 20, 31] = bad
        m = UI.compute_metrics(Eb.to(DEV), Hr.to(DEV), 4, (4, 7), check=False)
        f = m["flags"].cpu()
        assert int(f[1]) & 1 and not (int(f[0]) & 1) and not (int(f[2]) & 1), (bad, f)
        with pytest.raises(FloatingPointError):
            UI.compute_metrics(Eb.to(DEV), (Hr * 255).round().to(torch.uint8).to(DEV), 4, (4, 7), check=True)
        out, fl = UI._run(Eb.to(DEV) * 255, Hr.to(DEV) * 255, 4, None, False)      # generic (non-quantising) path
        assert int(fl[1].cpu()) & 1
    clean = UI.compute_metrics(E.to(DEV), Hr.to(DEV), 4, (4, 7), check=True)
    assert int(clean["flags"].max().cpu()) == 0

    class Net(torch.nn.Module):
        window_size = 8
        def __init__(self):
            super().__init__(); self.p = torch.nn.Parameter(torch.zeros(1, device=DEV))
        def forward(self, x):
            y = torch.nn.functional.interpolate(x, scale_factor=2, mode="nearest")
            y[0, 0, 3, 3] = float("nan")
            return y
    step = EV.make_cuda_step(Net(), 2, swinir_padding=False)
    lr = T.synthetic_lr(4, 32, 32, 1); hr = T.synthetic_hr(4, 64, 64, 2)
    with pytest.raises(FloatingPointError):
        EV.evaluate_patches(step, lr, hr, 2, device=torch.device(DEV))


def test_metric_shims_keep_uint8_levels_unscaled(L):
    """the mbatch_gpu_calculate_* shims take inputs in [0,255]; a uint8 target is levels, not [0,1]."""
    from sr_caco_2_b200 import utils_image as UI
    E, Hr = T.synthetic_pair(2, 48, 48, 9)
    e255 = (E.clamp(0, 1) * 255).round()
    h255 = (Hr * 255).round()
    a = UI.mbatch_gpu_calculate_psnr(e255.to(DEV), h255.to(DEV), border=2)
    b = UI.mbatch_gpu_calculate_psnr(e255.to(DEV), h255.to(torch.uint8).to(DEV), border=2)
    assert torch.allclose(a, b, rtol=0, atol=1e-9)


def test_plan_follows_in_place_parameter_updates(L):
    """the packed-weight plan is rebuilt after an in-place parameter change (EMA update, optimizer step)."""
    L.set_engine("tcgen05")
    cfg = O.SwinIRCfg(upscale=2, img_size=16, embed_dim=60, depths=[2], num_heads=[6], mlp_ratio=2.0,
                      upsampler="pixelshuffledirect")
    sd = T.swinir_state_dict(cfg, 3)
    net = make_swinir(cfg, sd)
    x = T.synthetic_lr(1, 16, 16, 2).to(DEV)
    y0 = net(x).clone()
    with torch.no_grad():
        net.conv_first.weight.mul_(1.5)
    y1 = net(x)
    sd2 = dict(sd); sd2["conv_first.weight"] = sd["conv_first.weight"] * 1.5
    ref = O.swinir_forward(sd2, cfg, x.cpu())
    assert float((y1.cpu() - ref).abs().max()) < 2e-3 and float((y1 - y0).abs().max()) > 1e-3


def test_cuda_graph_replay_is_bit_identical_to_eager_launches(L):
    """The forward of a shape is launched eagerly once, captured on the second call and replayed afterwards
    (network_swinir._forward_graph); options = SRK_OPT_NO_GRAPH keeps the plain launch sequence."""
    L.set_engine("tcgen05")
    cfg = O.SwinIRCfg(upscale=2, img_size=16, embed_dim=180, depths=[2, 2], num_heads=[6, 6], mlp_ratio=2.0,
                      upsampler="pixelshuffle")
    net = make_swinir(cfg, T.swinir_state_dict(cfg, 6))
    xs = [T.synthetic_lr(2, 24, 32, s).to(DEV) for s in (1, 2, 3, 4)]
    net.options = L.OPT_NO_GRAPH
    ref = [net(x).clone() for x in xs]
    net.options = 0
    L.launch_count(reset=True)
    got = [net(x) for x in xs]                        # eager, capture + replay, replay, replay
    n = L.launch_count()
    for a, b in zip(got, ref):
        assert torch.equal(a, b)
    key = next(iter(net._graphs))
    assert net._graphs[key]["graph"] is not None and n == 4 * net._graphs[key]["launches"]
    y2 = net(T.synthetic_lr(1, 16, 16, 9).to(DEV))    # another shape gets its own graph
    assert y2.shape == (1, 1, 32, 32) and len(net._graphs) == 2
