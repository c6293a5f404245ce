"""Weight repacking from the reference state_dict layout into the
padded, K-contiguous 16-bit operands of the GEMM kernels (see DESIGN.md "Data layout")."""
from __future__ import annotations

import torch

from . import _lib as L


def up(v: int, m: int) -> int:
    return (v + m - 1) // m * m


def _dt(code):
    return torch.bfloat16 if code == L.SRK_BF16 else torch.float16


def pad_bias(b: torch.Tensor, n_p: int) -> torch.Tensor:
    out = torch.zeros(n_p, dtype=torch.float32, device=b.device)
    out[: b.numel()] = b.float()
    return out.contiguous()


def pack_linear(w: torch.Tensor, n_p: int, k_p: int, dtype_code: int) -> torch.Tensor:
    """nn.Linear weight (N, K) -> (n_p, k_p) zero padded, K contiguous."""
    n, k = w.shape
    out = torch.zeros(n_p, k_p, dtype=torch.float32, device=w.device)
    out[:n, :k] = w.float()
    return out.to(_dt(dtype_code)).contiguous()


def pack_qkv(w: torch.Tensor, b: torch.Tensor, nh: int, d: int, dp: int, n_p: int, k_p: int,
             dtype_code: int):
    """qkv Linear (3C, C) whose rows are [q | k | v], each head-major (reshape(B_,N,3,nH,d),
    network_swinir.py:148-149) -> rows [which][head][dp] with head_dim zero-padded to dp."""
    C = w.shape[1]
    w3 = w.float().view(3, nh, d, C)
    wp = torch.zeros(3, nh, dp, k_p, dtype=torch.float32, device=w.device)
    wp[:, :, :d, :C] = w3
    out = torch.zeros(n_p, k_p, dtype=torch.float32, device=w.device)
    out[: 3 * nh * dp] = wp.view(3 * nh * dp, k_p)
    bp = torch.zeros(3, nh, dp, dtype=torch.float32, device=w.device)
    bp[:, :, :d] = b.float().view(3, nh, d)
    bias = torch.zeros(n_p, dtype=torch.float32, device=w.device)
    bias[: 3 * nh * dp] = bp.view(-1)
    return out.to(_dt(dtype_code)).contiguous(), bias.contiguous()


def fold_qkv_bias(w_packed: torch.Tensor, bias_packed: torch.Tensor, c: int) -> torch.Tensor:
    """Packed qkv weight (n_p, k_p) with the bias folded into the pad K columns c (bf16 of the bias) and
    c + 1 (bf16 of the remainder): with A[:, c] = A[:, c + 1] = 1 the GEMM adds bias to 2^-17 relative."""
    assert w_packed.shape[1] - c >= 2
    out = w_packed.clone()
    hi = bias_packed.to(w_packed.dtype)
    lo = (bias_packed - hi.float()).to(w_packed.dtype)
    out[:, c], out[:, c + 1] = hi, lo
    return out.contiguous()


def pack_qkv_heads(w_packed: torch.Tensor, nh: int, dp: int) -> torch.Tensor:
    """Packed qkv weight rows [which][head][dp] -> head-major rows [head][which][dp] (srk_attn_block streams one
    head's 3 * dp rows [q | k | v] per MMA job)."""
    k_p = w_packed.shape[1]
    return w_packed[: 3 * nh * dp].view(3, nh, dp, k_p).permute(1, 0, 2, 3).reshape(3 * nh * dp, k_p).contiguous()


def pack_proj(w: torch.Tensor, nh: int, d: int, dp: int, n_p: int, k_p: int, dtype_code: int):
    """proj Linear (C, C): input column (head, e) = head*d + e -> column head*dp + e."""
    C = w.shape[0]
    w3 = w.float().view(C, nh, d)
    wp = torch.zeros(C, nh, dp, dtype=torch.float32, device=w.device)
    wp[:, :, :d] = w3
    out = torch.zeros(n_p, k_p, dtype=torch.float32, device=w.device)
    out[:C, : nh * dp] = wp.view(C, nh * dp)
    return out.to(_dt(dtype_code)).contiguous()


def pack_conv3x3(w: torch.Tensor, b: torch.Tensor, cin_p: int, n_p: int, dtype_code: int,
                 pixel_shuffle_r: int = 0):
    """nn.Conv2d weight (Cout, Cin, 3, 3) -> (n_p, 9*cin_p) with k = (ky*3+kx)*cin_p + c.
    pixel_shuffle_r = 2: output rows are re-ordered from the reference channel c*4 + i*2 + j to
    (i*2+j)*(Cout/4) + c so that the GEMM epilogue writes NHWC pixels of the shuffled image
    directly (nn.PixelShuffle, network_swinir.py:675)."""
    cout, cin = w.shape[:2]
    wk = torch.zeros(cout, 3, 3, cin_p, dtype=torch.float32, device=w.device)
    wk[..., :cin] = w.float().permute(0, 2, 3, 1)
    wk = wk.reshape(cout, 9 * cin_p)
    bb = b.float()
    if pixel_shuffle_r:
        r2 = pixel_shuffle_r * pixel_shuffle_r
        cq = cout // r2
        perm = torch.arange(cout, device=w.device).view(cq, r2).t().reshape(-1)   # new row -> old
        wk, bb = wk[perm], bb[perm]
    out = torch.zeros(n_p, 9 * cin_p, dtype=torch.float32, device=w.device)
    out[:cout] = wk
    return out.to(_dt(dtype_code)).contiguous(), pad_bias(bb, n_p)


def pack_conv_in(w: torch.Tensor, b: torch.Tensor):
    """(C, 1, 3, 3) -> (C, 9) fp32."""
    return w.float().reshape(w.shape[0], 9).contiguous(), b.float().contiguous()


def pack_conv_out(w: torch.Tensor):
    """(1, Cin, 3, 3) -> (9, Cin) fp16, tap major."""
    return w.float()[0].permute(1, 2, 0).reshape(9, w.shape[1]).to(torch.float16).contiguous()


def fold_tail(up_convs, last_w: torch.Tensor, last_b: torch.Tensor, scale: int):
    """Compose the linear reconstruction tail -- log2(s) x [Conv3x3(F->4F) + PixelShuffle(2)] then
    Conv3x3(F->1) (network_swinir.py:661-680, 868; network_nlsn.py Upsampler + tail conv) -- into one
    5x5 convolution F -> s*s per border variant (include/srk.h, srk_tail_fold).

    up_convs: [(weight (4F,F,3,3), bias (4F))] in the reference layout (PixelShuffle channel order).
    The kernels are measured, not derived: the tail is a linear map, so its response to unit
    impulses on a 5x5 feature grid (fp64, plain torch ops on the parameter device, run once at
    plan-build time) IS the composed kernel, including the effect of every intermediate zero
    padding for a target pixel in the first / interior / last row and column.
    Returns (w (64, 1600) fp16, b (64) fp32, border_w (9, 64, 1600) fp16 in the ring-pass fragment
    order, border_b (9, 64) fp32, w_scale)."""
    import torch.nn.functional as Fn
    dev = last_w.device
    F = last_w.shape[1]
    if F != 64:
        raise ValueError("fold_tail: built for 64 feature channels")
    s2 = scale * scale
    if s2 > 64:
        raise ValueError("fold_tail: scale*scale must be <= 64")
    ups = [(w.detach().double(), b.detach().double()) for w, b in up_convs]
    lw, lb = last_w.detach().double(), last_b.detach().double()

    def tail(x):
        for w, b in ups:
            x = Fn.pixel_shuffle(Fn.conv2d(x, w, b, padding=1), 2)
        return Fn.conv2d(x, lw, lb, padding=1)

    G = 5
    # impulse k = (gy*5 + gx)*F + c at grid position (gy, gx), channel c; the last image is all zero
    X = torch.zeros(G * G * F + 1, F, G, G, dtype=torch.float64, device=dev)
    idx = torch.arange(G * G * F, device=dev)
    X[idx, idx % F, (idx // F) // G, (idx // F) % G] = 1.0
    # (in chunks: at X8 the fp64 intermediates of all 1601 impulse images at once would take several GB)
    out = torch.cat([tail(X[i:i + 128])[:, 0] for i in range(0, X.shape[0], 128)], 0)   # (1601, 5s, 5s)
    W = torch.zeros(9, 64, 25 * F, dtype=torch.float64, device=dev)
    Bv = torch.zeros(9, 64, dtype=torch.float64, device=dev)
    for vy in range(3):
        for vx in range(3):
            ty, tx = 2 * vy, 2 * vx                               # target pixel: first / interior / last
            blk = out[:, ty * scale:(ty + 1) * scale, tx * scale:(tx + 1) * scale].reshape(-1, s2)
            bias = blk[-1]
            resp = (blk[:-1] - bias).view(G, G, F, s2)             # [gy][gx][c][n]
            v = vy * 3 + vx
            Bv[v, :s2] = bias
            for dy in range(-2, 3):
                for dx in range(-2, 3):
                    gy, gx = ty + dy, tx + dx
                    if 0 <= gy < G and 0 <= gx < G:
                        tap = (dy + 2) * 5 + (dx + 2)
                        W[v, :s2, tap * F:(tap + 1) * F] = resp[gy, gx].t()
    wmax = float(W.abs().max())
    w_scale = 1.0 if wmax == 0.0 else 2.0 ** (13 - int(torch.floor(torch.log2(torch.tensor(wmax))).item()))
    Wh = (W * w_scale).to(torch.float16)
    Bs = (Bv * w_scale).float()
    # ring-pass layout: inside every 64-channel block k = ks*16 + 2t + 8w + h is stored at [t][ks][w][h]
    # (the mma.sync B fragments of a tap become 32 contiguous bytes per lane, elementwise.cu)
    Wb = Wh.view(9, 64, 25, 4, 2, 4, 2).permute(0, 1, 2, 5, 3, 4, 6).reshape(9, 64, 25 * F).contiguous()
    return (Wh[4].contiguous(), Bs[4].contiguous(), Wb, Bs.contiguous(), float(w_scale))


def pack_conv_nearest2x(w: torch.Tensor, b: torch.Tensor, cin_p: int, dtype_code: int):
    """Conv3x3 applied to a nearest x2 up-sampled image (network_swinir.py:953-960) as ONE 3x3 conv on the
    LOW-res grid with 4*Cout outputs in PixelShuffle order (i, j, c): output pixel (2y+i, 2x+j) reads
    up-sampled pixel (2y+i+dy, 2x+j+dx) = low-res pixel (y + (i+dy)//2, x + (j+dx)//2), so the taps that
    land on the same low-res pixel are summed.  The zero padding of the up-sampled image coincides with
    the zero padding of the low-res one (row -1 <-> row -1, row 2H <-> row H).
    w: (Cout, Cin, 3, 3) -> ((4*Cout, 9*cin_p) 16-bit, (4*Cout) fp32 bias)."""
    cout, cin = w.shape[:2]
    wf = w.float()
    comp = torch.zeros(2, 2, cout, 3, 3, cin_p, dtype=torch.float32, device=w.device)   # [i][j][c][oy+1][ox+1][cin]
    for i in range(2):
        for j in range(2):
            for dy in (-1, 0, 1):
                for dx in (-1, 0, 1):
                    oy, ox = (i + dy) // 2, (j + dx) // 2
                    comp[i, j, :, oy + 1, ox + 1, :cin] += wf[:, :, dy + 1, dx + 1]
    out = comp.reshape(4 * cout, 9 * cin_p)
    bias = b.float().repeat(4)
    return out.to(_dt(dtype_code)).contiguous(), bias.contiguous()
