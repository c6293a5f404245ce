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
