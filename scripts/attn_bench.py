"""Timing of the fused qkv + window-attention call at the cfg3 geometry (B=32, 64x64 tokens, 6 heads)."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sr_caco_2_b200 import _lib as L
L.load(); L.set_engine("tcgen05")
dev = "cuda:0"
B, H, W, nh = int(os.environ.get("SRK_B", 32)), 64, 64, 6
M = B * H * W
A = (torch.randn(M, 192, device=dev) * 0.5).bfloat16()
Wq = (torch.randn(576, 192, device=dev) * 0.08).bfloat16()
bias = torch.randn(576, device=dev) * 0.1
table = torch.randn(nh, 225, device=dev) * 0.5
out = torch.empty(M, 192, device=dev, dtype=torch.bfloat16)
g = L.GemmArgs(); g.res_scale, g.win_shift, g.ln_win_shift = 1.0, -1, -1
nobias = os.environ.get('SRK_NOBIAS') == '1'
kw = dict(A=A, a_mode=0, lda=192, nB=B, H=H, W=W, Wt=Wq, M=M, N=576, K=192, dtype=0, bias=(0 if nobias else bias), out16=out, ld16=192,
          out16_dtype=0, attn_table=table, attn_heads=nh, attn_scale=30 ** -0.5, attn_shift=4)
for k, v in kw.items():
    setattr(g, k, L.ptr(v) if isinstance(v, torch.Tensor) else v)
lib = L.load()
for _ in range(3): L.check(lib.srk_gemm(C.byref(g), L.stream_ptr()))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 20
e0.record()
for _ in range(n): L.check(lib.srk_gemm(C.byref(g), L.stream_ptr()))
e1.record(); torch.cuda.synchronize()
print(f"NOBIAS={int(nobias)} TC5={os.environ.get('SRK_ATTN_TC5','1')} DBG={os.environ.get('SRK_QA_DBG','0')}: {e0.elapsed_time(e1) / n * 1e3:.1f} us", flush=True)
