"""Experiment: per-role timeline of CTA 0 of the fused attention block kernel (built with -DSRK_AB_TRACE into a side library)."""
import ctypes as C, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
PKG = os.path.join(ROOT, "sr_caco_2_b200")
SO = os.path.join(ROOT, "scripts", "micro", "libsrk_trace.so")
if "--build" in sys.argv:
    from sr_caco_2_b200 import build as B
    srcs = [os.path.join(PKG, "csrc", f) for f in B.SOURCES]
    subprocess.check_call([B.nvcc_path(), *[f for f in B.NVCC_FLAGS if f not in ("-Xptxas", "-v")], "-DSRK_AB_TRACE", "-DSRK_MLP_TRACE", "-shared", "-o", SO, *srcs, "-lcudart"])
    sys.exit(0)
import torch
from sr_caco_2_b200 import _lib as L
L.LIB_PATH = SO
lib = L.load()
dev = "cuda:0"
B_, H, W, Cp = 32, 64, 64, 192
M = B_ * H * W
t16 = lambda *s: (torch.randn(*s, device=dev) * 0.1).bfloat16()
A, Wp = t16(M, Cp), t16(Cp, Cp)
bias = torch.zeros(576, device=dev); X = torch.randn(M, Cp, device=dev); X2 = torch.empty_like(X)
A16 = torch.empty(M, Cp, device=dev, dtype=torch.bfloat16); lng = torch.ones(192, device=dev); lnb = torch.zeros(192, device=dev)
whm = t16(576, Cp); tab = torch.randn(6, 225, device=dev) * 0.3
a = L.AttnBlockArgs()
a.A, a.lda, a.M, a.C, a.Cp, a.H, a.W, a.shift, a.num_heads = L.ptr(A), Cp, M, 180, Cp, H, W, 4, 6
a.Wqkv, a.Wproj, a.b_proj, a.rel_table, a.scale = L.ptr(whm), L.ptr(Wp), L.ptr(bias), L.ptr(tab), 30 ** -0.5
a.res, a.out32, a.ld32, a.out16, a.ld16, a.out16_dtype = L.ptr(X), L.ptr(X2), Cp, L.ptr(A16), Cp, 0
a.ln_g, a.ln_b, a.ln_C = L.ptr(lng), L.ptr(lnb), 180
for _ in range(3): L.check(lib.srk_attn_block(C.byref(a), L.stream_ptr()))
torch.cuda.synchronize()
buf = (C.c_longlong * (4 * 64 * 8))()
lib.srk_debug_ab_trace.restype = C.c_int
assert lib.srk_debug_ab_trace(buf) == 0
v = list(buf)
t0 = min(x for x in v if x > 0)
names = {0: "group 0, unit 3*it + h/2  [wait q_full, q ready, drained, bar1 passed, ao_empty ok, attention done, bar2 passed]",
         1: "group 1", 2: "group 0, final quarter 2*tp + qi  [residual requested, staged (T done), done]", 3: "group 1, final quarters"}
for role in (0, 1, 2, 3):
    print(names[role])
    for i in range(30):
        e = v[(role * 64 + i) * 8:(role * 64 + i) * 8 + 8]
        if any(e): print(f"  {i:3d} " + " ".join(f"{(x - t0) if x else -1:8d}" for x in e))
