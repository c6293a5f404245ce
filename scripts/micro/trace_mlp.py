"""Experiment: per-role timeline of CTA 0 of the fused MLP kernel (built with -DSRK_MLP_TRACE into a side library)."""
import ctypes as C, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
PKG = os.path.join(ROOT, "sr_caco_2_b200")
SO = os.path.join(ROOT, "scripts", "micro", "libsrk_trace.so")
if "--build" in sys.argv:
    from sr_caco_2_b200 import build as B
    srcs = [os.path.join(PKG, "csrc", f) for f in B.SOURCES]
    subprocess.check_call([B.nvcc_path(), *[f for f in B.NVCC_FLAGS if f not in ("-Xptxas", "-v")], "-DSRK_MLP_TRACE", "-shared", "-o", SO, *srcs, "-lcudart"])
    sys.exit(0)
import torch
from sr_caco_2_b200 import _lib as L
L.LIB_PATH = SO
lib = L.load()
dev = "cuda:0"
B_, H, W, Cp = 32, 64, 64, 192
M = B_ * H * W
t16 = lambda *s: (torch.randn(*s, device=dev) * 0.1).bfloat16()
A, W1, W2 = t16(M, Cp), t16(384, Cp), t16(Cp, 384)
bias = torch.zeros(576, device=dev); X = torch.randn(M, Cp, device=dev); X2 = torch.empty_like(X)
A16 = torch.empty(M, Cp, device=dev, dtype=torch.bfloat16); lng = torch.ones(180, device=dev); lnb = torch.zeros(180, device=dev)
m = L.MlpArgs()
m.A, m.lda, m.M, m.C, m.Cp, m.hid_p = L.ptr(A), Cp, M, 180, Cp, 384
m.W1, m.b1, m.W2, m.b2 = L.ptr(W1), L.ptr(bias), L.ptr(W2), L.ptr(bias)
m.res, m.out32, m.ld32, m.out16, m.ld16 = L.ptr(X), L.ptr(X2), Cp, L.ptr(A16), Cp
m.H, m.W, m.out16_dtype, m.ln_g, m.ln_b, m.ln_C, m.ln_win_shift = H, W, 0, L.ptr(lng), L.ptr(lnb), 180, 4
for _ in range(3): L.check(lib.srk_mlp(C.byref(m), L.stream_ptr()))
torch.cuda.synchronize()
buf = (C.c_longlong * (4 * 64 * 8))()
lib.srk_debug_mlp_trace.restype = C.c_int
assert lib.srk_debug_mlp_trace(buf) == 0
v = list(buf)
t0 = min(x for x in v if x > 0)
names = {0: "MMA  q  [loop top, fc1(q+2) issued, before w_full, fc2 issue | fc1: d1_empty ok, w_full ok]", 1: "GELU q  [wait d1, d1 ready, loads done, h_full | h_empty ok, piece0 done, piece1 load ok, piece1 done]", 2: "FIN  t  [wait d2, d2 ready, half0 done, half1 done]"}
for role in (0, 1, 2):
    print(names[role])
    for i in range(44):
        e = v[(role * 64 + i) * 8:(role * 64 + i) * 8 + 8]
        if any(e): print(f"  {i:3d} " + " ".join(f"{(x - t0) if x else -1:8d}" for x in e))
