"""Per-shape timing of the GEMM calls of cfg3 (B=32, 64x64): us, TFLOP/s, GB/s of algorithmic bytes."""
import ctypes as C, sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sr_caco_2_b200 import _lib as L
L.load()
dev = "cuda:0"
eng = sys.argv[1] if len(sys.argv) > 1 else "tcgen05"
L.set_engine(eng)
B, H, W = int(os.environ.get("SRK_B", 32)), 64, 64
M = B * H * W

def run(name, **kw):
    g = L.GemmArgs(); g.res_scale, g.win_shift, g.ln_win_shift = 1.0, -1, -1
    keep = []
    for k, v in kw.items():
        if isinstance(v, torch.Tensor): keep.append(v); v = L.ptr(v)
        setattr(g, k, v)
    lib = L.load()
    if os.environ.get("SRK_PROFILE_ONCE"):
        L.check(lib.srk_gemm(C.byref(g), L.stream_ptr())); torch.cuda.synchronize(); return 1.0
    for _ in range(3): L.check(lib.srk_gemm(C.byref(g), L.stream_ptr()))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record()
    for _ in range(n): L.check(lib.srk_gemm(C.byref(g), L.stream_ptr()))
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3  # us

def t16(*shape, dt=torch.bfloat16): return (torch.randn(*shape, device=dev) * 0.1).to(dt)
rows = []
def report(name, us, flops, bytes_):
    rows.append(dict(name=name, us=round(us, 1), tflops=round(flops / us / 1e6, 1), gbs=round(bytes_ / us / 1e3, 1)))
    print(f"{name:28s} {us:9.1f} us  {flops/us/1e6:8.1f} TFLOP/s  {bytes_/us/1e3:8.1f} GB/s", flush=True)

Cp = 192
A = t16(M, Cp); bias = torch.zeros(576, device=dev)
X = torch.randn(M, Cp, device=dev); X2 = torch.empty_like(X)
lng = torch.ones(180, device=dev); lnb = torch.zeros(180, device=dev)
# qkv
Wq = t16(576, Cp); QKV = torch.empty(M, 576, device=dev, dtype=torch.bfloat16)
us = run("qkv", A=A, a_mode=0, lda=Cp, nB=B, H=H, W=W, Wt=Wq, M=M, N=576, K=Cp, dtype=0, bias=bias, out16=QKV, ld16=576, out16_dtype=0)
report("qkv N576 K192 ->bf16", us, 2*M*576*192, M*(384+1152))
Wp = t16(192, 192); A16 = torch.empty(M, Cp, device=dev, dtype=torch.bfloat16)
us = run("proj", A=A, a_mode=0, lda=Cp, nB=B, H=H, W=W, Wt=Wp, M=M, N=192, K=192, dtype=0, bias=bias, res=X, out32=X2, ld32=Cp, win_shift=4,
         ln_g=lng, ln_b=lnb, ln_C=180, out16=A16, ld16=Cp, out16_dtype=0)
report("proj N192 K192 +res+LN", us, 2*M*192*192, M*(384+768+768+384))
us = run("proj_nores", A=A, a_mode=0, lda=Cp, nB=B, H=H, W=W, Wt=Wp, M=M, N=192, K=192, dtype=0, bias=bias, out16=A16, ld16=Cp, out16_dtype=0)
report("  (N192 K192 ->bf16 only)", us, 2*M*192*192, M*(384+384))
W1 = t16(384, 192); HID = torch.empty(M, 384, device=dev, dtype=torch.bfloat16)
us = run("fc1", A=A, a_mode=0, lda=Cp, nB=B, H=H, W=W, Wt=W1, M=M, N=384, K=192, dtype=0, bias=bias, act=1, out16=HID, ld16=384, out16_dtype=0)
report("fc1 N384 K192 GELU->bf16", us, 2*M*384*192, M*(384+768))
W2 = t16(192, 384)
us = run("fc2", A=HID, a_mode=0, lda=384, nB=B, H=H, W=W, Wt=W2, M=M, N=192, K=384, dtype=0, bias=bias, res=X, out32=X2, ld32=Cp,
         ln_g=lng, ln_b=lnb, ln_C=180, ln_win_shift=4, out16=A16, ld16=Cp, out16_dtype=0)
report("fc2 N192 K384 +res+LN", us, 2*M*192*384, M*(768+768+768+384))
# fused MLP
def run_mlp():
    m = L.MlpArgs()
    m.A, m.lda, m.M, m.C, m.Cp, m.hid_p = L.ptr(A), Cp, M, 180, Cp, 384
    m.W1, m.b1, m.W2, m.b2 = L.ptr(W1), L.ptr(bias), L.ptr(W2), L.ptr(bias)
    m.res, m.out32, m.ld32, m.out16, m.ld16 = L.ptr(X), L.ptr(X2), Cp, L.ptr(A16), Cp
    m.H, m.W, m.out16_dtype, m.ln_g, m.ln_b, m.ln_C, m.ln_win_shift = H, W, 0, L.ptr(lng), L.ptr(lnb), 180, 4
    lib = L.load()
    if os.environ.get("SRK_PROFILE_ONCE"):
        L.check(lib.srk_mlp(C.byref(m), L.stream_ptr())); torch.cuda.synchronize(); return 1.0
    for _ in range(3): L.check(lib.srk_mlp(C.byref(m), L.stream_ptr()))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): L.check(lib.srk_mlp(C.byref(m), L.stream_ptr()))
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 10 * 1e3
if eng == "tcgen05":
    us = run_mlp()
    report("fused MLP 192-384-192 +LN", us, 2*M*192*384*2, M*(384+768+768+384))
def run_attn_block():
    a = L.AttnBlockArgs()
    whm = t16(576, Cp); tab = torch.randn(6, 225, device=dev) * 0.3
    a.A, a.lda, a.M, a.C, a.Cp, a.H, a.W, a.shift, a.num_heads = L.ptr(A), Cp, M, 180, Cp, H, W, 4, 6
    a.Wqkv, a.Wproj, a.b_proj, a.rel_table, a.scale = L.ptr(whm), L.ptr(Wp), L.ptr(bias), L.ptr(tab), 30 ** -0.5
    a.res, a.out32, a.ld32, a.out16, a.ld16, a.out16_dtype = L.ptr(X), L.ptr(X2), Cp, L.ptr(A16), Cp, 0
    a.ln_g, a.ln_b, a.ln_C = L.ptr(lng192), L.ptr(lnb192), 180
    keep.extend([whm, tab])
    lib = L.load()
    if os.environ.get("SRK_PROFILE_ONCE"):
        L.check(lib.srk_attn_block(C.byref(a), L.stream_ptr())); torch.cuda.synchronize(); return 1.0
    for _ in range(3): L.check(lib.srk_attn_block(C.byref(a), L.stream_ptr()))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): L.check(lib.srk_attn_block(C.byref(a), L.stream_ptr()))
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 10 * 1e3
keep = []
lng192 = torch.ones(192, device=dev); lnb192 = torch.zeros(192, device=dev)
if eng == "tcgen05":
    us = run_attn_block()
    report("attn block (qkv+attn+proj+LN)", us, 2*M*(192*576 + 2*64*192 + 192*192), M*(384+768+768+384))
Ah = t16(M, Cp, dt=torch.float16); Wc = t16(192, 9*192, dt=torch.float16)
us = run("convCC", A=Ah, a_mode=1, lda=Cp, nB=B, H=H, W=W, Wt=Wc, M=M, N=192, K=9*192, dtype=1, bias=bias, res=X, out32=X2, ld32=Cp)
report("conv 192->192 +res", us, 2*M*192*1728, M*(384+768+768))
Wb = t16(64, 9*192, dt=torch.float16); U0 = torch.empty(M, 64, device=dev, dtype=torch.float16)
us = run("before_up", A=Ah, a_mode=1, lda=Cp, nB=B, H=H, W=W, Wt=Wb, M=M, N=64, K=9*192, dtype=1, bias=bias, act=2, out16=U0, ld16=64, out16_dtype=1)
report("conv 192->64 lrelu", us, 2*M*64*1728, M*(384+128))
Wu = t16(256, 9*64, dt=torch.float16)
for k, (hh, ww) in enumerate([(64, 64), (128, 128), (256, 256)]):
    Mi = B*hh*ww
    Ui = t16(Mi, 64, dt=torch.float16); Uo = torch.empty(Mi*4, 64, device=dev, dtype=torch.float16)
    us = run("up", A=Ui, a_mode=1, lda=64, nB=B, H=hh, W=ww, Wt=Wu, M=Mi, N=256, K=576, dtype=1, bias=bias, out16=Uo, ld16=64, out16_dtype=1, out16_mode=1)
    report(f"up{k} 64->256 ps @{hh}", us, 2*Mi*256*576, Mi*(128+512))
    del Ui, Uo
# EDSR body conv: 64 -> 64 on 64 x (64 x 64), 16-bit in / out, resident weights; 5x5 folded tail conv 64 -> 64 at 64x64
Be = 64
Me = Be * 64 * 64
Ae = t16(Me, 64, dt=torch.float16); We = t16(64, 9*64, dt=torch.float16); Oe = torch.empty(Me, 64, device=dev, dtype=torch.float16)
if eng == "tcgen05":
    for halo in (1, 0):
        L.load().srk_gemm_conv_halo(halo)
        us = run("edsr", A=Ae, a_mode=1, lda=64, nB=Be, H=64, W=64, Wt=We, M=Me, N=64, K=576, dtype=1, bias=bias, act=3, out16=Oe, ld16=64, out16_dtype=1)
        report(f"conv 64->64 relu B=64 halo={halo}", us, 2*Me*64*576, Me*(128+128))
        us = run("before_up", A=Ah, a_mode=1, lda=Cp, nB=B, H=H, W=W, Wt=Wb, M=M, N=64, K=9*192, dtype=1, bias=bias, act=2, out16=U0, ld16=64, out16_dtype=1)
        report(f"conv 192->64 lrelu halo={halo}", us, 2*M*64*1728, M*(384+128))
    L.load().srk_gemm_conv_halo(1)
# folded reconstruction tail of cfg3: 5x5 conv 64 -> 64 (= 8 x 8 sub-pixels) with the image epilogue, 32 x (64 x 64) -> 32 x 512 x 512
At = t16(M, 64, dt=torch.float16); Wt5 = t16(64, 25*64, dt=torch.float16); img = torch.empty(B, 1, 8*H, 8*W, device=dev)
for halo in ((1, 0) if eng == "tcgen05" else (1,)):
    if eng == "tcgen05": L.load().srk_gemm_conv_halo(halo)
    us = run("tail5", A=At, a_mode=1, conv_k=5, lda=64, nB=B, H=H, W=W, Wt=Wt5, M=M, N=64, K=25*64, dtype=1, bias=bias, img=img, img_s=8, img_scale=1.0, img_hc=8*H, img_wc=8*W)
    report(f"folded tail 5x5 64->8x8 img halo={halo}", us, 2*M*64*1600, M*(128+256))
if eng == "tcgen05": L.load().srk_gemm_conv_halo(1)
json.dump(rows, open(f"gpurun_out/gemm_bench_{eng}.json", "w"), indent=1)
