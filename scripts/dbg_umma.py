"""Experiment: shifted-window A operands for tcgen05.mma (see csrc/dbg_umma.cu)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import subprocess
from sr_caco_2_b200 import _lib as L
L.load()                                  # libsrk.so provides encode_map / fail / the error string
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "sr_caco_2_b200")
SO = os.path.join(ROOT, "gpurun_out", "libdbg_umma.so")
os.makedirs(os.path.dirname(SO), exist_ok=True)
# the experiment kernel is NOT part of libsrk.so: built here, linked against it
subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "--expt-relaxed-constexpr",
                       "-shared", "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(PKG, "csrc"),
                       os.path.join(PKG, "csrc", "dbg_umma.cu"), "-o", SO, "-L", PKG, "-l:libsrk.so",
                       "-Xlinker", "-rpath," + PKG])
lib = C.CDLL(SO)
lib.srk_dbg_umma_shift.restype = C.c_int
lib.srk_dbg_umma_shift.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
torch.manual_seed(0)
R = 240
A = torch.randn(R, 64, device="cuda").half()
Bm = torch.randn(64, 64, device="cuda").half()
for pitch in (8, 10, 12):
    for shift in (0, 1, 3, 8, 11, 13):
        if shift + 15 * pitch + 8 > R:
            continue
        rows = torch.tensor([shift + (r // 8) * pitch + r % 8 for r in range(128)], device="cuda")
        exp = A[rows].float() @ Bm.float().t()
        for bo in (0, shift & 7):
            D = torch.zeros(128, 64, device="cuda")
            rc = lib.srk_dbg_umma_shift(A.data_ptr(), R, Bm.data_ptr(), shift, pitch, bo, D.data_ptr(), None)
            torch.cuda.synchronize()
            err = float((D - exp).abs().max())
            print(f"pitch {pitch:2d} shift {shift:2d} base_offset {bo}: rc {rc} max err {err:.4f} {'OK' if err < 0.05 else 'MISMATCH'}", flush=True)
