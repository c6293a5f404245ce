"""ctypes binding of libsrk.so (see include/srk.h).  There is NO fallback: if the CUDA
library is missing or the device is not an sm_100 part, every entry point raises."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libsrk.so")

SRK_BF16, SRK_FP16 = 0, 1
ACT_NONE, ACT_GELU, ACT_LRELU, ACT_RELU, ACT_LRELU02 = 0, 1, 2, 3, 4
A_ROWS, A_CONV3X3 = 0, 1
O16_ROWS, O16_PIXSHUF2 = 0, 1
ENGINE_TCGEN05, ENGINE_MMA_SYNC = 0, 1
UPSAMPLER_PIXELSHUFFLE, UPSAMPLER_PIXELSHUFFLEDIRECT, UPSAMPLER_NEAREST_CONV = 0, 1, 2
MET_PSNR, MET_MSE, MET_NRMSE, MET_SSIM, MET_PSNR_Y, MET_N = 0, 1, 2, 3, 4, 5
MAX_ROI_THS = 8
OPT_NO_FUSED_MLP, OPT_NO_FUSED_ATTN, OPT_NO_FOLD_TAIL, OPT_NO_FOLD_QKV_BIAS, OPT_NO_FUSED_BLOCK, OPT_NO_GRAPH = 1, 2, 4, 8, 16, 32

vp, fp, ip = C.c_void_p, C.c_void_p, C.c_void_p   # all device pointers travel as void*


class GemmArgs(C.Structure):
    _fields_ = [("A", vp), ("a_mode", C.c_int), ("lda", C.c_int), ("nB", C.c_int), ("H", C.c_int),
                ("W", C.c_int), ("Wt", vp), ("M", C.c_int), ("N", C.c_int), ("K", C.c_int),
                ("dtype", C.c_int), ("bias", fp), ("act", C.c_int), ("res", fp),
                ("res_scale", C.c_float), ("out32", fp), ("ld32", C.c_int), ("win_shift", C.c_int),
                ("out16", vp), ("ld16", C.c_int), ("out16_dtype", C.c_int), ("out16_mode", C.c_int),
                ("ln_g", fp), ("ln_b", fp), ("ln_C", C.c_int), ("ln_win_shift", C.c_int),
                ("img", fp), ("img_s", C.c_int), ("img_scale", C.c_float), ("img_hc", C.c_int),
                ("img_wc", C.c_int), ("attn_table", fp), ("attn_heads", C.c_int),
                ("attn_scale", C.c_float), ("attn_shift", C.c_int), ("conv_k", C.c_int), ("ln_pad_one", C.c_int)]


class MlpArgs(C.Structure):
    _fields_ = [("A", vp), ("lda", C.c_int), ("M", C.c_int), ("C", C.c_int), ("Cp", C.c_int),
                ("hid_p", C.c_int), ("W1", vp), ("b1", fp), ("W2", vp), ("b2", fp), ("res", fp),
                ("out32", fp), ("ld32", C.c_int), ("out16", vp), ("ld16", C.c_int),
                ("out16_dtype", C.c_int), ("ln_g", fp), ("ln_b", fp), ("ln_C", C.c_int),
                ("ln_win_shift", C.c_int), ("H", C.c_int), ("W", C.c_int), ("ln_pad_one", C.c_int)]


class AttnBlockArgs(C.Structure):
    _fields_ = [("A", vp), ("lda", C.c_int), ("M", C.c_int), ("C", C.c_int), ("Cp", C.c_int), ("H", C.c_int),
                ("W", C.c_int), ("shift", C.c_int), ("num_heads", C.c_int), ("Wqkv", vp), ("Wproj", vp),
                ("b_proj", fp), ("rel_table", fp), ("scale", C.c_float), ("res", fp), ("out32", fp),
                ("ld32", C.c_int), ("out16", vp), ("ld16", C.c_int), ("out16_dtype", C.c_int),
                ("ln_g", fp), ("ln_b", fp), ("ln_C", C.c_int)]


class ConvParams(C.Structure):
    _fields_ = [("w", vp), ("b", fp), ("cin_p", C.c_int), ("n_p", C.c_int)]


class TailFold(C.Structure):
    _fields_ = [("w", vp), ("b", fp), ("border_w", vp), ("border_b", fp), ("w_scale", C.c_float)]


class StbParams(C.Structure):
    _fields_ = [("ln1_g", fp), ("ln1_b", fp), ("ln2_g", fp), ("ln2_b", fp),
                ("w_qkv", vp), ("w_proj", vp), ("w_fc1", vp), ("w_fc2", vp),
                ("b_qkv", fp), ("b_proj", fp), ("b_fc1", fp), ("b_fc2", fp),
                ("rel_table", fp), ("shift", C.c_int), ("num_heads", C.c_int), ("w_qkv_fb", vp), ("w_qkv_hm", vp)]


class SwinIRPlan(C.Structure):
    _fields_ = [("upscale", C.c_int), ("in_chans", C.c_int), ("window_size", C.c_int),
                ("embed_dim", C.c_int), ("hidden_dim", C.c_int), ("n_layers", C.c_int),
                ("upsampler", C.c_int), ("img_range", C.c_float),
                ("Cp", C.c_int), ("hid_p", C.c_int), ("dp", C.c_int), ("ao_p", C.c_int),
                ("depths", C.POINTER(C.c_int)), ("stbs", C.POINTER(StbParams)),
                ("rstb_convs", C.POINTER(ConvParams)),
                ("conv_first_w", fp), ("conv_first_b", fp),
                ("pe_norm_g", fp), ("pe_norm_b", fp), ("norm_g", fp), ("norm_b", fp),
                ("conv_after_body", ConvParams), ("conv_before_upsample", ConvParams),
                ("upsample", ConvParams * 4), ("n_upsample", C.c_int),
                ("conv_last_w", fp), ("conv_last_b", C.c_float),
                ("linear_dtype", C.c_int), ("conv_dtype", C.c_int), ("tail_fold", TailFold),
                ("resi_3conv", C.c_int), ("rstb_c0", C.POINTER(ConvParams)), ("rstb_c1", C.POINTER(ConvParams)),
                ("cab_c0", ConvParams), ("cab_c1", ConvParams), ("conv_hr", ConvParams), ("options", C.c_int)]


class EDSRPlan(C.Structure):
    _fields_ = [("in_chans", C.c_int), ("n_resblocks", C.c_int), ("n_feats", C.c_int),
                ("scale", C.c_int), ("res_scale", C.c_float), ("rgb_range", C.c_float),
                ("Fp", C.c_int), ("head_w", fp), ("head_b", fp),
                ("body", C.POINTER(ConvParams)), ("tail_up", ConvParams * 4),
                ("n_tail_up", C.c_int), ("tail_w", fp), ("tail_b", C.c_float),
                ("conv_dtype", C.c_int), ("tail_fold", TailFold), ("options", C.c_int)]


# every symbol include/srk.h declares: (restype, argtypes)
PROTOTYPES = {
    "srk_last_error": (C.c_char_p, []),
    "srk_version": (C.c_int, []),
    "srk_check_device": (C.c_int, [C.c_int]),
    "srk_set_engine": (C.c_int, [C.c_int]),
    "srk_get_engine": (C.c_int, []),
    "srk_index_map": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp]),
    "srk_metrics_scratch_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "srk_metrics_use_tile_kernel": (C.c_int, [C.c_int]),
    "srk_gemm_conv_halo": (C.c_int, [C.c_int]),
    "srk_metrics": (C.c_int, [fp, fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                              C.POINTER(C.c_int), C.c_int, vp, vp, vp, vp]),
    "srk_metrics_roi": (C.c_int, [fp, fp, fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp,
                                  vp, vp]),
    "srk_metrics_h8": (C.c_int, [fp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int, vp, vp,
                                 vp, vp]),
    "srk_bicubic_upsample": (C.c_int, [fp, C.c_int, C.c_int, C.c_int, C.c_int, fp, vp]),
    "srk_gemm": (C.c_int, [C.POINTER(GemmArgs), vp]),
    "srk_mlp": (C.c_int, [C.POINTER(MlpArgs), vp]),
    "srk_attn_block": (C.c_int, [C.POINTER(AttnBlockArgs), vp]),
    "srk_layernorm": (C.c_int, [fp, C.c_int, C.c_int, C.c_int, fp, fp, C.c_float, vp, C.c_int,
                                C.c_int, fp, C.c_int, C.c_int, C.c_int, vp]),
    "srk_window_attention": (C.c_int, [vp, C.c_int, vp, C.c_int, fp, C.c_int, C.c_int, C.c_int,
                                       C.c_int, C.c_int, C.c_float, C.c_int, vp]),
    "srk_tail_border": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(TailFold), C.c_float, fp,
                                  C.c_int, C.c_int, vp]),
    "srk_conv_in": (C.c_int, [fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, fp, fp,
                              C.c_int, fp, C.c_int, vp, C.c_int, C.c_int, vp]),
    "srk_conv_in_ln": (C.c_int, [fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, fp, fp,
                                 C.c_int, fp, fp, C.c_int, fp, fp, fp, fp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp]),
    "srk_conv_out": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.c_float,
                               C.c_float, fp, C.c_int, C.c_int, vp]),
    "srk_swinir_workspace_bytes": (C.c_size_t, [C.POINTER(SwinIRPlan), C.c_int, C.c_int, C.c_int]),
    "srk_swinir_forward": (C.c_int, [C.POINTER(SwinIRPlan), fp, fp, C.c_int, C.c_int, C.c_int, vp,
                                     C.c_size_t, vp]),
    "srk_edsr_workspace_bytes": (C.c_size_t, [C.POINTER(EDSRPlan), C.c_int, C.c_int, C.c_int]),
    "srk_edsr_forward": (C.c_int, [C.POINTER(EDSRPlan), fp, fp, C.c_int, C.c_int, C.c_int, vp,
                                   C.c_size_t, vp]),
    "srk_launch_count": (C.c_longlong, [C.c_int]),
    "srk_profile": (C.c_int, [C.c_int]),
    "srk_profile_read": (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_longlong), C.c_int]),
}


class SrkError(RuntimeError):
    pass


_lib = None


def load():
    """Load libsrk.so (raises if it has not been built: there is no CPU / eager fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SrkError(f"{LIB_PATH} not found: build it with `python -m sr_caco_2_b200.build` "
                       "(the CUDA path is the only path; there is no fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)           # AttributeError if a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int):
    if rc != 0:
        msg = load().srk_last_error().decode(errors="replace")
        codes = {-1: "invalid argument", -2: "unsupported", -3: "CUDA error", -4: "workspace",
                 -5: "wrong architecture"}
        if rc == -2:
            raise NotImplementedError(f"libsrk: {msg}")
        if rc == -1:
            raise ValueError(f"libsrk: {msg}")
        raise SrkError(f"libsrk [{codes.get(rc, rc)}]: {msg}")


_device_ok = set()


def require_device(t):
    """The tensor must live on an sm_100 CUDA device; anything else is an error."""
    import torch
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise SrkError("sr_caco_2_b200 runs on CUDA (sm_100a) tensors only; got a "
                       f"{'CPU tensor' if isinstance(t, torch.Tensor) else type(t)} "
                       "(there is no CPU fallback)")
    idx = t.device.index if t.device.index is not None else torch.cuda.current_device()
    if idx not in _device_ok:
        check(load().srk_check_device(idx))
        _device_ok.add(idx)
    return idx


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def set_engine(name: str):
    eng = {"tcgen05": ENGINE_TCGEN05, "mma_sync": ENGINE_MMA_SYNC}[name]
    check(load().srk_set_engine(eng))


def get_engine() -> str:
    return {ENGINE_TCGEN05: "tcgen05", ENGINE_MMA_SYNC: "mma_sync"}[load().srk_get_engine()]


_graph_launches = 0          # kernel launches replayed from CUDA graphs (the library only counts what it launches itself)


def add_graph_launches(n: int):
    global _graph_launches
    _graph_launches += int(n)


def launch_count(reset=False) -> int:
    """Kernels of this library launched on the calling thread since the last reset, including graph replays."""
    global _graph_launches
    n = int(load().srk_launch_count(1 if reset else 0)) + _graph_launches
    if reset:
        _graph_launches = 0
    return n


PROF_FAMILIES = ("gemm", "attention", "layernorm", "conv_in", "conv_out", "metrics", "gemm_res_ln", "attn_block", "mlp")


profiling = False            # per-launch event timing is on: forwards run eagerly (a graph replay has no per-kernel events)


def profile(enable: bool):
    global profiling
    check(load().srk_profile(1 if enable else 0))
    profiling = bool(enable)


def profile_read(reset=True):
    ms = (C.c_double * len(PROF_FAMILIES))()
    calls = (C.c_longlong * len(PROF_FAMILIES))()
    check(load().srk_profile_read(ms, calls, 1 if reset else 0))
    return ({k: float(ms[i]) for i, k in enumerate(PROF_FAMILIES)},
            {k: int(calls[i]) for i, k in enumerate(PROF_FAMILIES)})
