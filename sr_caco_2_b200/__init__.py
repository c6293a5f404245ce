"""B200-native (sm_100a) evaluation hot path of sbelharbi/sr-caco-2: SwinIR / EDSR forward and
PSNR / SSIM / NRMSE scoring behind the reference's constructor and metric APIs.
See DESIGN.md (path, layouts, kernels) and INTEGRATION.md (how the reference binds to it)."""
from . import _lib
from ._lib import SrkError, get_engine, launch_count, set_engine
from .interpolate import Interpolate, bicubic_upsample
from .network_edsr import EDSR, EDSR_LIIF
from .network_swinir import SwinIR
from .select_network import define_G

__all__ = ["SwinIR", "EDSR", "EDSR_LIIF", "define_G", "Interpolate", "bicubic_upsample", "SrkError", "set_engine", "get_engine",
           "launch_count"]
