"""Import shims that let the UNMODIFIED reference (`/root/reference`) be imported in the
build container so that it can pin the oracle port (`oracle/sr_oracle.py`).

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (`sr_caco_2_b200/`) may import this
module; it is used by `tests/golden/make_golden.py` (fixture generation, build container
only) and by the `not gpu` tests that cross-check the oracle port against the live reference
when `/root/reference` happens to be present.  `/root/reference` does not exist on the GPU
box, so `available()` is False there and callers must skip.

Shims (all off the arithmetic path):
  * a bare `dlib` package whose __path__ points at /root/reference/dlib, which skips
    dlib/__init__.py:7-28 (that file drags in the unrelated WSOL stack);
  * `timm.models.layers` with DropPath (identity at eval), to_2tuple and trunc_normal_
    (used at network_swinir.py:14);
  * inert stubs for matplotlib, munch, more_itertools, pretrainedmodels, skimage,
    pydensecrf, bilateralfilter, pynvml.smi, kornia and a top level `utils` package.
"""
import importlib.machinery
import os
import sys
import types

REF_ROOT = os.environ.get("SRK_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "dlib", "models"))


class _Anything(types.ModuleType):
    """Module stub: any attribute is another stub / a no-op callable."""

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        sub = _Anything(self.__name__ + "." + name)
        setattr(self, name, sub)
        return sub

    def __call__(self, *a, **k):
        return None


def _stub(name):
    if name in sys.modules:
        return sys.modules[name]
    m = _Anything(name)
    m.__path__ = []
    m.__spec__ = importlib.machinery.ModuleSpec(name, loader=None, is_package=True)
    sys.modules[name] = m
    parent, _, child = name.rpartition(".")
    if parent:
        setattr(_stub(parent), child, m)
    return m


_installed = False


def install():
    """Idempotently install the shims.  Raises if the reference is not present."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    import torch.nn as nn

    dl = types.ModuleType("dlib")
    dl.__path__ = [os.path.join(REF_ROOT, "dlib")]
    sys.modules["dlib"] = dl

    class DropPath(nn.Module):  # eval-mode identity (drop_path only matters when training)
        def __init__(self, p=0.0):
            super().__init__()

        def forward(self, x):
            return x

    try:
        import timm.models.layers  # noqa: F401  (use the real one if it is installed)
    except Exception:
        layers = types.ModuleType("timm.models.layers")
        layers.DropPath = DropPath
        layers.trunc_normal_ = nn.init.trunc_normal_
        layers.to_2tuple = lambda x: tuple(x) if isinstance(x, (tuple, list)) else (x, x)
        timm = types.ModuleType("timm")
        timm_models = types.ModuleType("timm.models")
        timm.models = timm_models
        timm_models.layers = layers
        sys.modules["timm"] = timm
        sys.modules["timm.models"] = timm_models
        sys.modules["timm.models.layers"] = layers

    for name in ["matplotlib", "matplotlib.pyplot", "matplotlib.colors", "matplotlib.patches",
                 "matplotlib.cm", "matplotlib.font_manager", "matplotlib.ticker",
                 "mpl_toolkits", "mpl_toolkits.mplot3d", "mpl_toolkits.axes_grid1",
                 "munch", "more_itertools", "pretrainedmodels", "pretrainedmodels.utils",
                 "skimage", "skimage.filters", "skimage.metrics", "skimage.transform",
                 "skimage.io", "skimage.util", "skimage.exposure",
                 "pydensecrf", "pydensecrf.densecrf", "pydensecrf.utils", "bilateralfilter",
                 "kornia", "tifffile", "utils"]:
        try:
            __import__(name)
        except Exception:
            _stub(name)
    try:
        import pynvml.smi  # noqa: F401
    except Exception:
        _stub("pynvml.smi")
    _installed = True


def swinir_class():
    install()
    from dlib.models.network_swinir import SwinIR
    return SwinIR


def nlsn_primitives():
    install()
    from dlib.models.network_nlsn import default_conv, ResBlock, Upsampler
    return default_conv, ResBlock, Upsampler


def utils_image():
    install()
    from dlib.utils import utils_image as ui
    return ui
