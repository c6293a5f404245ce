"""The drop-in boundary exercised through the REFERENCE'S OWN callers (CPU only, build container only: skipped when
/root/reference is absent).  The same `args` object the reference builds (`get_config('swinir')` -> `Dict2Obj`) is
handed to both `define_G`s; a `{'params': state_dict}` checkpoint goes through the reference's own
`ModelBase.load_network` code into this package's module; the metric shims keep the reference's signatures and the
argument conventions `_compute_metrics` uses (dlib/utils/utils_trainer.py:998-1030).

Reference call sites: eval.py:116-139, dlib/models/select_network.py:24-50, dlib/models/model_base.py:182-200,
dlib/models/model_plain.py:178-200, 398-402."""
import copy
import inspect
import os
import sys
import types

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import ref_import as R  # noqa: E402

pytestmark = pytest.mark.skipif(not R.available(), reason="the reference tree is only present in the build container")


def _ref_args(scale, in_chans=1, img_size=None):
    R.install()
    import dlib.dllogger as DLLogger
    try:
        DLLogger.init_arb(backends=[], is_master=False)             # define_G logs the parameter count
    except Exception:
        pass
    from dlib.utils.utils_config import get_config
    from dlib.utils.tools import Dict2Obj
    cfg = copy.deepcopy(get_config("swinir"))
    cfg["scale"] = scale
    cfg["netG"]["swinir_upscale"] = scale
    cfg["netG"]["swinir_in_chans"] = in_chans                      # 1-channel microscopy patches (main.py --n_channels 1)
    cfg["netG"]["swinir_img_size"] = img_size or 128 // scale     # h_size // scale (utils_parser.py)
    return Dict2Obj(cfg)


@pytest.mark.parametrize("scale", [2, 4, 8])
def test_same_args_object_builds_the_same_network(scale):
    """define_G(args) of the reference and of this package on the SAME object: identical state_dict keys, shapes,
    dtypes, parameter count."""
    R.install()
    from dlib.models.select_network import define_G as ref_define_G
    import sr_caco_2_b200 as S
    args = _ref_args(scale)
    ref = ref_define_G(args)
    ours = S.define_G(args)
    sr, so = ref.state_dict(), ours.state_dict()
    assert list(sr.keys()) == list(so.keys())
    for k in sr:
        assert sr[k].shape == so[k].shape and sr[k].dtype == so[k].dtype, k
    assert sum(p.numel() for p in ref.parameters()) == sum(p.numel() for p in ours.parameters())
    assert [n for n, _ in ref.named_buffers()] == [n for n, _ in ours.named_buffers()]
    # the buffers the reference derives at construction hold the same values
    for (n, a), (_, b) in zip(ref.named_buffers(), ours.named_buffers()):
        assert torch.equal(a, b), n


def test_reference_load_network_loads_a_params_checkpoint(tmp_path):
    """`{'params': sd}` file -> the reference's own ModelBase.load_network (strict and non-strict branches) -> this
    package's module; the loaded values equal the checkpoint."""
    R.install()
    from dlib.models.model_base import ModelBase
    from dlib.models.select_network import define_G as ref_define_G
    import sr_caco_2_b200 as S
    args = _ref_args(4)
    ref = ref_define_G(args)
    path = str(tmp_path / "10000_G.pth")
    torch.save({"params": ref.state_dict()}, path)
    stub = types.SimpleNamespace(get_bare_model=lambda n: n)
    for strict in (True, False):
        ours = S.define_G(args)
        ModelBase.load_network(stub, path, ours, strict=strict, param_key="params")
        for k, v in ref.state_dict().items():
            assert torch.equal(ours.state_dict()[k], v), (strict, k)
    # a bare state_dict file (no 'params' key) loads as well, as the reference's branch does
    torch.save(ref.state_dict(), path)
    ours = S.define_G(args)
    ModelBase.load_network(stub, path, ours, strict=True, param_key="params")
    assert torch.equal(ours.state_dict()["conv_first.weight"], ref.state_dict()["conv_first.weight"])


def test_metric_shims_keep_the_reference_signatures():
    """`_compute_metrics` calls the shims positionally (E, H) with border= / roi= keywords and, for NRMSE, with
    img= / y= keywords (utils_trainer.py:998-1030): names, order and defaults must match the reference's."""
    UIr = R.utils_image()
    from sr_caco_2_b200 import utils_image as UI
    for name in ("mbatch_gpu_calculate_psnr", "mbatch_gpu_calculate_mse", "mbatch_gpu_calculate_nrmse",
                 "mbatch_gpu_calculate_ssim", "tensor2uint82float"):
        pr = inspect.signature(getattr(UIr, name)).parameters
        po = inspect.signature(getattr(UI, name)).parameters
        assert list(pr.keys()) == list(po.keys()), name
        for k in pr:
            assert pr[k].default == po[k].default or (pr[k].default is inspect._empty and po[k].default is inspect._empty), (name, k)
    # tensor2uint82float is the same function on CPU tensors
    x = torch.rand(2, 1, 9, 7) * 1.4 - 0.2
    assert torch.equal(UI.tensor2uint82float(x), UIr.tensor2uint82float(x))


def test_cpu_tensors_and_unbuilt_networks_fail_loudly():
    """No silent fallback: a CPU forward raises, an RGB network raises at construction."""
    import sr_caco_2_b200 as S
    from sr_caco_2_b200 import _lib as L
    args = _ref_args(2)
    net = S.define_G(args).eval()
    with pytest.raises(L.SrkError):
        net(torch.rand(1, 1, 16, 16))
    with pytest.raises(NotImplementedError):
        S.define_G(_ref_args(2, in_chans=3))
