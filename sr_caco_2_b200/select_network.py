"""`define_G(args)` with the reference's contract (dlib/models/select_network.py:19-291) for the
two net families on the hot path: reads `args.netG['net_type']` and the `'<nt>_*'` hyper
parameters exactly as the reference does (:24-50) and returns an nn.Module."""
from __future__ import annotations

import re

SWINIR = "swinir"          # dlib/utils/constants.py:32
EDSR_LIIF = "EDSR_LIIF"    # dlib/utils/constants.py:39


def _safe(name: str) -> str:  # dlib/utils/shared.py:272-277 safe_str_var
    out = re.sub("[^0-9a-zA-Z_]", "_", name)
    return "_" + out if out[0].isdigit() else out


def define_G(args):
    opt = args.netG if hasattr(args, "netG") else args["netG"]
    net_type = opt["net_type"]
    nt = _safe(net_type)
    if net_type == SWINIR:
        from .network_swinir import SwinIR
        return SwinIR(upscale=opt[f"{nt}_upscale"], in_chans=opt[f"{nt}_in_chans"],
                      img_size=opt[f"{nt}_img_size"], window_size=opt[f"{nt}_window_size"],
                      img_range=opt[f"{nt}_img_range"], depths=opt[f"{nt}_depths"],
                      embed_dim=opt[f"{nt}_embed_dim"], num_heads=opt[f"{nt}_num_heads"],
                      mlp_ratio=opt[f"{nt}_mlp_ratio"], upsampler=opt[f"{nt}_upsampler"],
                      resi_connection=opt[f"{nt}_resi_connection"])
    if net_type == EDSR_LIIF:
        from .network_edsr import EDSR_LIIF as net
        return net(in_chans=opt[f"{nt}_in_chans"], n_resblocks=opt[f"{nt}_n_resblocks"],
                   n_feats=opt[f"{nt}_n_feats"], scale=opt[f"{nt}_upscale"],
                   rgb_range=opt[f"{nt}_img_range"], local_ensemble=True, feat_unfold=True,
                   cell_decode=True)
    raise NotImplementedError(f"net_type {net_type!r}: only the SwinIR and EDSR families are on "
                              "the hot path this package accelerates")
