"""Host-side evaluation loop: the B200-native restatement of `fast_eval`
(dlib/utils/utils_trainer.py:533-762) for the hot path only -- forward with the caller-side
window padding of `_forward_with_padding` (:829-862), uint8 quantisation + PSNR / MSE / NRMSE /
SSIM / PSNR_Y (+ ROI-marginalised variants, :874-930) and the metric-sum exchange across
ranks (:653-674).

Multi-GPU: patches are independent, so each rank evaluates an exact contiguous shard of the
patch list (no padding by repetition, unlike DistributedSampler: the result equals the
single-GPU mean) and ONE all-reduce(SUM) of an fp64 vector of 11 values replaces the
reference's 11 all_gathers + barriers (dlib/utils/utils_parallel.py:13-22).
"""
from __future__ import annotations

from typing import Callable, Dict, Optional, Sequence, Tuple

import torch

METRICS = ("psnr", "mse", "nrmse", "ssim", "psnr_y")


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Exact contiguous split of n items over `world` ranks (sizes differ by at most one)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pad_for_windows(lr: torch.Tensor, window_size: int = 8) -> torch.Tensor:
    """Caller-side padding of `_forward_with_padding` (utils_trainer.py:838-852): ALWAYS adds at
    least one window, by mirroring the last rows / columns (edge pixel repeated)."""
    h, w = lr.shape[-2:]
    hp = (h // window_size + 1) * window_size - h
    wp = (w // window_size + 1) * window_size - w
    lr = torch.cat([lr, torch.flip(lr[:, :, h - hp:, :], [2])], 2)
    lr = torch.cat([lr, torch.flip(lr[:, :, :, w - wp:], [3])], 3)
    return lr


def forward_with_padding(net, lr: torch.Tensor, scale: int, swinir_padding: bool) -> torch.Tensor:
    h, w = lr.shape[-2:]
    if swinir_padding:
        e = net(pad_for_windows(lr, getattr(net, "window_size", 8)))
        return e[..., : h * scale, : w * scale]
    return net(lr)


def reduce_sums(vec: torch.Tensor, group=None) -> torch.Tensor:
    """SUM all-reduce of the metric-sum vector (NCCL over NVLink on GPUs, gloo in CPU tests)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(vec, op=dist.ReduceOp.SUM, group=group)
    return vec


def gather_details(block: torch.Tensor, n: int, rank: int, world: int, group=None) -> torch.Tensor:
    """All ranks' per-image blocks (n_r, 10) -> the full (n, 10) block in patch order on every rank: ONE all_gather of
    equally padded blocks (shards differ by at most one row) replaces the reference's per-metric
    sync_dict_across_gpus loop of `_sync_details_across_gpu` (utils_trainer.py:800-826)."""
    import torch.distributed as dist
    if world <= 1 or not (dist.is_available() and dist.is_initialized()):
        return block
    width = -(-n // world)
    pad = torch.zeros(width, block.shape[1], dtype=block.dtype, device=block.device)
    pad[: block.shape[0]] = block
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    rows = []
    for r in range(world):
        lo, hi = shard_range(n, r, world)
        rows.append(parts[r][: hi - lo])
    return torch.cat(rows, 0)


def details_dicts(block: torch.Tensor, ids: Sequence) -> Tuple[dict, dict]:
    """(n, 10) per-image block -> the reference's `details` / `roi_details` dictionaries
    {image id: {metric: value}} (utils_trainer.py:1047-1066)."""
    vals = block.detach().cpu().tolist()
    det, roi = {}, {}
    for i, im_id in enumerate(ids):
        det[im_id] = {m: vals[i][k] for k, m in enumerate(METRICS)}
        roi[im_id] = {m: vals[i][5 + k] for k, m in enumerate(METRICS)}
    return det, roi


def write_details(details: dict, roi_details: Optional[dict], save_dir: str, ds_name: str) -> None:
    """`details_<ds>.yml` / `roi_details_<ds>.yml` as the reference writes them (utils_trainer.py:1140-1147)."""
    import os
    import yaml
    os.makedirs(save_dir, exist_ok=True)
    with open(os.path.join(save_dir, f"details_{ds_name}.yml"), "w") as fd:
        yaml.dump(details, fd)
    if roi_details is not None:
        with open(os.path.join(save_dir, f"roi_details_{ds_name}.yml"), "w") as fd:
            yaml.dump(roi_details, fd)


def write_current_perf_eval(means: Dict[str, float], split: str, ds_name: str, save_dir: Optional[str], name_f: str,
                            current_step: int = 0, current_epoch: int = 0, roi: bool = False) -> dict:
    """Summary file of one evaluation in the shape of the reference's `write_current_perf_eval`
    (dlib/utils/utils_tracker.py:133-165).  A single evaluation has one value per metric, so last == best."""
    import os
    import yaml
    pre = "roi_" if roi else ""
    out = {}
    for m in METRICS:
        out[f"last_{m}"] = float(means[pre + m])
        out[f"best_{m}"] = float(means[pre + m])
    out.update(dataset=ds_name, split=split, current_step=current_step, current_epoch=current_epoch)
    if save_dir is not None:
        os.makedirs(save_dir, exist_ok=True)
        with open(os.path.join(save_dir, name_f), "w") as f:
            yaml.dump(out, f)
    return out


def evaluate_patches(step_fn: Callable[[torch.Tensor, torch.Tensor], torch.Tensor],
                     lr: torch.Tensor, hr: torch.Tensor, batch_size: int,
                     rank: int = 0, world: int = 1, group=None,
                     device: Optional[torch.device] = None, ids: Optional[Sequence] = None) -> Dict[str, float]:
    """Evaluate this rank's shard of (lr, hr) with `step_fn(lr_b, hr_b) -> (b, 10) fp64`
    (5 full-image metrics then 5 ROI-marginalised ones) and return the global means.
    ids: image ids of ALL n patches -> the per-image (n, 10) block is gathered from every rank (one all_gather)
    and returned as the reference's `details` / `roi_details` dictionaries under the keys of the same name."""
    n = lr.shape[0]
    lo, hi = shard_range(n, rank, world)
    dev = device if device is not None else lr.device
    acc = torch.zeros(11, dtype=torch.float64, device=dev)
    blocks = []
    for s in range(lo, hi, batch_size):
        e = min(s + batch_size, hi)
        vals = step_fn(lr[s:e], hr[s:e])
        acc[:10] += vals.to(dev, torch.float64).sum(0)
        acc[10] += e - s
        if ids is not None:
            blocks.append(vals.to(dev, torch.float64))
    # the reference's check_negative_non_float guard (utils_trainer.py:933-958): the step functions OR the per-image
    # flags of every batch into one device word; it is read here, once per sweep (no per-batch host sync)
    fl = getattr(step_fn, "flags", None)
    if fl is not None and fl.get("acc") is not None:
        f = int(fl["acc"].item())
        fl["acc"] = None
        if f & 1:
            raise FloatingPointError("non-finite value (inf/nan) in the SR output or a metric")
        if f & 2:
            raise FloatingPointError("negative metric value")
    acc = reduce_sums(acc, group)
    tot = acc.cpu()
    cnt = float(tot[10])
    out = {"n": cnt}
    for i, m in enumerate(METRICS):
        out[m] = float(tot[i]) / max(cnt, 1.0)
        out["roi_" + m] = float(tot[5 + i]) / max(cnt, 1.0)
    if ids is not None:
        assert len(ids) == n, "one id per patch"
        block = torch.cat(blocks, 0) if blocks else torch.zeros(0, 10, dtype=torch.float64, device=dev)
        full = gather_details(block, n, rank, world, group)
        out["details"], out["roi_details"] = details_dicts(full, ids)
    return out


def make_cuda_step(net, scale: int, swinir_padding: bool, border: Optional[int] = None,
                   roi_ths: Sequence[int] = (4, 5, 6, 7, 8, 9, 10), check: bool = False):
    """The product step: SR forward + single-pass metrics on the current CUDA device."""
    from . import utils_image as UI

    b = scale if border is None else border      # fast_eval: border = args.scale (:562)

    copy_stream = {}

    def step(lr_b: torch.Tensor, hr_b: torch.Tensor) -> torch.Tensor:
        dev = next(net.parameters()).device
        cur = torch.cuda.current_stream(dev)
        lr_b = lr_b.to(dev, non_blocking=True)
        if lr_b.dtype == torch.uint8:                   # uint8 shipping: uint2tensor on the device
            lr_b = lr_b.float().div(255.0)
        if hr_b.device != dev:
            # the target image is 8^2 x larger than the input and only the metrics kernel reads it:
            # its host->device copy runs on a side stream underneath the network forward
            if dev not in copy_stream:
                copy_stream[dev] = torch.cuda.Stream(dev)
            side = copy_stream[dev]
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                hr_b = hr_b.to(dev, non_blocking=True)
            hr_b.record_stream(cur)
            e = forward_with_padding(net, lr_b, scale, swinir_padding)
            cur.wait_stream(side)
        else:
            e = forward_with_padding(net, lr_b, scale, swinir_padding)
        m = UI.compute_metrics(e, hr_b, b, roi_ths, check=check)
        f = m["flags"].max()
        flags["acc"] = f if flags["acc"] is None else torch.maximum(flags["acc"], f)
        cols = [m[k] for k in METRICS]
        if len(roi_ths):
            cols += [m["roi_" + k] for k in METRICS]
        else:
            cols += [torch.zeros_like(cols[0])] * 5
        return torch.stack(cols, 1)

    flags = {"acc": None}
    step.flags = flags
    return step


def make_bicubic_step(scale: int, border: Optional[int] = None,
                      roi_ths: Sequence[int] = (4, 5, 6, 7, 8, 9, 10), check: bool = False, device=None):
    """The model-free baseline sweep of `evaluate()` (utils_trainer.py:1263-1280): bicubic
    up-scaling (Interpolate.forward :120-147) + the same single-pass metrics."""
    from . import utils_image as UI
    from .interpolate import bicubic_upsample

    b = scale if border is None else border

    def step(lr_b: torch.Tensor, hr_b: torch.Tensor) -> torch.Tensor:
        dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        lr_b = lr_b.to(dev, non_blocking=True)
        if lr_b.dtype == torch.uint8:
            lr_b = lr_b.float().div(255.0)
        hr_b = hr_b.to(dev, non_blocking=True)
        m = UI.compute_metrics(bicubic_upsample(lr_b, scale), hr_b, b, roi_ths, check=check)
        f = m["flags"].max()
        flags["acc"] = f if flags["acc"] is None else torch.maximum(flags["acc"], f)
        cols = [m[k] for k in METRICS]
        cols += [m["roi_" + k] for k in METRICS] if len(roi_ths) else [torch.zeros_like(cols[0])] * 5
        return torch.stack(cols, 1)

    flags = {"acc": None}
    step.flags = flags
    return step
