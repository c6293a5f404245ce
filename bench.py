#!/usr/bin/env python
"""Benchmark of the SR-CACO-2 evaluation hot path on B200 (contract: see the task statement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--workload cfg3|cfg1|cfg2|cfg4|cfg5] [--engine tcgen05|mma_sync] [--sweep N_PATCHES]

--sweep N_PATCHES switches to the STRONG-scaling form of BASELINE.json configs[3]: a fixed list of N patches is
sharded over the ranks (ragged shards, tail batches), every rank evaluates its shard, ONE all-reduce of the 11
metric sums, and rank 0 re-evaluates the whole list alone (untimed) to assert that the sharded means equal the
single-GPU means.

A "step" is one pass of the hot path over one batch of synthetic patches: SwinIR forward
(SwinIR-classical X8, 64x64 -> 512x512, batch 32 per GPU = BASELINE.json configs[2]) followed by
uint8 quantisation + PSNR / MSE / NRMSE / SSIM / PSNR_Y (and the 7 ROI thresholds) against the HR
target.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ROI_THS = (4, 5, 6, 7, 8, 9, 10)


def synthetic_batch(B, h, w, scale, seed):
    """LR in [0,1) and an HR target stored as uint8-representable values (what the loader delivers)."""
    g = torch.Generator().manual_seed(seed)
    lr = torch.rand(B, 1, h, w, generator=g)
    hr = (torch.rand(B, 1, h * scale, w * scale, generator=g) * 255).round() / 255
    return lr, hr


def read_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"tensor_burst": d.get("bf16_tflops"), "tensor_sustained": d.get("bf16_tflops_sustained"),
                "hbm": d.get("hbm_gbs"), "source": "measured (MEASURED_PEAKS.json)"}
    return {"tensor_burst": 1590.0, "tensor_sustained": 1400.0, "hbm": 6650.0,
            "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """SM clock / power / throttle reasons sampled DURING the timed region: NVML polled every
    ~2 ms from a thread (nvidia-smi's own loop starts too slowly for a 100 ms region); falls
    back to one nvidia-smi query when NVML is unavailable."""
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20),
               ("sw_power_cap", 0x4))

    def __init__(self, index):
        self.index, self.samples, self.stop_flag, self.t, self.nv = index, [], False, None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _poll(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                try:
                    rs = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    rs = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.samples.append((sm, pw, rs))
            except Exception:
                break
            time.sleep(0.002)

    def start(self):
        if self.nv is not None:
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()

    def stop(self):
        if self.nv is None:
            return self._smi_once()
        self.stop_flag = True
        self.t.join(timeout=2)
        if not self.samples:
            return self._smi_once()
        sm = sorted(x[0] for x in self.samples)
        bits = 0
        for x in self.samples:
            bits |= x[2]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.max_sm, "power_w_max": max(x[1] for x in self.samples),
                "samples": len(sm), "source": "nvml", "reasons": sorted(n for n, b in self.REASONS if bits & b)}

    def _smi_once(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            out = subprocess.check_output(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}",
                                           "--format=csv,noheader,nounits"], text=True, timeout=10)
            f = [x.strip() for x in out.strip().splitlines()[0].split(",")]
            reasons = [n for n, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7])
                       if v.lower().startswith("active")]
            return {"sm_mhz": float(f[0]), "sm_max_mhz": float(f[1]), "power_w_max": float(f[2]), "samples": 1,
                    "source": "nvidia-smi (after the timed region)", "reasons": reasons}
        except Exception:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi / NVML unavailable"]}


def host_cpu_model():
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.lower().startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def patch_pair(idx, h, w, scale, seed):
    """Patch `idx` of a synthetic evaluation list: the same bytes whatever the rank count (uint8 levels as stored)."""
    g = torch.Generator().manual_seed(seed * 100003 + idx)
    lr = torch.randint(0, 256, (1, h, w), generator=g, dtype=torch.uint8)
    hr = torch.randint(0, 256, (1, h * scale, w * scale), generator=g, dtype=torch.uint8)
    return lr, hr


def gpu_eager_run(kind, kw, sd_cpu, h, w, seed, batch, dev):
    """SURVEY 8(d) "the bar to beat on the same box": the reference algorithm as plain PyTorch eager modules on this
    GPU (the oracle port IS plain PyTorch: cuBLAS / cuDNN / ATen kernels), forward + the reference's metric formulas,
    in fp32 without TF32, fp32 with TF32 allowed, and autocast(bf16).  CUDA-event timed, 1 warm-up + 2 steps."""
    from oracle import sr_oracle as O
    out = {}
    if kind == "swinir":
        cfg = O.SwinIRCfg(**{k: kw[k] for k in ("upscale", "in_chans", "img_size", "window_size", "img_range",
                                                 "depths", "embed_dim", "num_heads", "mlp_ratio", "upsampler",
                                                 "resi_connection")})
        scale = cfg.upscale
        fwd = lambda sd, x: O.swinir_forward(sd, cfg, x)
    else:
        cfg = O.EDSRCfg(**{k: kw[k] for k in ("in_chans", "n_resblocks", "n_feats", "scale", "rgb_range")})
        scale = cfg.scale
        fwd = lambda sd, x: O.edsr_forward(sd, cfg, x)
    sd = {k: v.to(dev) for k, v in sd_cpu.items()}
    x, hr = synthetic_batch(batch, h, w, scale, seed)
    x, hr = x.to(dev), hr.to(dev)

    def step():
        y = fwd(sd, x)
        m = O.all_metrics(y.float(), hr, scale)
        r = O.roi_marginal_metrics(y.float(), hr, scale, ROI_THS)
        return m["psnr"].sum() + r["psnr"].sum()
    modes = (("fp32_no_tf32", False, None), ("fp32_tf32", True, None), ("autocast_bf16", True, torch.bfloat16))
    prev = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    try:
        for name, tf32, ac in modes:
            torch.backends.cuda.matmul.allow_tf32 = tf32
            torch.backends.cudnn.allow_tf32 = tf32
            with torch.no_grad(), torch.autocast("cuda", dtype=ac, enabled=ac is not None):
                step(); torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(2):
                    step()
                e1.record(); torch.cuda.synchronize()
            out[name] = batch * 2 / (e0.elapsed_time(e1) * 1e-3)
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = prev
    return {"unit": "patches/s", "batch": batch, "steps": 2, "warmup": 1, "values": out,
            "what": "oracle/sr_oracle.py (plain PyTorch restatement of the reference modules and metric formulas) "
                    "run eagerly on this GPU"}


def cpu_reference_run(kind, kw, sd, h, w, seed, steps, warmup, sample_b):
    """The reference's CPU path for this workload: oracle port (oracle/sr_oracle.py) forward +
    metrics on the host cores.  Used ONLY as the measured baseline."""
    from oracle import sr_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    if kind == "swinir":
        cfg = O.SwinIRCfg(**{k: kw[k] for k in ("upscale", "in_chans", "img_size", "window_size", "img_range",
                                                 "depths", "embed_dim", "num_heads", "mlp_ratio", "upsampler",
                                                 "resi_connection")})
        scale = cfg.upscale
        fwd = lambda x: O.swinir_forward(sd, cfg, x)
    else:
        cfg = O.EDSRCfg(**{k: kw[k] for k in ("in_chans", "n_resblocks", "n_feats", "scale", "rgb_range")})
        scale = cfg.scale
        fwd = lambda x: O.edsr_forward(sd, cfg, x)
    x, hr = synthetic_batch(sample_b, h, w, scale, seed)

    def step():
        y = fwd(x)
        m = O.all_metrics(y, hr, scale)
        r = O.roi_marginal_metrics(y, hr, scale, ROI_THS)
        return float(m["psnr"].sum() + r["psnr"].sum())
    with torch.no_grad():
        for _ in range(warmup):
            step()
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
        dt = time.perf_counter() - t0
    return sample_b * steps / dt, dt / steps, torch.get_num_threads()


def run_sweep(args, kind, kw, desc, net, B, h, w, scale, seed, swin_pad, rank, world, dev, metric, unit, W):
    """Strong scaling (BASELINE.json configs[3]: the 1471-patch test sweep): a fixed patch list, exact ragged shards
    (evaluator.shard_range), tail batches, one all-reduce of the 11 metric sums.  Inputs are device resident (uint8
    levels) when the timed region starts."""
    import torch.distributed as dist
    from sr_caco_2_b200 import evaluator as EV, _lib as L
    n = args.sweep
    lo, hi = EV.shard_range(n, rank, world)
    own = range(n) if rank == 0 else range(lo, hi)              # rank 0 also holds the whole list for the untimed cross-check
    pairs = {i: patch_pair(i, h, w, scale, seed) for i in own}
    lr_all = torch.stack([pairs[i][0] for i in own]).to(dev)
    hr_all = torch.stack([pairs[i][1] for i in own]).to(dev)
    base = 0 if rank == 0 else lo
    lr_sh, hr_sh = lr_all[lo - base:hi - base], hr_all[lo - base:hi - base]
    step_fn = EV.make_cuda_step(net, scale, swin_pad, roi_ths=ROI_THS, check=False)

    def sweep(lr, hr):
        acc = torch.zeros(11, dtype=torch.float64, device=dev)
        for s0 in range(0, lr.shape[0], B):
            vals = step_fn(lr[s0:s0 + B], hr[s0:s0 + B])
            acc[:10] += vals.sum(0)
            acc[10] += vals.shape[0]
        return acc

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    for _ in range(max(W, 3)):
        step_fn(lr_sh[:B], hr_sh[:B])
    barrier()
    sampler = ClockSampler(dev.index)
    if rank == 0:
        sampler.start()
    L.launch_count(reset=True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = max(args.steps, 1)
    barrier()
    ev0.record()
    for _ in range(K):
        acc = sweep(lr_sh, hr_sh)
        if world > 1:
            dist.all_reduce(acc)
    ev1.record()
    barrier()
    launches = L.launch_count()
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    if rank != 0:
        return
    means = (acc[:10] / acc[10]).cpu()
    single = sweep(lr_all, hr_all)                                 # untimed: the whole list on this GPU alone
    means1 = (single[:10] / single[10]).cpu()
    rel = float(((means - means1).abs() / means1.abs().clamp_min(1e-300)).max())
    assert int(acc[10].item()) == n and rel < 1e-9, f"sharded means differ from the single-GPU means: {rel}"
    shards = [EV.shard_range(n, r, world) for r in range(world)]
    line = {"metric": metric, "value": n * K / (ms * 1e-3), "unit": unit, "n_gpus": world, "steps": K, "warmup": max(W, 3),
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "bf16 linear/attention + fp16 conv operands, fp32 accumulate", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {desc}", "sweep_patches": n, "per_gpu_batch": B,
                       "shard_sizes": [b - a for a, b in shards], "tail_batch": [(b - a) % B for a, b in shards],
                       "step": "one step = the whole sweep + its all-reduce", "geometry": args.geometry,
                       "l2": "the sweep's inputs and activations (GBs) are far larger than the 126 MB L2"},
            "clocks": clocks, "gpu_launches": launches,
            "sharded_vs_single_gpu_means_max_rel_diff": rel,
            "means": {k: float(means[i]) for i, k in enumerate(EV.METRICS)}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg3")
    ap.add_argument("--engine", default="tcgen05", choices=["tcgen05", "mma_sync"])
    ap.add_argument("--batch", type=int, default=0, help="override the per-GPU batch")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--geometry", default="direct", choices=["direct", "evalpy"],
                    help="direct: net(x) on hxw; evalpy: caller-side +1 window padding (72x72)")
    ap.add_argument("--sweep", type=int, default=0, help="strong scaling: evaluate a fixed list of this many patches")
    ap.add_argument("--no-eager-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    from sr_caco_2_b200 import configs as CF
    kind, kw, B, h, w, desc = CF.WORKLOADS[args.workload]
    seed = 100 + int(args.workload[3:])
    if args.batch:
        B = args.batch
    scale = kw["upscale"] if kind == "swinir" else kw["scale"]
    W = max(args.warmup, 3)
    K = max(args.steps, 1)
    metric, unit = "SR patches/sec (forward + PSNR/SSIM scoring)", "patches/s"

    # ---------------------------------------------------------------------------------------
    if args.impl == "reference":
        if rank != 0:
            return
        sample_b = 4 if kind == "swinir" else 8
        steps = min(max(K, 6), 8)
        torch.manual_seed(seed)
        sd = {k: v.detach().clone() for k, v in CF.build(kind, kw).state_dict().items()}
        v, spt, thr = cpu_reference_run(kind, kw, sd, h, w, seed, steps, 1, sample_b)
        line = {"impl": "reference", "metric": metric, "value": v, "unit": unit, "n_gpus": args.gpus,
                "steps": steps, "warmup": 1, "ms_per_step": spt * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"{args.workload}: {desc}", "sample_batch": sample_b, "host_cpu": host_cpu_model()},
                "cpu_baseline": {"value": v, "unit": unit, "cores": thr, "kind": "port", "host_cpu": host_cpu_model(),
                                 "sample": f"{steps} steps x batch {sample_b} of the same workload through "
                                           "oracle/sr_oracle.py (fp32 PyTorch CPU restatement of the reference; "
                                           "/root/reference is not present on the GPU box)"},
                "e2e": {"value": v, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    # ---------------------------------------------------------------------------------------
    import torch.distributed as dist
    import sr_caco_2_b200 as S
    from sr_caco_2_b200 import utils_image as UI, evaluator as EV, _lib as L

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # NCCL prints its version banner on stdout at the first collective: keep stdout for the ONE JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    S.set_engine(args.engine)

    torch.manual_seed(seed)                       # random-init weights of the named architecture
    net = CF.build(kind, kw)
    sd_cpu = {k: v.detach().clone() for k, v in net.state_dict().items()}
    net = net.to(dev).eval()
    swin_pad = kind == "swinir" and args.geometry == "evalpy"

    if args.sweep > 0:
        run_sweep(args, kind, kw, desc, net, B, h, w, scale, seed, swin_pad, rank, world, dev, metric, unit, W)
        if world > 1:
            dist.destroy_process_group()
        return

    # synthetic inputs: NB distinct batches, rotated, resident in HBM for `value`
    NB = 3
    pairs = [synthetic_batch(B, h, w, scale, seed + 10 * rank + i) for i in range(NB)]
    lr_host = [p[0].pin_memory() for p in pairs]
    hr_host = [p[1].pin_memory() for p in pairs]
    lr_dev = [t.to(dev) for t in lr_host]
    hr_dev = [t.to(dev) for t in hr_host]
    acc = torch.zeros(11, dtype=torch.float64, device=dev)

    def device_step(i):
        e = EV.forward_with_padding(net, lr_dev[i % NB], scale, swin_pad)
        m = UI.compute_metrics(e, hr_dev[i % NB], scale, ROI_THS, check=False)
        cols = torch.stack([m[k] for k in EV.METRICS] + [m["roi_" + k] for k in EV.METRICS], 1)
        acc[:10] += cols.sum(0)
        acc[10] += B

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(W):
        device_step(i)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    acc.zero_()
    L.launch_count(reset=True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for i in range(K):
        device_step(i)
    if world > 1:
        dist.all_reduce(acc)          # the path's only exchange: 11 fp64 metric sums over NCCL
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = L.launch_count()
    clocks = sampler.stop() if rank == 0 else None
    # the same K steps again with a CUDA-event pair around every kernel-family call (srk_profile):
    # per-family device time for the roofline of the dominant kernel (kept out of `value`'s region
    # because the ~400 event records per step cost ~2 %)
    L.profile_read(reset=True)
    L.profile(True)
    for i in range(K):
        device_step(i)
    torch.cuda.synchronize()
    L.profile(False)
    prof_ms, prof_calls = L.profile_read(reset=True)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * B * K / (ms * 1e-3)

    # ---- e2e: host buffers in, host result out, every step ---------------------------------
    # The plugin call a user makes: evaluator.make_cuda_step on HOST tensors.  The loader's images are uint8
    # (SURVEY 8f-1): LR and HR travel as the stored uint8 levels, `uint2tensor` (/255) runs on the device and the
    # metrics kernel reads the target bytes directly.  The fp32 shipping of round 1 is timed beside it.
    res_host = [torch.empty(B, 10, dtype=torch.float64).pin_memory() for _ in range(2)]
    res_evt = [torch.cuda.Event() for _ in range(2)]
    step_fn = EV.make_cuda_step(net, scale, swin_pad, roi_ths=ROI_THS, check=False)
    lr8_host = [(t * 255).round().to(torch.uint8).pin_memory() for t in lr_host]
    hr8_host = [(t * 255).round().to(torch.uint8).pin_memory() for t in hr_host]

    def time_e2e(lrs, hrs, pipelined=True):
        """Every step: H2D of that step's inputs from pinned host memory, the plugin call, D2H of its (B, 10) score
        block and a host-side read of it.  pipelined: the host reads the scores of step i after it has enqueued step
        i + 1 (two pinned result buffers, one event each) -- what `evaluator.evaluate_patches` does for a whole sweep;
        otherwise the host waits for every step before it enqueues the next one."""
        acc = [0.0]
        def consume(slot):
            res_evt[slot].synchronize()
            acc[0] += res_host[slot][:, 0].sum().item()
        def run(n):
            pending = None
            for i in range(n):
                vals = step_fn(lrs[i % NB], hrs[i % NB])
                slot = i & 1
                res_host[slot].copy_(vals, non_blocking=True)
                res_evt[slot].record()
                if not pipelined:
                    consume(slot)
                    continue
                if pending is not None:
                    consume(pending)
                pending = slot
            if pending is not None:
                consume(pending)
        run(2)
        barrier()
        ev0.record()
        run(K)
        ev1.record()
        barrier()
        tt = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return world * B * K / (float(tt.item()) * 1e-3)

    e2e_value = time_e2e(lr8_host, hr8_host)
    e2e_sync = time_e2e(lr8_host, hr8_host, pipelined=False)
    e2e_fp32 = time_e2e(lr_host, hr_host)
    h2d = lr8_host[0].numel() + hr8_host[0].numel()
    d2h = res_host[0].numel() * 8

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline -------------------------------------------------------------------------
    # Headline: the tensor-core roofline of the WHOLE step (north_star / SURVEY 8d): FLOPs of one step / device time of
    # one step against the measured cuBLAS bf16 peak -- the burst figure when the SM clock held its maximum through the
    # timed region (median >= 1900 MHz), else the sustained one.  Both FLOP counts are given: the network as the
    # reference executes it layer by layer, and what this implementation executes (folded reconstruction tail).
    peaks = read_peaks()
    hn, wn = (h, w)
    if swin_pad:
        hn, wn = (h // 8 + 1) * 8, (w // 8 + 1) * 8
    fl = CF.swinir_flops if kind == "swinir" else CF.edsr_flops
    gflop_ref = fl(kw, hn, wn) / 1e9
    folded = args.engine == "tcgen05" and kw.get("upsampler", "pixelshuffle") == "pixelshuffle"
    gflop_exec = fl(kw, hn, wn, folded_tail=folded) / 1e9
    burst = bool(clocks and clocks.get("sm_mhz") and clocks["sm_mhz"] >= 1900.0 and peaks["tensor_burst"])
    peak = peaks["tensor_burst"] if burst else (peaks["tensor_sustained"] or peaks["tensor_burst"])
    per_gpu = value / world
    ach_exec, ach_ref = per_gpu * gflop_exec / 1e3, per_gpu * gflop_ref / 1e3
    roofline = {"bound": "tensor", "achieved": ach_exec, "peak": peak, "unit": "TFLOP/s", "frac": ach_exec / peak,
                "frac_executed_flops": ach_exec / peak, "frac_reference_flops": ach_ref / peak,
                "achieved_reference_flops": ach_ref, "traffic": None,
                "peak_source": peaks["source"] + (", cuBLAS bf16 burst (SM clock at maximum through the timed region)" if burst
                                                  else ", cuBLAS bf16 sustained"),
                "kernel": "whole step: every launch of the forward + metrics (CUDA events around the K timed steps)",
                "gflop_per_patch_executed": gflop_exec, "gflop_per_patch_reference": gflop_ref,
                "scope": "per GPU (rank 0)",
                "family_ms_per_step": {k: v / K for k, v in prof_ms.items() if prof_calls.get(k)},
                "family_launches_per_step": {k: v / K for k, v in prof_calls.items() if v}}
    tp = os.path.join(ROOT, "profiles", "r02_step_traffic.json")
    tj = json.load(open(tp)) if os.path.exists(tp) and args.workload == "cfg3" and args.geometry == "direct" and B == 32 else None
    if tj:
        roofline["traffic"] = tj["step"]["dram_bytes"]
        roofline["traffic_note"] = ("dram__bytes_read.sum + dram__bytes_write.sum summed over the %d launches of one step "
                                    "(ncu launch list profiles/r02_launches_cfg3_step.csv)" % tj["step"]["launches"])
    # The two kernels that make up a Swin block, each with its HBM view (algorithmic bytes per token on the UN-padded
    # embedding) and its tensor view, from CUDA-event pairs around every such launch (second pass of the same K steps).
    if kind == "swinir":
        C_, hid_ = kw["embed_dim"], int(kw["embed_dim"] * kw["mlp_ratio"])
        tok = B * hn * wn
        doms = []
        for fam, name, bytes_tok, flop_tok in (
                ("attn_block", "attn_block_tc5_kernel: LN1 rows -> qkv -> window attention -> proj + residual -> LN2",
                 2 * C_ + 4 * C_ + 4 * C_ + 2 * C_, 2.0 * (3 * C_ * C_ + 2 * 64 * C_ + C_ * C_)),
                ("mlp", "mlp_tc5_kernel: fc1 + GELU + fc2 + residual + next LN1",
                 2 * C_ + 4 * C_ + 4 * C_ + 2 * C_, 2.0 * (2 * C_ * hid_)),
                ("gemm_res_ln", "gemm_tc5_kernel<192, E_RES_LN>: proj / fc2 row GEMM + residual + LayerNorm (unfused blocks only)",
                 2 * C_ + 4 * C_ + 4 * C_ + 2 * C_, 2.0 * C_ * C_)):
            if prof_calls.get(fam, 0) == 0:
                continue
            t_l = prof_ms[fam] / prof_calls[fam] * 1e-3
            d = {"kernel": name, "launches_per_step": prof_calls[fam] / K, "avg_launch_us": t_l * 1e6,
                 "share_of_step": prof_ms[fam] / K / (ms / K),
                 "hbm": {"algorithmic_bytes_per_token": bytes_tok, "achieved": tok * bytes_tok / t_l / 1e9, "unit": "GB/s",
                         "peak": peaks["hbm"], "frac": tok * bytes_tok / t_l / 1e9 / peaks["hbm"]},
                 "tensor": {"algorithmic_flop_per_token": flop_tok, "achieved": tok * flop_tok / t_l / 1e12, "unit": "TFLOP/s",
                            "peak": peak, "frac": tok * flop_tok / t_l / 1e12 / peak}}
            if tj and fam in tj.get("by_family", {}):
                d["hbm"]["traffic"] = tj["by_family"][fam]["dram_bytes"] / max(tj["by_family"][fam]["launches"], 1)
            doms.append(d)
        doms.sort(key=lambda d: -d["share_of_step"])
        if doms:
            roofline["dominant_kernel"] = doms[0]
            roofline["other_block_kernels"] = doms[1:]

    line = {"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16 linear/attention + fp16 conv operands, fp32 accumulate", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {desc}", "net_input": f"{hn}x{wn}",
                       "geometry": args.geometry, "engine": args.engine, "per_gpu_batch": B,
                       "l2": "per-step working set (>2 GB of activations) is far larger than the 126 MB L2; "
                             f"{NB} distinct input batches are rotated"},
            "clocks": clocks, "gpu_launches": launches,
            "e2e": {"value": e2e_value, "unit": unit, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "shipping": "uint8 LR + uint8 HR (the loader's stored levels), /255 on the device",
                    "host_loop": "every step: H2D of its inputs from pinned memory, plugin call, D2H + host read of its (B, 10) scores; "
                                 "the host reads step i's scores after enqueueing step i + 1 (as evaluator.evaluate_patches does)",
                    "host_waits_every_step": {"value": e2e_sync},
                    "fp32_shipping": {"value": e2e_fp32, "h2d_bytes_per_step": 4 * h2d}},
            "roofline": roofline}
    if world == 1 and not args.no_eager_baseline:
        try:
            line["gpu_eager_baseline"] = gpu_eager_run(kind, kw, sd_cpu, h, w, seed, min(B, 8), dev)
        except Exception as exc:                                   # e.g. out of memory on a shared box: report, do not fail the bench
            line["gpu_eager_baseline"] = {"unavailable": repr(exc)[:200]}

    if not args.no_cpu_baseline and world >= 1:
        sample_b = 4 if kind == "swinir" else 8
        cpu_steps = 6 if kind == "swinir" else 8                 # ~10 s of host work for cfg3
        v, spt, thr = cpu_reference_run(kind, kw, sd_cpu, h, w, seed, cpu_steps, 1, sample_b)
        line["cpu_baseline"] = {"value": v, "unit": unit, "cores": thr, "kind": "port", "host_cpu": host_cpu_model(),
                                "sample": f"{cpu_steps} steps x batch {sample_b} (1 warm-up) of the same workload through "
                                          "oracle/sr_oracle.py on the host CPU, fp32"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
