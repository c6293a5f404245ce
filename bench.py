#!/usr/bin/env python
"""Benchmark of the SR-CACO-2 evaluation hot path on B200 (contract: see the task statement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--workload cfg3|cfg1|cfg2|cfg4|cfg5] [--engine tcgen05|mma_sync]

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
        steps = min(K, 3)
        torch.manual_seed(seed)
        sd = {k: v.detach().clone() for k, v in CF.build(kind, kw).state_dict().items()}
        v, spt, thr = cpu_reference_run(kind, kw, sd, h, w, seed, steps, 1, sample_b)
        line = {"impl": "reference", "metric": metric, "value": v, "unit": unit, "n_gpus": args.gpus,
                "steps": steps, "warmup": 1, "ms_per_step": spt * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"{args.workload}: {desc}", "sample_batch": sample_b},
                "cpu_baseline": {"value": v, "unit": unit, "cores": thr, "kind": "port",
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
    res_host = torch.empty(B, 10, dtype=torch.float64).pin_memory()
    step_fn = EV.make_cuda_step(net, scale, swin_pad, roi_ths=ROI_THS, check=False)

    def e2e_step(i):
        vals = step_fn(lr_host[i % NB], hr_host[i % NB])
        res_host.copy_(vals, non_blocking=True)
        torch.cuda.current_stream().synchronize()      # the caller reads the scores every step
        return res_host[:, 0].sum().item()

    for i in range(2):
        e2e_step(i)
    barrier()
    ev0.record()
    for i in range(K):
        e2e_step(i)
    ev1.record()
    barrier()
    t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * B * K / (float(t.item()) * 1e-3)
    h2d = lr_host[0].numel() * 4 + hr_host[0].numel() * 4
    d2h = res_host.numel() * 8

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (tensor-core GEMM family) ---------------------------
    peaks = read_peaks()
    hn, wn = (h, w)
    if swin_pad:
        hn, wn = (h // 8 + 1) * 8, (w // 8 + 1) * 8
    fl = CF.swinir_flops if kind == "swinir" else CF.edsr_flops
    gflop_patch = fl(kw, hn, wn) / 1e9                       # the network as the reference executes it
    folded = (args.engine == "tcgen05" and os.environ.get("SRK_FOLD_TAIL", "1") != "0"
              and kw.get("upsampler", "pixelshuffle") == "pixelshuffle")
    gflop_exec = fl(kw, hn, wn, folded_tail=folded) / 1e9    # what this implementation executes
    peak = peaks["tensor_sustained"] or peaks["tensor_burst"]
    step_tflops = value / world * gflop_exec / 1e3
    roofline = {"bound": "tensor", "achieved": None, "peak": peak, "unit": "TFLOP/s", "frac": None,
                "traffic": None, "peak_source": peaks["source"] + ", sustained bf16",
                "scope": "per GPU (rank 0)", "whole_step_achieved": step_tflops, "whole_step_frac": step_tflops / peak,
                "gflop_per_patch": gflop_exec, "reference_gflop_per_patch": gflop_patch,
                "flops_note": "executed FLOPs; the linear reconstruction tail is folded into one 5x5 conv "
                              "(srk_tail_fold), the reference's layer-by-layer count is reference_gflop_per_patch"
                              if folded else "executed == reference layer-by-layer count"}
    gemm_ms = prof_ms.get("gemm", 0.0) + prof_ms.get("gemm_res_ln", 0.0)
    gemm_calls = prof_calls.get("gemm", 0) + prof_calls.get("gemm_res_ln", 0)
    family = None
    if gemm_ms > 0:
        fused_attn = kind == "swinir" and prof_ms.get("attention", 0.0) == 0.0   # attention ran inside the qkv GEMM kernel
        gf = fl(kw, hn, wn, gemm_only=True, attention_in_gemm=fused_attn, folded_tail=folded)
        ach = gf * B * K / (gemm_ms * 1e-3) / 1e12
        family = {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                  "kernel": "srk tcgen05 GEMM kernel family incl. the fused qkv+window-attention kernel (every launch of "
                            "K steps, CUDA events on the launch stream)",
                  "algorithmic_gflop_per_patch_in_gemm": gf / 1e9, "launches_per_step": gemm_calls / K,
                  "ms_per_step": gemm_ms / K}
        roofline.update(achieved=ach, frac=ach / peak, kernel=family["kernel"],
                        algorithmic_gflop_per_patch_in_gemm=gf / 1e9, launches_per_step=gemm_calls / K,
                        ms_per_step=gemm_ms / K, family_ms_per_step={k: v / K for k, v in prof_ms.items()})
    else:
        roofline.update(achieved=step_tflops, frac=step_tflops / peak,
                        kernel="whole step (per-kernel event timing unavailable)")
    tp = os.path.join(ROOT, "profiles", "r01_gemm_traffic.json")
    tj = json.load(open(tp)) if os.path.exists(tp) and args.workload == "cfg3" and args.geometry == "direct" and B == 32 else None
    if tj:
        roofline["traffic"] = tj["gemm_family"]["dram_bytes"]
        roofline["traffic_note"] = ("dram__bytes_read.sum + dram__bytes_write.sum summed over the %d GEMM launches of one "
                                    "step (ncu launch list profiles/r01_launches_cfg3_step.csv); same per-step scope "
                                    "as `achieved`" % tj["gemm_family"]["launches"])
    # The dominant single kernel of a SwinIR step is the row GEMM with residual + fused LayerNorm epilogue
    # (proj and fc2 of every block: gemm_tc5_kernel<192, E_RES_LN>): it is HBM bound, so ITS roofline is the
    # headline one and the tensor-pipe view of the whole GEMM family moves to `gemm_family`.
    if kind == "swinir" and prof_ms.get("gemm_res_ln", 0.0) > 0 and peaks.get("hbm"):
        Cp_, hid_p_ = (kw["embed_dim"] + 63) // 64 * 64, (int(kw["embed_dim"] * kw["mlp_ratio"]) + 63) // 64 * 64
        nh_max = max(kw["num_heads"])
        ao_p_ = ((-(-(kw["embed_dim"] // nh_max) // 16) * 16) * nh_max + 63) // 64 * 64
        tok = B * hn * wn
        # per token: A (16-bit) + residual in (fp32) + residual out (fp32) + LayerNorm output (16-bit)
        b_proj, b_fc2 = 2 * ao_p_ + 10 * Cp_, 2 * hid_p_ + 10 * Cp_
        n_l = prof_calls["gemm_res_ln"] / K
        bytes_launch = tok * (b_proj + b_fc2) / 2.0
        t_launch = prof_ms["gemm_res_ln"] / prof_calls["gemm_res_ln"] * 1e-3
        ach_gbs = bytes_launch / t_launch / 1e9
        hbm = {"bound": "hbm", "achieved": ach_gbs, "peak": peaks["hbm"], "unit": "GB/s", "frac": ach_gbs / peaks["hbm"],
               "traffic": None, "peak_source": peaks["source"] + ", device copy",
               "kernel": "gemm_tc5_kernel<192, E_RES_LN>: proj / fc2 row GEMM + bias + fp32 residual + fused LayerNorm, "
                         "%.0f launches per step, %.1f %% of the step" % (n_l, 100.0 * prof_ms["gemm_res_ln"] / K / (ms / K)),
               "algorithmic_bytes_per_launch": bytes_launch,
               "algorithmic_bytes_per_token": {"proj": b_proj, "fc2": b_fc2},
               "avg_launch_us": t_launch * 1e6, "launches_per_step": n_l,
               "share_of_step": prof_ms["gemm_res_ln"] / K / (ms / K), "scope": "per GPU (rank 0)"}
        if tj and "by_kernel" in tj:
            for name, kinfo in tj["by_kernel"].items():
                if name.startswith("gemm_tc5_kernel<192, 2, 0, 0>"):
                    hbm["traffic"] = kinfo["dram_bytes"] / kinfo["launches"]
                    hbm["traffic_note"] = ("dram__bytes_read.sum + dram__bytes_write.sum per launch, average over the %d "
                                           "launches of this kernel in profiles/r01_launches_cfg3_step.csv; below the "
                                           "algorithmic bytes because part of the written rows is still dirty in the "
                                           "126 MB L2 when the next kernel reads them" % kinfo["launches"])
        hbm["gemm_family"] = family
        hbm["whole_step"] = {k: roofline[k] for k in ("whole_step_achieved", "whole_step_frac", "gflop_per_patch",
                                                       "reference_gflop_per_patch", "flops_note")}
        hbm["whole_step"]["family_ms_per_step"] = {k: v / K for k, v in prof_ms.items()}
        roofline = hbm

    line = {"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16 linear/attention + fp16 conv operands, fp32 accumulate", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {desc}", "net_input": f"{hn}x{wn}",
                       "geometry": args.geometry, "engine": args.engine, "per_gpu_batch": B,
                       "l2": "per-step working set (>2 GB of activations) is far larger than the 126 MB L2; "
                             f"{NB} distinct input batches are rotated"},
            "clocks": clocks, "gpu_launches": launches,
            "e2e": {"value": e2e_value, "unit": unit, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "roofline": roofline}

    if not args.no_cpu_baseline and world >= 1:
        sample_b = 4 if kind == "swinir" else 8
        cpu_steps = 6 if kind == "swinir" else 8                 # ~10 s of host work for cfg3
        v, spt, thr = cpu_reference_run(kind, kw, sd_cpu, h, w, seed, cpu_steps, 1, sample_b)
        line["cpu_baseline"] = {"value": v, "unit": unit, "cores": thr, "kind": "port",
                                "sample": f"{cpu_steps} steps x batch {sample_b} (1 warm-up) of the same workload through "
                                          "oracle/sr_oracle.py on the host CPU, fp32"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
