"""Times the metric kernels alone at the BASELINE scoring size (32 x 512 x 512, border 8, 7 ROI thresholds):
the row-streaming hot-path kernel against the round-1 tile kernel, fp32 and uint8 targets.  CUDA events around
20 back-to-back calls (init + kernel + finalize), inputs 100 MB (L2 holds them: the number is the on-chip rate;
the DRAM rate is in the ncu launch list of the step, where the estimate comes straight from the network)."""
import json
import sys
import os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sr_caco_2_b200 import _lib as L, utils_image as UI

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
B, S = 32, 512
H = (torch.rand(B, 1, S, S, device=dev, generator=g) * 255).round() / 255
E = H + 0.05 * torch.randn(H.shape, device=dev, generator=g)
H8 = (H * 255).round().to(torch.uint8)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
lib = L.load()
out = {}
cases = [("stream_f32", 0, H), ("stream_u8", 0, H8), ("tile_f32", 1, H), ("tile_u8", 1, H8)]
cases += [(f"stream_f32_cps{c}", c << 8, H) for c in (2, 3, 5, 6, 8, 12)]
for name, tile, tgt in cases:
    lib.srk_metrics_use_tile_kernel(tile)
    for _ in range(3):
        UI.compute_metrics(E, tgt, 8, (4, 5, 6, 7, 8, 9, 10), check=False)
    ts = []
    for _ in range(20):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        UI.compute_metrics(E, tgt, 8, (4, 5, 6, 7, 8, 9, 10), check=False)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    out[name] = {"median_us": ts[len(ts) // 2], "min_us": ts[0]}
lib.srk_metrics_use_tile_kernel(0)
print(json.dumps(out))
