"""One warm-up step + one measured step of the cfg3 hot path (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sr_caco_2_b200 as S
from sr_caco_2_b200 import configs as CF, utils_image as UI
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
kind, kw, B, h, w, desc = CF.WORKLOADS[wl]
B = int(os.environ.get("SRK_B", B))
torch.manual_seed(0)
net = CF.build(kind, kw).cuda().eval()
from sr_caco_2_b200 import _lib as L
net.options = L.OPT_NO_GRAPH          # plain launches: one profiler record per kernel
scale = kw.get("upscale", kw.get("scale"))
x = torch.rand(B, 1, h, w, device="cuda")
hr = (torch.rand(B, 1, h * scale, w * scale, device="cuda") * 255).round() / 255
for _ in range(int(os.environ.get("SRK_STEPS", 2))):
    y = net(x)
    m = UI.compute_metrics(y, hr, scale, (4, 5, 6, 7, 8, 9, 10), check=False)
torch.cuda.synchronize()
print("ok", float(m["psnr"].mean()))
