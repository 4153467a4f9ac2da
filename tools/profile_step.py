"""Per-kernel device time of the eager train step (2x128^3 bf16) from torch.profiler (CUPTI activity records:
warm caches, real back-to-back execution — unlike the ncu launch list, which replays every kernel cold and serialised)."""
import sys, os, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from multimodal_segmentation_project_b200 import functional as F
from multimodal_segmentation_project_b200.dp import DataParallelTrainer
from multimodal_segmentation_project_b200.models.unet import UNet3D
from multimodal_segmentation_project_b200.synthetic import structured_volume
from multimodal_segmentation_project_b200.utils import metrics as M

dev = torch.device("cuda")
torch.manual_seed(0)
model = UNet3D(1, 4, dropout_rate=0.0).to(dev).train()
tr = DataParallelTrainer(model, M.combined_loss, lr=1e-3, autocast_dtype=torch.bfloat16, metrics_fn="confusion")
x, y = structured_volume(2, 128, seed=1234)
x, y = x.to(dev).bfloat16(), y.to(dev).to(torch.uint8)
for _ in range(3):
    tr.step(x, y)
torch.cuda.synchronize()
steps = 3
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(steps):
        tr.step(x, y)
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
t0, t1 = None, None
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        agg[e.name][0] += 1
        agg[e.name][1] += e.device_time if hasattr(e, "device_time") else e.cuda_time
tot = sum(v for _, v in agg.values())
print(f"{sum(c for c, _ in agg.values()) // steps} launches/step, sum of kernel time {tot / steps / 1e3:.3f} ms/step")
for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{v / steps / 1e3:8.3f} ms {100 * v / tot:5.1f}% x{c // steps:3d}  avg {v / c:7.1f} us  {n[:100]}")
# per-launch durations of the convolution kernels of the LAST profiled step, in launch order
# (forward: enc0.c1 .. bott .. dec3.c1, then the data gradients in reverse order)
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and ("conv3d_tc" in e.name or "splitk" in e.name)]
evs.sort(key=lambda e: e.time_range.start)
per = len(evs) // steps
last = evs[-per:]
print("conv launches of one step (us):")
print(" ".join(f"{'T3' if 'tc3' in e.name else ('sk' if 'splitk' in e.name else 'T2')}:{(e.device_time if hasattr(e, 'device_time') else e.cuda_time):.0f}" for e in last))
