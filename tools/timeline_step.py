"""Timeline of ONE graph-replayed train step (2 x 128^3 bf16): every kernel with start / duration / stream from CUPTI activity
records, then: span of the step, time with no kernel running anywhere, busy time per stream, the longest idle gaps."""
import sys, os, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
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
tr.capture(x, y, warmup=3)
for _ in range(5):
    tr.replay()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        tr.replay()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
# split into the three replays by the largest gaps is fragile: use kernel count instead
per = len(evs) // 3
last = evs[-per:]
t0 = last[0].time_range.start
rows = []
for e in last:
    s = e.time_range.start - t0
    d = e.time_range.end - e.time_range.start
    rows.append((s, d, getattr(e, "stream", getattr(e, "device_resource_id", -1)), e.name.replace("(anonymous namespace)::", "").replace("void ", "").split("(")[0][:60]))
span = max(s + d for s, d, _, _ in rows)
# union of busy intervals
iv = sorted((s, s + d) for s, d, _, _ in rows)
busy, cur_s, cur_e, gaps = 0.0, iv[0][0], iv[0][1], []
for s, e in iv[1:]:
    if s > cur_e:
        busy += cur_e - cur_s
        gaps.append((s - cur_e, cur_e))
        cur_s, cur_e = s, e
    else:
        cur_e = max(cur_e, e)
busy += cur_e - cur_s
print(f"{len(rows)} kernels, span {span:.1f} us, some kernel running {busy:.1f} us, nothing running {span - busy:.1f} us over {len(gaps)} gaps")
streams = collections.defaultdict(float)
for s, d, st, _ in rows:
    streams[st] += d
for st, v in sorted(streams.items(), key=lambda kv: -kv[1]):
    print(f"  stream {st}: kernel time {v:.1f} us")
print("largest idle gaps (us, at):", [(round(g, 1), round(a, 1)) for g, a in sorted(gaps, reverse=True)[:12]])
print("gap histogram:", collections.Counter(min(int(g), 10) for g, _ in gaps))
byname = collections.defaultdict(lambda: [0.0, 0])
for s, d, st, n in rows:
    byname[(st, n)][0] += d
    byname[(st, n)][1] += 1
for (st, n), (v, c) in sorted(byname.items(), key=lambda kv: -kv[1][0]):
    print(f"  {v:8.1f} us x{c:3d} s{st} {n}")
if "--all" in sys.argv:
    for s, d, st, n in rows:
        print(f"{s:9.1f} {d:8.1f} s{st} {n}")
