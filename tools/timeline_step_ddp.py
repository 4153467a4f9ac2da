"""Rank-0 timeline of one graph-replayed 2-GPU train step: where the NCCL kernels sit and what the main stream does meanwhile.
torchrun --nproc-per-node 2 tools/timeline_step_ddp.py"""
import sys, os, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from torch.profiler import profile, ProfilerActivity
from multimodal_segmentation_project_b200.dp import DataParallelTrainer
from multimodal_segmentation_project_b200.models.unet import UNet3D
from multimodal_segmentation_project_b200.synthetic import structured_volume
from multimodal_segmentation_project_b200.utils import metrics as M

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
torch.manual_seed(0)
model = UNet3D(1, 4, dropout_rate=0.0).to(dev).train()
tr = DataParallelTrainer(model, M.combined_loss, lr=1e-3, autocast_dtype=torch.bfloat16, metrics_fn="confusion")
x, y = structured_volume(2, 128, seed=1234 + rank)
x, y = x.to(dev).bfloat16(), y.to(dev).to(torch.uint8)
tr.capture(x, y, warmup=3)
for _ in range(5):
    tr.replay()
torch.cuda.synchronize(); dist.barrier()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        tr.replay()
    torch.cuda.synchronize()
if rank == 0:
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    per = len(evs) // 3
    last = evs[-per:]
    t0 = last[0].time_range.start
    rows = [(e.time_range.start - t0, e.time_range.end - e.time_range.start, e.name[:70]) for e in last]
    span = max(s + d for s, d, _ in rows)
    print(f"{len(rows)} kernels, span {span:.1f} us")
    for s, d, n in rows:
        if "nccl" in n.lower() or "AllReduce" in n:
            print(f"NCCL  start {s:8.1f} us  dur {d:7.1f} us  {n}")
    # what ran while NCCL kernels were active, and how long it took compared with ... just list neighbours
    nccl = [(s, s + d) for s, d, n in rows if "nccl" in n.lower()]
    for a, b in nccl:
        print(f"--- kernels overlapping NCCL [{a:.0f}, {b:.0f}]:")
        for s, d, n in rows:
            if s < b and s + d > a and "nccl" not in n.lower():
                print(f"      {s:8.1f} {d:7.1f} {n.replace('(anonymous namespace)::', '')[:60]}")
    print("last 12 kernels of the step:")
    for s, d, n in rows[-12:]:
        print(f"      {s:8.1f} {d:7.1f} {n.replace('(anonymous namespace)::', '')[:60]}")
dist.barrier()
torch.cuda.synchronize()
os._exit(0)
