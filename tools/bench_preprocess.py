"""CUDA-event timing of the device-side input pipeline on a 512x512x256 volume (268 MB fp32, 537 MB int64 labels)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_segmentation_project_b200.utils import preprocess as P
dev = torch.device("cuda")
x = (torch.rand(256, 512, 512, device=dev) * 2000 - 800)
lab = torch.randint(0, 256, (256, 512, 512), device=dev)
def t(fn, nbytes, name):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(5):
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(e))
    ms = sorted(ts)[len(ts) // 2]
    print(f"{name:34s} {ms*1e3:8.1f} us  {nbytes/ms/1e6:7.0f} GB/s of algorithmic traffic ({nbytes/1e6:.0f} MB)")
n = x.numel()
t(lambda: P.preprocess_ct(x), 8 * n, "preprocess_ct (read + write fp32)")
t(lambda: P.preprocess_mri(x), 4 * n * (2 + 3 + 1) + 4 * n, "preprocess_mri (6 reads + 1 write)")
t(lambda: P.remap_labels(lab, "chaos_mri"), 16 * n, "remap_labels int64 -> int64")
t(lambda: P.remap_labels(lab, "chaos_mri", torch.uint8), 9 * n, "remap_labels int64 -> uint8")
