"""Bottleneck knobs of the persistent conv kernel (needs a -DB200_TC_DEBUG build): times the two top-level layer shapes
with MMAs / epilogue stores / epilogue TMEM loads / TMA loads switched off one at a time."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_segmentation_project_b200 import _lib, functional as F

def timeit(fn, iters=7, flush=None):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None: flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]

dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
F.set_conv_persistent(1)
for Cin, Cout in ((16, 16), (32, 16), (16, 32)):
    x = torch.randn(2, 128, 128, 128, Cin, device=dev).bfloat16()
    w = torch.randn(Cout, Cin, 3, 3, 3, device=dev) * 0.05
    b = torch.randn(Cout, device=dev)
    wp = F.pack_conv3_weights(w, _lib.PACK_FPROP_TC, torch.bfloat16)
    for skip, name in ((0, "normal"), (12, "MMA only"), (1, "no MMA"), (1 | 12, "barriers only")):
        os.environ["B200_TC3_SKIP"] = str(skip)
        ms = timeit(lambda: F.conv3d_k3_raw(x, None, wp, b, Cout, 0, impl=2), flush=flush)
        print(f"{Cin:3d}->{Cout:3d} skip={skip:2d} {name:32s} {ms * 1e3:8.1f} us", flush=True)
    os.environ["B200_TC3_SKIP"] = "0"
    os.environ["B200_TC_DEBUG"] = "1"
    F.conv3d_k3_raw(x, None, wp, b, Cout, 0, impl=2)
    torch.cuda.synchronize()
    del os.environ["B200_TC_DEBUG"]
