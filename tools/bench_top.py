"""Top-level (2 x 128^3) convolution layers: row-streaming kernel vs the persistent window kernel, CUDA events, L2 flushed."""
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
for c0, c1, co0, co1 in ((16, 0, 16, 0), (16, 16, 16, 0), (16, 0, 16, 16)):
    x0 = torch.randn(2, 128, 128, 128, c0, device=dev).bfloat16()
    x1 = torch.randn(2, 128, 128, 128, c1, device=dev).bfloat16() if c1 else None
    w = torch.randn(co0 + co1, c0 + c1, 3, 3, 3, device=dev) * 0.05
    b = torch.randn(co0 + co1, device=dev)
    wp = F.pack_conv3_weights(w, _lib.PACK_FPROP_TC, torch.bfloat16)
    gf = 2.0 * 2 * 128 ** 3 * 27 * (c0 + c1) * (co0 + co1) / 1e9
    for rs in (1, 0):
        F.set_conv_rowstream(bool(rs))
        ms = timeit(lambda: F.conv3d_k3_raw(x0, x1, wp, b, co0, co1, impl=2), flush=flush)
        print(f"({c0}+{c1})->({co0}+{co1}) rowstream={rs}: {ms * 1e3:7.1f} us  {gf / ms:7.1f} TFLOP/s", flush=True)
    if os.environ.get("B200_KNOBS"):
        F.set_conv_rowstream(True)
        for skip, name in ((12, "MMA only"), (1, "no MMA"), (4, "no epilogue"), (8, "no TMA"), (13, "barriers only")):
            os.environ["B200_TC4_SKIP"] = str(skip)
            ms = timeit(lambda: F.conv3d_k3_raw(x0, x1, wp, b, co0, co1, impl=2), flush=flush)
            print(f"    skip={skip:2d} {name:16s} {ms * 1e3:7.1f} us", flush=True)
        os.environ["B200_TC4_SKIP"] = "0"
    F.set_conv_rowstream(True)
