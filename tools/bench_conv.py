"""Per-layer timing of the 3x3x3 convolution kernels at BASELINE config #2 shapes (B=2, 128^3):
tcgen05 path vs CUDA-core path vs torch/cuDNN bf16 channels_last_3d (context only). CUDA events,
L2 flushed between iterations."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_segmentation_project_b200 import _lib, functional as F

LAYERS = [  # name, S, Cin, Cout
    ("enc0.c1", 128, 16, 16), ("enc1.c0", 64, 16, 32), ("enc1.c1", 64, 32, 32), ("enc2.c0", 32, 32, 64), ("enc2.c1", 32, 64, 64),
    ("enc3.c0", 16, 64, 128), ("enc3.c1", 16, 128, 128), ("bott.c0", 8, 128, 256), ("bott.c1", 8, 256, 256),
    ("dec0.c0", 16, 256, 128), ("dec1.c0", 32, 128, 64), ("dec2.c0", 64, 64, 32), ("dec3.c0", 128, 32, 16),
]

def timeit(fn, iters=5, flush=None):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None: flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]

def main():
    dev = torch.device("cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    impls = [int(a) for a in sys.argv[1:] if a.isdigit()] or [2]
    rows = []
    for name, S, Cin, Cout in LAYERS:
        N = 2
        x = torch.randn(N, S, S, S, Cin, device=dev).bfloat16()
        w = torch.randn(Cout, Cin, 3, 3, 3, device=dev) * 0.05
        b = torch.randn(Cout, device=dev)
        flops = 2.0 * N * S ** 3 * 27 * Cin * Cout
        row = {"layer": name, "S": S, "Cin": Cin, "Cout": Cout, "gflop": flops / 1e9}
        for impl in impls:
            wp = F.pack_conv3_weights(w, _lib.PACK_FPROP_TC if impl == 2 else _lib.PACK_FPROP, torch.bfloat16)
            ms = timeit(lambda: F.conv3d_k3_raw(x, None, wp, b, Cout, 0, impl=impl), flush=flush)
            row[f"impl{impl}_ms"] = ms; row[f"impl{impl}_tflops"] = flops / ms / 1e9
        dy = torch.randn(N, S, S, S, Cout, device=dev).bfloat16()
        for wi in (2, 1):
            if wi == 1 and S >= 64 and "--slow" not in sys.argv:
                continue
            F.set_wgrad_impl(wi)
            try:
                ms = timeit(lambda: F.conv3d_wgrad_raw(x, None, dy, want_bias=False), flush=flush)
                row[f"wgrad{wi}_ms"] = ms; row[f"wgrad{wi}_tflops"] = flops / ms / 1e9
            except Exception as e:
                row[f"wgrad{wi}_err"] = str(e)[:80]
            F.set_wgrad_impl(0)
        xt = x.permute(0, 4, 1, 2, 3)  # NCDHW view of channels-last memory
        wt = w.bfloat16().contiguous(memory_format=torch.channels_last_3d)
        ms = timeit(lambda: torch.nn.functional.conv3d(xt, wt, b.bfloat16(), padding=1), flush=flush)
        row["cudnn_ms"] = ms; row["cudnn_tflops"] = flops / ms / 1e9
        rows.append(row)
        print(json.dumps(row), flush=True)

if __name__ == "__main__":
    main()
