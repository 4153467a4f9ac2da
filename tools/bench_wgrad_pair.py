"""A/B timing of the 16-channel-slab weight-gradient kernels at the BASELINE config #2 shapes: plane-pair kernel
(wgrad_tc4.cu) against the M = 64 kernel (wgrad_tc2.cu).  CUDA events, L2 flushed between iterations."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_segmentation_project_b200 import functional as F

LAYERS = [("enc0.c1/dec3.c1", 128, 16, 0, 16), ("dec3.c0", 128, 16, 16, 16)]


def timeit(fn, flush, iters=7):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


def main():
    dev = torch.device("cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    dsegs = [int(a) for a in sys.argv[1:]] or [0]
    for name, S, c0, c1, Cout in LAYERS:
        N = 2
        x0 = torch.randn(N, S, S, S, c0, device=dev).bfloat16()
        x1 = torch.randn(N, S, S, S, c1, device=dev).bfloat16() if c1 else None
        dy = torch.randn(N, S, S, S, Cout, device=dev).bfloat16()
        flops = 2.0 * N * S ** 3 * 27 * (c0 + c1) * Cout
        row = {"layer": name, "S": S, "Cin": c0 + c1, "Cout": Cout}
        F.set_wgrad_pair(False)
        ref, _ = F.conv3d_wgrad_raw(x0, x1, dy, want_bias=False)
        ms = timeit(lambda: F.conv3d_wgrad_raw(x0, x1, dy, want_bias=False), flush)
        row["tc2_ms"] = round(ms, 4); row["tc2_tflops"] = round(flops / ms / 1e9, 1)
        for ds in dsegs:
            F.set_wgrad_pair(True, ds)
            out, _ = F.conv3d_wgrad_raw(x0, x1, dy, want_bias=False)
            ms = timeit(lambda: F.conv3d_wgrad_raw(x0, x1, dy, want_bias=False), flush)
            row[f"pair_d{ds}_ms"] = round(ms, 4); row[f"pair_d{ds}_tflops"] = round(flops / ms / 1e9, 1)
            row[f"pair_d{ds}_rel"] = float(((out - ref).norm() / ref.norm()).item())
        F.set_wgrad_pair(True, 0)
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
