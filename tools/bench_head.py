"""Fused head kernels alone at the benchmark shape (2 x 128^3, 16 channels, 4 classes): CUDA events, L2 flushed."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_segmentation_project_b200 import _lib, functional as F

dev = torch.device("cuda")
L = _lib.load()
N, S3, C, K = 2, 128 ** 3, 16, 4
x = torch.randn(N, 128, 128, 128, C, device=dev).bfloat16()
stats = torch.rand(4, C, device=dev) + 0.5
w = torch.randn(K, C, device=dev) * 0.1
b = torch.zeros(K, device=dev)
for lab_dtype in (torch.int64, torch.uint8):
    y = torch.randint(0, K, (N, 1, 128, 128, 128), device=dev).to(lab_dtype)
    logits = torch.empty(N, K, 128, 128, 128, device=dev)
    sums = torch.empty(4 + 4 * K, dtype=torch.float64, device=dev)
    conf = torch.empty(K, K, dtype=torch.int64, device=dev)
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    lb = 1 if lab_dtype == torch.uint8 else 8
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    def fwd():
        _lib.check(L.b200_head_fwd(P(x), P(stats[0]), P(stats[1]), P(stats[2]), P(w), P(b), 1, P(y), lb, N, S3, C, K, P(logits), P(sums), P(conf), st()))
    coef = torch.tensor([1.0 / (N * S3), 0, 0, -1e-7, -1e-7, -1e-7, 0, 1e-8, 1e-8, 1e-8], device=dev)
    go = torch.ones(1, device=dev)
    nb = L.b200_head_blocks(N, S3)
    gy = torch.empty_like(x); wpart = torch.empty(nb * 68, device=dev); bnpart = torch.empty(nb * 32, device=dev)
    dw = torch.empty(K, C, device=dev); db = torch.empty(K, device=dev)
    def bwd():
        _lib.check(L.b200_head_bwd(P(logits), P(y), lb, P(coef), P(go), P(x), P(stats[0]), P(stats[1]), P(stats[2]), P(stats[3]), P(w), N, S3, C, K,
                                   P(gy), P(wpart), P(bnpart), P(dw), P(db), st()))
    for name, fn, nbytes in (("head_fwd", fwd, N * S3 * (32 + lb + 16)), ("head_bwd", bwd, N * S3 * (16 + lb + 32 + 32))):
        fn(); torch.cuda.synchronize()
        ts = []
        for _ in range(7):
            flush.zero_()
            a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); e.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(e))
        ms = sorted(ts)[3]
        print(f"{name} labels {lab_dtype}: {ms * 1e3:7.1f} us  {nbytes / ms / 1e6:7.1f} GB/s", flush=True)
