"""CUDA-event timing of the HBM-bound elementwise kernels at the top-level activation size (2x128^3x16 bf16)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_segmentation_project_b200 import _lib, functional as F
from multimodal_segmentation_project_b200.functional import _ptr, _stream, _dt, check
L = _lib.load(); dev = torch.device("cuda")
N, S, C = 2, 128 ** 3, 16
x = torch.randn(N, 128, 128, 128, C, device=dev).bfloat16(); gy = torch.randn_like(x); y = torch.empty_like(x)
stats = torch.rand(4, C, device=dev) + 0.5
part = torch.empty(L.b200_bn_partials_bytes(C) // 4, device=dev)
sums = torch.zeros(2 * C, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def t(fn, nbytes, name):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(7):
        flush.zero_(); a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(e))
    ms = sorted(ts)[len(ts) // 2]
    print(f"{name:22s} {ms*1e3:7.1f} us  {nbytes/ms/1e6:7.0f} GB/s  ({nbytes/1e6:.0f} MB algorithmic)")
B = x.numel() * 2
t(lambda: check(L.b200_bn_stats(1, _ptr(x), N * S, C, _ptr(part), _stream())), B, "bn_stats")
t(lambda: check(L.b200_bn_act_fwd(1, _ptr(x), _ptr(y), _ptr(stats[0]), _ptr(stats[1]), _ptr(stats[2]), None, 1, N, S, C, _stream())), 2 * B, "bn_act_fwd")
t(lambda: check(L.b200_bn_act_bwd_reduce(1, _ptr(gy), _ptr(x), _ptr(stats[0]), _ptr(stats[1]), _ptr(stats[2]), _ptr(stats[3]), None, 1, N, S, C, _ptr(part), _stream())), 2 * B, "bn_act_bwd_reduce")
t(lambda: check(L.b200_bn_act_bwd_apply(1, _ptr(gy), _ptr(x), _ptr(y), _ptr(stats[0]), _ptr(stats[1]), _ptr(stats[2]), _ptr(stats[3]), None, 1, _ptr(sums), 1, N, S, C, _stream())), 3 * B, "bn_act_bwd_apply")
xp = torch.empty(N, 64, 64, 64, C, device=dev, dtype=torch.bfloat16)
t(lambda: check(L.b200_maxpool2_fwd(1, _ptr(x), _ptr(xp), N, 128, 128, 128, C, _stream())), B + B // 8, "maxpool2_fwd")
t(lambda: check(L.b200_maxpool2_bwd(1, _ptr(x), _ptr(xp), _ptr(y), N, 128, 128, 128, C, _stream())), 2 * B + B // 8, "maxpool2_bwd")
lg = torch.randn(N, 4, 128, 128, 128, device=dev); tg = torch.randint(0, 4, (N, 1, 128, 128, 128), device=dev)
t(lambda: F.confusion_counts(lg, tg), lg.numel() * 4 + tg.numel() * 8, "confusion")
lg.requires_grad_(True)
from multimodal_segmentation_project_b200.utils import metrics as M
t(lambda: M.combined_loss(lg, tg), lg.numel() * 4 + tg.numel() * 8, "seg_loss fwd(+finalize)")
# the tiny per-channel kernels (latency-bound): partial sums -> per-channel vectors
gam = torch.ones(C, device=dev); bet = torch.zeros(C, device=dev); rm_ = torch.zeros(C, device=dev); rv_ = torch.ones(C, device=dev)
nbt = torch.zeros(1, dtype=torch.int64, device=dev)
t(lambda: check(L.b200_bn_finalize(1, _ptr(x), _ptr(part), N * S, C, _ptr(gam), _ptr(bet), 1e-5, 0.1, 1, _ptr(rm_), _ptr(rv_), _ptr(nbt),
                                   _ptr(stats[0]), _ptr(stats[1]), _ptr(stats[2]), _ptr(stats[3]), _stream())), 592 * 2 * C * 4, "bn_finalize")
t(lambda: check(L.b200_bn_bwd_finalize(_ptr(part), N * S, C, _ptr(gam), _ptr(bet), _ptr(sums), _stream())), 296 * 2 * C * 4, "bn_bwd_finalize")
