"""Targets for `ncu --set full` at the benchmark shapes (2 x 128^3): the kernel bench.py reports in `roofline` — conv3d_tc4_kernel on
decoder.3.c0's fprop, (16+16) -> 16 — its data gradient, the matching wgrad_tc2_kernel, the persistent window kernel on the 64^3
level (conv3d_tc3_kernel, (32+32) -> 32) and the fused head."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_segmentation_project_b200 import _lib, functional as F
dev = torch.device("cuda"); S, N = 128, 2
x0 = torch.randn(N, S, S, S, 16, device=dev).bfloat16(); x1 = torch.randn(N, S, S, S, 16, device=dev).bfloat16()
dy = torch.randn(N, S, S, S, 16, device=dev).bfloat16()
w = torch.randn(16, 32, 3, 3, 3, device=dev) * 0.05; b = torch.zeros(16, device=dev)
wp = F.pack_conv3_weights(w, _lib.PACK_FPROP_TC, torch.bfloat16)
wpd = F.pack_conv3_weights(w, _lib.PACK_DGRAD_TC, torch.bfloat16)
for _ in range(3):
    F.conv3d_k3_raw(x0, x1, wp, b, 16, 0, impl=2)            # conv3d_tc4_kernel<0, 0>
    F.conv3d_k3_raw(dy, None, wpd, None, 16, 16, impl=2)     # conv3d_tc4_kernel<0, 0>, n_tile 32
    F.conv3d_wgrad_raw(x0, x1, dy, want_bias=False)          # wgrad_tc2_kernel
torch.cuda.synchronize(); print("ok top level")
S2 = 64
q0 = torch.randn(N, S2, S2, S2, 32, device=dev).bfloat16(); q1 = torch.randn(N, S2, S2, S2, 32, device=dev).bfloat16()
dy2 = torch.randn(N, S2, S2, S2, 32, device=dev).bfloat16()
w2 = torch.randn(32, 64, 3, 3, 3, device=dev) * 0.05; b2 = torch.zeros(32, device=dev)
wp2 = F.pack_conv3_weights(w2, _lib.PACK_FPROP_TC, torch.bfloat16)
for _ in range(3):
    F.conv3d_k3_raw(q0, q1, wp2, b2, 32, 0, impl=2)          # conv3d_tc3_kernel<0>
    F.conv3d_wgrad_raw(q0, q1, dy2, want_bias=False)         # wgrad_tc3_kernel
torch.cuda.synchronize(); print("ok 64^3 level")
