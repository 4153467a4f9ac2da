"""The kernel bench.py reports in `roofline`: conv3d_tc2_kernel on decoder.3.c0's fprop shape (2x128^3, (16+16)->16), plus the
matching wgrad_tc2_kernel. Target for `ncu --set full`."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_segmentation_project_b200 import _lib, functional as F
dev = torch.device("cuda"); S, N = 128, 2
x0 = torch.randn(N, S, S, S, 16, device=dev).bfloat16(); x1 = torch.randn(N, S, S, S, 16, device=dev).bfloat16()
dy = torch.randn(N, S, S, S, 16, device=dev).bfloat16()
w = torch.randn(16, 32, 3, 3, 3, device=dev) * 0.05; b = torch.zeros(16, device=dev)
wp = F.pack_conv3_weights(w, _lib.PACK_FPROP_TC, torch.bfloat16)
for _ in range(3):
    F.conv3d_k3_raw(x0, x1, wp, b, 16, 0, impl=2)
    F.conv3d_wgrad_raw(x0, x1, dy, want_bias=False)
torch.cuda.synchronize(); print("ok")
# wide-row weight gradient (wgrad_tc3_kernel): decoder.2.c0's shape, 2x64^3, (32+32) -> 32
S2 = 64
q0 = torch.randn(N, S2, S2, S2, 32, device=dev).bfloat16(); q1 = torch.randn(N, S2, S2, S2, 32, device=dev).bfloat16()
dy2 = torch.randn(N, S2, S2, S2, 32, device=dev).bfloat16()
for _ in range(3):
    F.conv3d_wgrad_raw(q0, q1, dy2, want_bias=False)
torch.cuda.synchronize(); print("ok wide")
