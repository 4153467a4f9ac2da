"""Runs the top-level conv layer (2x128^3, 16->16) fprop + dgrad-shaped + wgrad a few times: target for ncu."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_segmentation_project_b200 import _lib, functional as F
dev = torch.device("cuda")
S, N = 128, 2
cin = int(sys.argv[1]) if len(sys.argv) > 1 else 16
cout = int(sys.argv[2]) if len(sys.argv) > 2 else 16
x = torch.randn(N, S, S, S, cin, device=dev).bfloat16()
dy = torch.randn(N, S, S, S, cout, device=dev).bfloat16()
w = torch.randn(cout, cin, 3, 3, 3, device=dev) * 0.05
b = torch.randn(cout, device=dev)
wp = F.pack_conv3_weights(w, _lib.PACK_FPROP_TC, torch.bfloat16)
for _ in range(3):
    F.conv3d_k3_raw(x, None, wp, b, cout, 0, impl=2)
    F.conv3d_wgrad_raw(x, None, dy, want_bias=False)
torch.cuda.synchronize()
print("ok")
