"""ncu target: the voxel-pair weight gradient (wgrad_tc4.cu) on the encoder.0.c1 / decoder.3.c0 shapes, a few launches each."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_segmentation_project_b200 import functional as F

dev = torch.device("cuda")
N, S = 2, 128
for c1 in (0, 16):
    x0 = torch.randn(N, S, S, S, 16, device=dev).bfloat16()
    x1 = torch.randn(N, S, S, S, 16, device=dev).bfloat16() if c1 else None
    dy = torch.randn(N, S, S, S, 16, device=dev).bfloat16()
    for _ in range(3):
        F.conv3d_wgrad_raw(x0, x1, dy, want_bias=False)
    torch.cuda.synchronize()
