"""Deep-layer (16^3 / 8^3) convolution timings, L2-warm, back-to-back launches (what the step sees)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_segmentation_project_b200 import _lib, functional as F
dev = torch.device("cuda"); N = 2
for name, S, cin, cout in [("enc3.c0", 16, 64, 128), ("enc3.c1", 16, 128, 128), ("bott.c0", 8, 128, 256), ("bott.c1", 8, 256, 256), ("dec0.c0", 16, 256, 128), ("enc2.c1", 32, 64, 64)]:
    x = torch.randn(N, S, S, S, cin, device=dev).bfloat16()
    w = torch.randn(cout, cin, 3, 3, 3, device=dev) * 0.05
    b = torch.randn(cout, device=dev)
    wp = F.pack_conv3_weights(w, _lib.PACK_FPROP_TC, torch.bfloat16)
    for _ in range(3): F.conv3d_k3_raw(x, None, wp, b, cout, 0, impl=2)
    torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): F.conv3d_k3_raw(x, None, wp, b, cout, 0, impl=2)
    e.record(); torch.cuda.synchronize()
    us = a.elapsed_time(e) / 20 * 1e3
    fl = 2.0 * N * S ** 3 * 27 * cin * cout
    print(f"{name} S={S} {cin}->{cout}: {us:6.1f} us  {fl / us / 1e6:6.0f} TFLOP/s   (REPEAT={os.environ.get('B200_TC_REPEAT', '1')} SPLITK={os.environ.get('B200_CONV_SPLITK', '1')})")
