import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_segmentation_project_b200 import _lib, functional as F
dev = torch.device("cuda"); S, N = 128, 2
cin, cout = int(sys.argv[1]), int(sys.argv[2])
x = torch.randn(N, S, S, S, cin, device=dev).bfloat16()
w = torch.randn(cout, cin, 3, 3, 3, device=dev) * 0.05
b = torch.randn(cout, device=dev)
wp = F.pack_conv3_weights(w, _lib.PACK_FPROP_TC, torch.bfloat16)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ts = []
for i in range(8):
    flush.zero_()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); F.conv3d_k3_raw(x, None, wp, b, cout, 0, impl=2); e.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(e))
dy = torch.randn(N, S, S, S, cout, device=dev).bfloat16()
tw = []
for i in range(6):
    flush.zero_()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); F.conv3d_wgrad_raw(x, None, dy, want_bias=False); e.record(); torch.cuda.synchronize()
    tw.append(a.elapsed_time(e))
print(f"{cin}->{cout} wgrad {sorted(tw)[len(tw)//2]*1e3:.1f} us  (B200_WG_REPEAT={os.environ.get('B200_WG_REPEAT','1')})")
print(f"{cin}->{cout} fprop {sorted(ts)[len(ts)//2]*1e3:.1f} us  (B200_TC_REPEAT={os.environ.get('B200_TC_REPEAT','1')})")
