"""Targets for `ncu --set full`: the up-convolution kernels at the top level of the benchmark step (coarse 2 x 64^3 x 32 -> fine 2 x 128^3 x 16)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_segmentation_project_b200 import _lib
L = _lib.load()
dev = torch.device("cuda"); N, Sc = 2, 64
P = lambda t: ctypes.c_void_p(t.data_ptr())
st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
xc = torch.randn(N, Sc, Sc, Sc, 32, device=dev).bfloat16()
fine = torch.randn(N, 2 * Sc, 2 * Sc, 2 * Sc, 16, device=dev).bfloat16()
out = torch.empty_like(fine); gxc = torch.empty_like(xc)
wt, bt = torch.randn(32, 16, 2, 2, 2, device=dev) * 0.1, torch.zeros(16, device=dev)
ws_bytes = L.b200_convt2_wgrad_workspace(32, 16, N, Sc, Sc, Sc)
ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev); dwt = torch.empty_like(wt); dbt = torch.empty_like(bt)
for _ in range(3):
    _lib.check(L.b200_convt2_fwd(1, P(xc), P(wt), P(bt), P(out), N, Sc, Sc, Sc, 32, 16, st()))
    _lib.check(L.b200_convt2_bwd_data(1, P(fine), P(wt), P(gxc), N, Sc, Sc, Sc, 32, 16, st()))
    _lib.check(L.b200_convt2_bwd_weight(1, P(xc), P(fine), P(dwt), P(dbt), P(ws), ws_bytes, N, Sc, Sc, Sc, 32, 16, st()))
torch.cuda.synchronize(); print("ok")
