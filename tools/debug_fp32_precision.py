import sys, os, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as TF
from multimodal_segmentation_project_b200 import functional as F
def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()
def cl(x): return x.permute(0, 2, 3, 4, 1).contiguous().cuda()
def cf(x): return x.float().permute(0, 4, 1, 2, 3).contiguous().cpu()
torch.manual_seed(0)
for (N, C0, C, S) in [(2, 64, 128, 4), (2, 16, 16, 32), (2, 128, 256, 2), (2, 32, 64, 8)]:
    conv = torch.nn.Conv3d(C0, C, 3, padding=1); bn = torch.nn.BatchNorm3d(C)
    x = torch.randn(N, C0, S, S, S) + 0.5; gy = torch.randn(N, C, S, S, S)
    cr, br = copy.deepcopy(conv).double(), copy.deepcopy(bn).double()
    xr = x.double().requires_grad_(True)
    pre = cr(xr); pre.retain_grad()
    yr = TF.relu(br(pre)); yr.backward(gy.double())
    cc, bc = copy.deepcopy(conv).cuda(), copy.deepcopy(bn).cuda()
    xc = cl(x).requires_grad_(True)
    y = F.conv_bn_act(xc, None, cc, bc, None, True, impl=1); y.backward(cl(gy))
    # fp32 torch on CPU for calibration
    c32, b32 = copy.deepcopy(conv), copy.deepcopy(bn)
    x32 = x.clone().requires_grad_(True); y32 = TF.relu(b32(c32(x32))); y32.backward(gy)
    print(f"block N{N} {C0}->{C} S{S}: y {rel(cf(y), yr):.1e} (torch32 {rel(y32, yr):.1e}) dx {rel(cf(xc.grad), xr.grad):.1e} ({rel(x32.grad, xr.grad):.1e}) "
          f"dW {rel(cc.weight.grad, cr.weight.grad):.1e} ({rel(c32.weight.grad, cr.weight.grad):.1e}) dgamma {rel(bc.weight.grad, br.weight.grad):.1e} ({rel(b32.weight.grad, br.weight.grad):.1e}) "
          f"dbeta {rel(bc.bias.grad, br.bias.grad):.1e} ({rel(b32.bias.grad, br.bias.grad):.1e})")
for (N, C, S) in [(2, 64, 8), (2, 16, 32)]:
    x = torch.relu(torch.randn(N, C, S, S, S)); gy = torch.randn(N, C, S // 2, S // 2, S // 2)
    xr = x.double().requires_grad_(True); TF.max_pool3d(xr, 2, 2).backward(gy.double())
    xc = cl(x).requires_grad_(True); F.maxpool2(xc).backward(cl(gy))
    print(f"pool C{C} S{S}: dx {rel(cf(xc.grad), xr.grad):.1e}")
for (N, Ci, Co, S) in [(2, 256, 128, 2), (2, 64, 32, 8), (2, 32, 16, 16)]:
    x = torch.randn(N, Ci, S, S, S); w = torch.randn(Ci, Co, 2, 2, 2) / Ci ** 0.5; b = torch.randn(Co); gy = torch.randn(N, Co, 2 * S, 2 * S, 2 * S)
    xr, wr, br = x.double().requires_grad_(True), w.double().requires_grad_(True), b.double().requires_grad_(True)
    TF.conv_transpose3d(xr, wr, br, stride=2).backward(gy.double())
    xc = cl(x).requires_grad_(True); wc, bc = w.cuda().requires_grad_(True), b.cuda().requires_grad_(True)
    F.conv_transpose2(xc, wc, bc).backward(cl(gy))
    print(f"convT {Ci}->{Co} S{S}: dx {rel(cf(xc.grad), xr.grad):.1e} dW {rel(wc.grad, wr.grad):.1e} db {rel(bc.grad, br.grad):.1e}")
