"""One launch of every tcgen05 / TMEM / mbarrier kernel on the smallest shape that selects it, for compute-sanitizer
(memcheck / racecheck / synccheck, one tool per run): conv3d_tc4 (plain, stats, BN-backward epilogue), conv3d_tc3 (plain, stats),
conv3d_tc2 (+ split-K), wgrad_tc2 / wgrad_tc3, ConvTranspose kernels, fused head forward / backward, one 32^3 train step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_segmentation_project_b200 import _lib, functional as F
from multimodal_segmentation_project_b200.models.unet import UNet3D
from multimodal_segmentation_project_b200.synthetic import structured_volume
from multimodal_segmentation_project_b200.utils import metrics as M

dev = torch.device("cuda")
g = torch.Generator(device="cuda").manual_seed(0)
def rnd(*s): return torch.randn(*s, device=dev, generator=g).bfloat16()

def conv_block(N, D, H, W, cin, cout):
    conv = torch.nn.Conv3d(cin, cout, 3, padding=1).cuda(); bn = torch.nn.BatchNorm3d(cout).cuda()
    conv2 = torch.nn.Conv3d(cout, cout, 3, padding=1).cuda(); bn2 = torch.nn.BatchNorm3d(cout).cuda()
    x = rnd(N, D, H, W, cin).requires_grad_(True)
    h, ctx = F.conv_bn_act(x, None, conv, bn, None, True, return_ctx=True)
    y = F.conv_bn_act(h, None, conv2, bn2, None, True, prev=ctx)
    y.float().square().mean().backward()
    torch.cuda.synchronize()

conv_block(2, 24, 200, 128, 16, 16); print("row-streaming kernel (fprop + stats, dgrad + BN-backward epilogue), wgrad_tc2: ok", flush=True)
conv_block(2, 44, 64, 64, 32, 32); print("persistent window kernel (fprop + stats, dgrad), wgrad_tc3: ok", flush=True)
conv_block(1, 8, 16, 16, 64, 128); print("two-CTA kernel with split-K, wide wgrad: ok", flush=True)
up = torch.nn.ConvTranspose3d(64, 32, 2, stride=2).cuda()
xt = rnd(1, 4, 8, 8, 64).requires_grad_(True)
F.conv_transpose2(xt, up.weight, up.bias).float().square().mean().backward(); torch.cuda.synchronize(); print("ConvTranspose kernels: ok", flush=True)
x, y = structured_volume(2, 32, seed=1)
net = UNet3D(1, 4, dropout_rate=0.0).cuda().train()
with torch.autocast("cuda", dtype=torch.bfloat16):
    _, loss, conf = net.forward_with_loss(x.cuda().bfloat16(), y.cuda().to(torch.uint8), M.combined_loss, want_confusion=True)
loss.backward(); torch.cuda.synchronize()
print("32^3 train step with the fused head: ok, loss", float(loss), "counts", int(conf.sum()), flush=True)
