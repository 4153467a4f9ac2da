"""Timing of SURVEY §8d configurations 2-5 on one B200 (CUDA events, eager launches; not collected by pytest).

  cfg 2  train step 2x128^3 bf16: this library (eager; bench.py reports the graph-captured number) next to the SAME step
         written with torch ops (the oracle restatement of the reference) under CUDA bf16 autocast = PyTorch eager + cuDNN
         on the same GPU ("second bar" of §8d);
  cfg 3  distillation step: frozen teacher forward + student forward/backward, distillation_loss(alpha=.7, T=2);
  cfg 4  DANN step: source + target forwards, ce_tversky task loss, domain CE through gradient reversal;
  cfg 5  512x512x256 volume, sliding window 128^3 stride 128 and stride 64, eval-mode forward + stitched confusion counts.

Lives under tests/ because it runs the oracle (test infrastructure) as a baseline; run with  python tests/bench_configs.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from multimodal_segmentation_project_b200 import functional as F
from multimodal_segmentation_project_b200.inference import evaluate_volume
from multimodal_segmentation_project_b200.models.unet import UNet3D
from multimodal_segmentation_project_b200.models.unet_dann import UNet3D as UNet3DDann
from multimodal_segmentation_project_b200.synthetic import structured_volume
from multimodal_segmentation_project_b200.train_dann import DomainDiscriminator, domain_cross_entropy, grad_reverse
from multimodal_segmentation_project_b200.utils import metrics as M
from oracle import metrics_oracle as OM
from oracle.unet_oracle import clone_for_autograd, init_state_dict, train_step_grads

dev = torch.device("cuda")


def timed(fn, warm=2, iters=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(e))
    return sorted(ts)[len(ts) // 2]


sd = init_state_dict(1, 4, seed=0)
x, y = structured_volume(2, 128, seed=1234)
xc, yc = x.to(dev), y.to(dev)
vox = 2 * 128 ** 3

# ---- cfg 2 -----------------------------------------------------------------------------------------------------------
net = UNet3D(1, 4, dropout_rate=0.0).to(dev).train(); net.load_state_dict(sd)

def ours_step():
    net.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = net(xc)
    M.combined_loss(out.float(), yc).backward()

ms = timed(ours_step)
print(f"cfg2 ours  (eager fwd+loss+bwd, bf16)            : {ms:8.2f} ms  {vox / ms / 1e3:8.1f} Mvox/s")
sd_cuda = {k: v.to(dev) for k, v in sd.items()}

def torch_step():
    train_step_grads(sd_cuda, xc, yc, OM.combined_loss, autocast_dtype=torch.bfloat16)

torch.backends.cudnn.benchmark = True
ms_t = timed(torch_step)
print(f"cfg2 torch (eager, bf16 autocast, cuDNN, same GPU): {ms_t:8.2f} ms  {vox / ms_t / 1e3:8.1f} Mvox/s   -> ours is {ms_t / ms:.1f}x faster")

# ---- cfg 3 -----------------------------------------------------------------------------------------------------------
teacher = UNet3D(1, 4, dropout_rate=0.0).to(dev).eval(); teacher.load_state_dict(init_state_dict(1, 4, seed=1))
for q in teacher.parameters():
    q.requires_grad = False

def kd_step():
    net.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        sl = net(xc)
        with torch.no_grad():
            tl = teacher(xc)
    M.distillation_loss(sl.float(), tl.float(), yc, alpha=0.7, temperature=2.0).backward()

ms = timed(kd_step)
print(f"cfg3 distillation step (teacher fwd + student fwd/bwd): {ms:8.2f} ms  {vox / ms / 1e3:8.1f} Mvox/s")

# ---- cfg 4 -----------------------------------------------------------------------------------------------------------
seg = UNet3DDann(1, 4, dropout_rate=0.0).to(dev).train(); seg.load_state_dict(sd)
disc = DomainDiscriminator(256).to(dev).train()
xs, ys = structured_volume(1, 128, seed=21)
xt, _ = structured_volume(1, 128, seed=22)
xs, ys, xt = xs.to(dev), ys.to(dev), xt.to(dev)
dom_labels = torch.tensor([0, 1], device=dev)

def dann_step(lam=0.2):
    seg.zero_grad(set_to_none=True); disc.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        o_s, f_s = seg(xs, return_features=True)
        _, f_t = seg(xt, return_features=True)
    task = M.combined_ce_tversky_loss(o_s.float(), ys, alpha=0.5, beta=0.5)
    dom = domain_cross_entropy(torch.cat([disc(grad_reverse(f_s, lam)), disc(grad_reverse(f_t, lam))], 0), dom_labels)
    (task + lam * dom).backward()

ms = timed(dann_step)
print(f"cfg4 DANN step (source + target 1x128^3 each)     : {ms:8.2f} ms  {vox / ms / 1e3:8.1f} Mvox/s")

# ---- cfg 5 -----------------------------------------------------------------------------------------------------------
del xc, yc
torch.cuda.empty_cache()
vol = torch.rand(1, 1, 256, 512, 512, device=dev)
lab = torch.randint(0, 4, (1, 1, 256, 512, 512), device=dev)
net.eval(); net.compute_dtype = torch.bfloat16
nv = vol.numel()
for stride in (128, 64):
    with torch.no_grad():
        ms = timed(lambda: evaluate_volume(net, vol, lab, window=128, stride=stride), warm=1, iters=3)
    print(f"cfg5 sliding window 128^3 stride {stride:3d} on 512x512x256 (+ confusion counts, organ metrics): {ms:8.1f} ms  {nv / ms / 1e3:8.1f} Mvox/s of volume")
