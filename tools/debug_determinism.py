import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_segmentation_project_b200.models.unet_dann import UNet3D as UNet3DDann
from multimodal_segmentation_project_b200.synthetic import structured_volume
from multimodal_segmentation_project_b200.utils import metrics as M
from oracle.unet_oracle import init_state_dict
sd = init_state_dict(1, 4, seed=0)
S = int(sys.argv[1]) if len(sys.argv) > 1 else 32
xs, ys = structured_volume(2, S, seed=21)
def run(two):
    seg = UNet3DDann(1, 4, dropout_rate=0.0).cuda(); seg.load_state_dict(sd); seg.train()
    o, f = seg(xs.cuda(), return_features=True)
    loss = M.combined_ce_tversky_loss(o, ys.cuda(), alpha=0.5, beta=0.5)
    if two:
        loss = loss + f.sum() * 0.01
    loss.backward()
    torch.cuda.synchronize()
    return o.detach().clone(), {k: p.grad.clone() for k, p in seg.named_parameters()}
for two in (False, True):
    o1, g1 = run(two); o2, g2 = run(two)
    print('two' if two else 'one', 'logits equal', torch.equal(o1, o2))
    bad = [(k, (g1[k] - g2[k]).abs().max().item() / (g1[k].abs().max().item() + 1e-30)) for k in g1 if not torch.equal(g1[k], g2[k])]
    print(len(bad), 'tensors differ run-to-run', sorted(bad, key=lambda t: -t[1])[:8])
