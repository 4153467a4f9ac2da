"""Generates tests/golden/ref_checkpoint_f8_16.pth and ref_checkpoint_next_step.npz with the UNMODIFIED reference
(imported from /root/reference) and torch.optim.AdamW, exactly as train_unet.py:378,221-226,477-486 do:

    python -m oracle.make_golden_checkpoint

TEST INFRASTRUCTURE (see oracle/__init__.py).  The .pth file is what a reference training run writes (model + optimizer
state after two steps); the .npz holds the batch of a third step and the parameters the reference has after it, so a test
can load the checkpoint into the drop-in model + FlatAdamW, take that step on the GPU and compare."""
from __future__ import annotations

import os

import numpy as np
import torch

from .make_golden import OUT, _import_reference


def main():
    ref_unet, _, RM, _ = _import_reference()
    torch.manual_seed(5)
    net = ref_unet.UNet3D(in_channels=1, out_channels=4, features=[8, 16], dropout_rate=0.0)
    net.train()
    opt = torch.optim.AdamW(net.parameters(), lr=1e-3, weight_decay=1e-2)
    g = torch.Generator().manual_seed(21)
    batches = [(torch.rand(2, 1, 16, 16, 16, generator=g), torch.randint(0, 4, (2, 1, 16, 16, 16), generator=g)) for _ in range(3)]
    losses = []
    for x, y in batches[:2]:
        opt.zero_grad()
        loss = RM.combined_loss(net(x), y)
        loss.backward()
        opt.step()
        losses.append(float(loss))
    torch.save({
        "epoch": 25,
        "model_state_dict": net.state_dict(),
        "optimizer_state_dict": opt.state_dict(),
        "train_loss": losses[-1], "val_loss": losses[-1], "train_dice": 0.25, "val_dice": 0.25, "encoder_frozen": False,
    }, os.path.join(OUT, "ref_checkpoint_f8_16.pth"))
    x, y = batches[2]
    opt.zero_grad()
    loss = RM.combined_loss(net(x), y)
    loss.backward()
    opt.step()
    out = {"x": x.numpy(), "y": y.numpy(), "loss": np.float32(float(loss))}
    for k, v in net.state_dict().items():
        out["after/" + k] = v.detach().cpu().numpy()
    st = opt.state_dict()["state"]
    out["exp_avg_0"] = st[0]["exp_avg"].numpy()
    out["step"] = np.float32(float(st[0]["step"]))
    np.savez_compressed(os.path.join(OUT, "ref_checkpoint_next_step.npz"), **out)
    print("checkpoint golden written to", OUT)


if __name__ == "__main__":
    main()
