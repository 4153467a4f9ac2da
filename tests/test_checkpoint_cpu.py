"""Checkpoint wire format, CPU part: files written by the reference (tests/golden/ref_checkpoint_f8_16.pth,
made by oracle/make_golden_checkpoint.py from the unmodified reference) load into the drop-in model; DDP 'module.' prefix
and bare state_dict files are accepted like test_model.py:381-385 / distill_unet.py:20-29 do; files we write have the
reference's keys."""
import io
import os

import pytest
import torch

from multimodal_segmentation_project_b200.checkpoint import load_checkpoint, model_state_from, save_checkpoint
from multimodal_segmentation_project_b200.models.unet import UNet3D

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _ref_ckpt():
    return torch.load(os.path.join(GOLD, "ref_checkpoint_f8_16.pth"), weights_only=False)


def test_reference_checkpoint_loads_into_drop_in_model():
    net = UNet3D(1, 4, features=[8, 16], dropout_rate=0.0)
    ckpt = load_checkpoint(os.path.join(GOLD, "ref_checkpoint_f8_16.pth"), net)
    assert ckpt["epoch"] == 25 and ckpt["encoder_frozen"] is False
    ref_sd = _ref_ckpt()["model_state_dict"]
    sd = net.state_dict()
    assert list(sd.keys()) == list(ref_sd.keys())
    for k in sd:
        assert torch.equal(sd[k], ref_sd[k]), k


def test_module_prefix_and_bare_state_dict():
    ref_sd = _ref_ckpt()["model_state_dict"]
    prefixed = {"module." + k: v for k, v in ref_sd.items()}
    assert list(model_state_from({"model_state_dict": prefixed}).keys()) == list(ref_sd.keys())
    net = UNet3D(1, 4, features=[8, 16])
    out = load_checkpoint(prefixed, net)          # bare state_dict object with DDP prefix
    assert "model_state_dict" in out
    assert torch.equal(net.state_dict()["final_conv.weight"], ref_sd["final_conv.weight"])
    with pytest.raises(RuntimeError):             # strict loading like the reference: wrong architecture fails
        load_checkpoint(ref_sd, UNet3D(1, 4, features=[8, 16, 32]))
    with pytest.raises(KeyError):                 # optimizer requested but the file has none
        load_checkpoint(ref_sd, net, optimizer=torch.optim.AdamW(net.parameters()))


def test_saved_file_has_reference_keys_and_loads_into_torch_adamw():
    net = UNet3D(1, 4, features=[8, 16])
    opt = torch.optim.AdamW(net.parameters(), lr=1e-3, weight_decay=1e-2)
    buf = io.BytesIO()
    save_checkpoint(buf, net, opt, epoch=3, train_loss=1.0, val_loss=2.0, train_dice=0.1, val_dice=0.2, encoder_frozen=True)
    buf.seek(0)
    obj = torch.load(buf, weights_only=False)
    assert set(obj.keys()) == set(_ref_ckpt().keys())
    net2 = UNet3D(1, 4, features=[8, 16])
    opt2 = torch.optim.AdamW(net2.parameters())
    buf.seek(0)
    load_checkpoint(buf, net2, opt2)
    assert opt2.param_groups[0]["weight_decay"] == 1e-2


def test_oracle_continues_reference_checkpoint():
    """CPU pin of the whole chain: reference checkpoint -> oracle forward/backward (functional restatement) -> torch AdamW
    with the checkpoint's optimizer state -> the parameters the reference itself has after its third step."""
    import numpy as np
    from oracle import metrics_oracle as OM
    from oracle.unet_oracle import train_step_grads

    ckpt = _ref_ckpt()
    nxt = np.load(os.path.join(GOLD, "ref_checkpoint_next_step.npz"))
    sd = {k: v.clone() for k, v in ckpt["model_state_dict"].items()}
    x, y = torch.from_numpy(nxt["x"]), torch.from_numpy(nxt["y"])
    loss, _, grads, new_buffers = train_step_grads(sd, x, y, OM.combined_loss)
    assert abs(float(loss) - float(nxt["loss"])) <= 1e-6
    names = [k for k in sd if k in grads]
    params = [torch.nn.Parameter(sd[k].clone()) for k in names]
    opt = torch.optim.AdamW(params, lr=123.0)            # hyper-parameters come from the checkpoint
    opt.load_state_dict(ckpt["optimizer_state_dict"])
    for p, k in zip(params, names):
        p.grad = grads[k]
    opt.step()
    for p, k in zip(params, names):
        if k.endswith("double_conv.0.bias") or k.endswith("double_conv.4.bias"):
            continue  # pre-BatchNorm conv biases: their gradient is round-off noise around 0 and AdamW turns its SIGN into lr-sized steps
        ref = torch.from_numpy(nxt["after/" + k])
        assert torch.allclose(p.detach(), ref, rtol=1e-5, atol=2e-6), k
