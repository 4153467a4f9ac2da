"""Checkpoint wire format of the reference.

The reference writes ``{'epoch', 'model_state_dict', 'optimizer_state_dict', 'train_loss', 'val_loss', 'train_dice',
'val_dice', 'encoder_frozen'}`` with ``accelerator.save`` (train_unet.py:477-486) and reads either that dictionary or a
bare ``state_dict``, with an optional ``module.`` prefix left by DDP (distill_unet.py:20-29, test_model.py:381-385).
Because the drop-in modules keep the reference's 136-entry ``state_dict`` and ``FlatAdamW`` speaks
``torch.optim.AdamW``'s ``state_dict`` format, files move in both directions unchanged."""
from __future__ import annotations

from typing import Any, Dict, Optional

import torch


def model_state_from(checkpoint: Any) -> Dict[str, torch.Tensor]:
    """The model state_dict inside a checkpoint object (dictionary with 'model_state_dict' or a bare state_dict), with
    DDP's 'module.' prefix removed exactly as test_model.py:384 does (str.replace)."""
    sd = checkpoint["model_state_dict"] if isinstance(checkpoint, dict) and "model_state_dict" in checkpoint else checkpoint
    return {k.replace("module.", ""): v for k, v in sd.items()}


def load_checkpoint(path_or_obj, model: torch.nn.Module, optimizer: Optional[torch.optim.Optimizer] = None, map_location=None,
                    strict: bool = True) -> Dict[str, Any]:
    """Loads a reference-format checkpoint (file path or already-loaded object) into ``model`` (and ``optimizer``).
    Returns the checkpoint dictionary (a bare state_dict is returned as {'model_state_dict': ...})."""
    ckpt = torch.load(path_or_obj, map_location=map_location, weights_only=False) if isinstance(path_or_obj, (str, bytes)) or hasattr(path_or_obj, "read") \
        else path_or_obj
    model.load_state_dict(model_state_from(ckpt), strict=strict)
    is_full = isinstance(ckpt, dict) and "model_state_dict" in ckpt
    if optimizer is not None:
        if not is_full or ckpt.get("optimizer_state_dict") is None:
            raise KeyError("checkpoint holds no 'optimizer_state_dict'")
        optimizer.load_state_dict(ckpt["optimizer_state_dict"])
    return ckpt if is_full else {"model_state_dict": ckpt}


def save_checkpoint(path, model: torch.nn.Module, optimizer: Optional[torch.optim.Optimizer] = None, epoch: int = 0, **extra) -> None:
    """Writes the reference's dictionary (train_unet.py:477-486); ``extra`` carries train_loss / val_dice / encoder_frozen ..."""
    module = getattr(model, "module", model)  # unwrap DDP-style containers like accelerator.unwrap_model
    obj = {"epoch": int(epoch), "model_state_dict": module.state_dict(),
           "optimizer_state_dict": optimizer.state_dict() if optimizer is not None else None}
    obj.update(extra)
    torch.save(obj, path)
