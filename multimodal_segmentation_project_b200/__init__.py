"""B200-native (sm_100a) implementation of the 3D U-Net segmentation hot path of
fransiskusbudi/multimodal_segmentation_project, behind the reference's own Python API.

    from multimodal_segmentation_project_b200.models.unet import UNet3D
    from multimodal_segmentation_project_b200.models.unet_dann import UNet3D as UNet3DDann
    from multimodal_segmentation_project_b200.utils.metrics import combined_loss, calculate_dice, ...
    from multimodal_segmentation_project_b200.train_dann import grad_reverse, DomainDiscriminator
"""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
__version__ = "0.1.0"
