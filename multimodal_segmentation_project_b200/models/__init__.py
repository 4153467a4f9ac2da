from .unet import DoubleConv, UNet3D  # noqa: F401
