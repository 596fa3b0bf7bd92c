"""Torch denoiser scaffolding (random-init; checkpoints are offline).  NOT the product: the north-star keeps the
denoiser forward in torch.  These modules exist so the sampler can be driven end to end by networks with the
architecture, parameter count and I/O contract of the three the reference uses:
NCSN++ (VP, continuous) for CIFAR-10, DiT-XL/2, and an SD3-medium-shaped MMDiT stand-in."""
from .unet_vp import NCSNppVP  # noqa: F401
from .dit import DiT, dit_xl_2  # noqa: F401
from .mmdit import MMDiT, mmdit_sd3_medium  # noqa: F401
