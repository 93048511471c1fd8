"""Drop-in for the external `DCN` package (tteepe/DCNv2) the reference imports at
models/backbones/pose_dla_dcn.py:11 and resnet_dcn.py:14."""
from .dcn_v2 import DCN

__all__ = ["DCN"]
