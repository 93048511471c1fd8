from .ctdet import ctdet_decode
from .multi_pose import multi_pose_decode

__all__ = ["ctdet_decode", "multi_pose_decode"]
