"""Shim for the three timm helpers reference layers/win_attention.py:3 imports."""
import torch.nn as nn
from torch.nn.init import trunc_normal_  # noqa: F401


def to_2tuple(x):
    return tuple(x) if isinstance(x, (tuple, list)) else (x, x)


class DropPath(nn.Identity):
    def __init__(self, drop_prob=0.0):
        super().__init__()
