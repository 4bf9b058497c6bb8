"""Parameter containers with the reference's module tree (=> identical state-dict keys).

These classes reproduce the *structure* of compress/layers/{gdn,layers,win_attention,masking}.py and
compress/ops/parametrizers.py so that ``state_dict()`` / ``load_state_dict()`` are interchangeable with
the reference model (SURVEY.md §5 "state-dict key layout is part of the drop-in contract").  They hold
parameters only: the arithmetic runs in the CUDA engine (engine.py), so calling ``forward`` on a
container raises instead of silently running a PyTorch/cuDNN path.
"""
from __future__ import annotations

import torch
import torch.nn as nn
from torch import Tensor

from ._lib import PcodecError
from .entropy_models import LowerBound


class _EngineOnly(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover - guard
        raise PcodecError(f"{type(self).__name__} is a parameter container; run it through the model's CUDA engine")


def conv(in_channels, out_channels, kernel_size=5, stride=2):
    """models/utils.py:186-193."""
    return nn.Conv2d(in_channels, out_channels, kernel_size=kernel_size, stride=stride, padding=kernel_size // 2)


def deconv(in_channels, out_channels, kernel_size=5, stride=2):
    """models/utils.py:196-204."""
    return nn.ConvTranspose2d(in_channels, out_channels, kernel_size=kernel_size, stride=stride,
                              output_padding=stride - 1, padding=kernel_size // 2)


def conv3x3(in_ch: int, out_ch: int, stride: int = 1) -> nn.Module:
    return nn.Conv2d(in_ch, out_ch, kernel_size=3, stride=stride, padding=1)


def subpel_conv3x3(in_ch: int, out_ch: int, r: int = 1) -> nn.Sequential:
    """layers/layers.py:20-24."""
    return nn.Sequential(nn.Conv2d(in_ch, out_ch * r ** 2, kernel_size=3, padding=1), nn.PixelShuffle(r))


def conv1x1(in_ch: int, out_ch: int, stride: int = 1) -> nn.Module:
    return nn.Conv2d(in_ch, out_ch, kernel_size=1, stride=stride)


class NonNegativeParametrizer(nn.Module):
    """ops/parametrizers.py:23-49."""

    pedestal: Tensor

    def __init__(self, minimum: float = 0, reparam_offset: float = 2 ** -18):
        super().__init__()
        self.minimum = float(minimum)
        self.reparam_offset = float(reparam_offset)
        pedestal = self.reparam_offset ** 2
        self.register_buffer("pedestal", torch.Tensor([pedestal]))
        bound = (self.minimum + self.reparam_offset ** 2) ** 0.5
        self.lower_bound = LowerBound(bound)

    def init(self, x: Tensor) -> Tensor:
        return torch.sqrt(torch.max(x + self.pedestal, self.pedestal))

    def forward(self, x: Tensor) -> Tensor:
        out = self.lower_bound(x)
        return out ** 2 - self.pedestal


class GDN(_EngineOnly):
    """layers/gdn.py:16-63 (parameters + reparametrisation; the normalisation itself is a fused conv epilogue)."""

    def __init__(self, in_channels: int, inverse: bool = False, beta_min: float = 1e-6, gamma_init: float = 0.1):
        super().__init__()
        self.inverse = bool(inverse)
        self.beta_reparam = NonNegativeParametrizer(minimum=float(beta_min))
        self.beta = nn.Parameter(self.beta_reparam.init(torch.ones(in_channels)))
        self.gamma_reparam = NonNegativeParametrizer()
        self.gamma = nn.Parameter(self.gamma_reparam.init(float(gamma_init) * torch.eye(in_channels)))

    def effective(self):
        """(beta [C], gamma [C_out, C_in]) after reparametrisation (gdn.py:53-55)."""
        with torch.no_grad():
            return self.beta_reparam(self.beta), self.gamma_reparam(self.gamma)


class ResidualUnit(_EngineOnly):
    """layers/layers.py:39-59."""

    def __init__(self, N: int):
        super().__init__()
        self.conv = nn.Sequential(conv1x1(N, N // 2), nn.GELU(), conv3x3(N // 2, N // 2), nn.GELU(), conv1x1(N // 2, N))
        self.relu = nn.GELU()


class WindowAttention(_EngineOnly):
    """layers/win_attention.py:37-82 (parameters, relative-position index buffer)."""

    def __init__(self, dim=192, window_size=(8, 8), num_heads=8, qkv_bias=True):
        super().__init__()
        self.dim, self.window_size, self.num_heads = dim, window_size, num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.relative_position_bias_table = nn.Parameter(
            torch.zeros((2 * window_size[0] - 1) * (2 * window_size[1] - 1), num_heads))
        coords = torch.stack(torch.meshgrid([torch.arange(window_size[0]), torch.arange(window_size[1])], indexing="ij"))
        flat = torch.flatten(coords, 1)
        rel = (flat[:, :, None] - flat[:, None, :]).permute(1, 2, 0).contiguous()
        rel[:, :, 0] += window_size[0] - 1
        rel[:, :, 1] += window_size[1] - 1
        rel[:, :, 0] *= 2 * window_size[1] - 1
        self.register_buffer("relative_position_index", rel.sum(-1))
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)
        nn.init.trunc_normal_(self.relative_position_bias_table, std=0.02)

    def bias_matrix(self) -> Tensor:
        """[heads, T, T] relative-position bias (win_attention.py:97-100)."""
        T = self.window_size[0] * self.window_size[1]
        with torch.no_grad():
            b = self.relative_position_bias_table[self.relative_position_index.view(-1)].view(T, T, -1)
            return b.permute(2, 0, 1).contiguous()


class WinBasedAttention(_EngineOnly):
    """layers/win_attention.py:118-207."""

    def __init__(self, dim=192, num_heads=8, window_size=8, shift_size=0):
        super().__init__()
        assert 0 <= shift_size < window_size, "shift_size must in 0-window_size"
        self.dim, self.num_heads, self.window_size, self.shift_size = dim, num_heads, window_size, shift_size
        self.attn = WindowAttention(dim, window_size=(window_size, window_size), num_heads=num_heads)
        self.drop_path = nn.Identity()


class Win_noShift_Attention(_EngineOnly):
    """layers/layers.py:31-75."""

    def __init__(self, dim, num_heads=8, window_size=8, shift_size=0):
        super().__init__()
        self.dim, self.num_heads, self.window_size, self.shift_size = dim, num_heads, window_size, shift_size
        self.conv_a = nn.Sequential(ResidualUnit(dim), ResidualUnit(dim), ResidualUnit(dim))
        self.conv_b = nn.Sequential(WinBasedAttention(dim=dim, num_heads=num_heads, window_size=window_size,
                                                      shift_size=shift_size),
                                    ResidualUnit(dim), ResidualUnit(dim), ResidualUnit(dim), conv1x1(dim, dim))


class ChannelMask(_EngineOnly):
    """layers/masking.py:9-295.  Only the parameter-free policies used on the inference path are supported
    ('point-based-std', 'two-levels', None); the learnable / random / scalable_res ablation policies are
    training-time experiments outside the hot path."""

    SUPPORTED = ("point-based-std", "two-levels", None)

    def __init__(self, mask_policy, scalable_levels, dim_chunk, num_levels, gamma_bound=1e-9, double_dim=False):
        super().__init__()
        self.mask_policy = mask_policy
        self.scalable_levels = scalable_levels
        self.quality_list = list(range(scalable_levels))
        self.dim_chunk = dim_chunk
        self.num_levels = num_levels
        self.double_dim = double_dim

    @staticmethod
    def mode_for(mask_pol, pr):
        """-> ("ones"|"zeros"|"threshold", q) following masking.py:199-226.  Raises NotImplementedError for
        policies without an inference role (the reference does the same for unknown names, masking.py:295)."""
        if mask_pol is None:
            return "ones", None
        if mask_pol == "point-based-std":
            if pr >= 10:
                return "ones", None
            if pr == 0:
                return "zeros", None
            return "threshold", 1.0 - (pr * 0.1)
        if mask_pol == "two-levels":
            return ("zeros", None) if pr == 0 else ("ones", None)
        raise NotImplementedError(f"mask policy {mask_pol!r} is not on the B200 inference path")


class ResidualBlock(_EngineOnly):
    """models/utils.py:59-87: conv3x3 - LeakyReLU - conv3x3 - LeakyReLU, plus identity (1x1 `skip` conv when the channel
    count changes).  Parameter container of the REM wrapper's LatentRateReduction nets."""

    def __init__(self, in_ch: int, out_ch: int):
        super().__init__()
        self.conv1 = conv3x3(in_ch, out_ch)
        self.nonlin = nn.LeakyReLU(inplace=True)
        self.conv2 = conv3x3(out_ch, out_ch)
        self.skip = conv1x1(in_ch, out_ch) if in_ch != out_ch else None


class LatentRateReduction(_EngineOnly):
    """CHProgREM.py:12-85 (module tree only; the arithmetic runs in rem.py on the CUDA engine)."""

    def __init__(self, dim_chunk: int = 32, mu_std: bool = False, dimension: str = "middle"):
        super().__init__()
        self.dim_block = dim_chunk
        self.mu_std = mu_std
        N = dim_chunk
        extra = 1 if dimension == "big" else 0
        self.enc_base_entropy_params = nn.Sequential(ResidualBlock(2 * N, N), *[ResidualBlock(N, N) for _ in range(1 + extra)])
        self.enc_enh_entropy_params = nn.Sequential(ResidualBlock(2 * N if mu_std else N, N),
                                                    *[ResidualBlock(N, N) for _ in range(1 + extra)])
        self.enc_base_rep = nn.Sequential(*[ResidualBlock(N, N) for _ in range(2 + extra)])
        self.enc = nn.Sequential(ResidualBlock(3 * N, 2 * N), *[ResidualBlock(2 * N, 2 * N) for _ in range(1 + extra)],
                                 ResidualBlock(2 * N, 2 * N if mu_std else N))
