"""CPU ORACLE (test infrastructure): entropy-model restatement.

Restates, in torch-CPU / numpy, the reference's ``EntropyBottleneck`` and ``GaussianConditional``
(entropy_models/entropy_models.py) as table holders + pure functions, and wraps the two coders
the tests compare against:

  * ``CPortCoder``  — oracle/rans_port.c via ctypes (our plain-C restatement), and
  * ``RefCoder``    — the reference's own rans_interface.cpp compiled into oracle/_ref.

Never imported by the product package.
"""
from __future__ import annotations

import ctypes
import math
import os
import subprocess
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(HERE, "librans_port.so")


# ----------------------------------------------------------------------------------------------
# coders
# ----------------------------------------------------------------------------------------------

def build_c_port(force: bool = False) -> str:
    src = os.path.join(HERE, "rans_port.c")
    if force or not os.path.isfile(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", _LIB_PATH, src, "-lm"])
    return _LIB_PATH


_lib = None


def _c_lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build_c_port())
        vp, cl = ctypes.c_void_p, ctypes.c_long
        _lib.rans_port_encode.restype = cl
        _lib.rans_port_encode.argtypes = [vp, vp, cl, vp, cl, vp, vp, vp, cl]
        _lib.rans_port_decode.restype = cl
        _lib.rans_port_decode.argtypes = [vp, cl, vp, cl, vp, cl, vp, vp, vp]
        _lib.pmf_to_quantized_cdf_port.restype = ctypes.c_int
        _lib.pmf_to_quantized_cdf_port.argtypes = [vp, ctypes.c_int, ctypes.c_int, vp]
    return _lib


def _i32(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.int32))


class CPortCoder:
    """Same call shape as compressai.ans.RansEncoder/RansDecoder (rans_interface.cpp:193-275) on numpy."""

    name = "c-port"

    def encode_with_indexes(self, symbols, indexes, cdfs, cdf_sizes, offsets) -> bytes:
        lib = _c_lib()
        sy, ix, cd, cs, of = _i32(symbols).ravel(), _i32(indexes).ravel(), _i32(cdfs), _i32(cdf_sizes), _i32(offsets)
        cap = 4 * (sy.size * 12 + 8)
        out = np.empty(cap, dtype=np.uint8)
        n = lib.rans_port_encode(sy.ctypes.data, ix.ctypes.data, sy.size, cd.ctypes.data, cd.shape[1],
                                 cs.ctypes.data, of.ctypes.data, out.ctypes.data, cap)
        assert n >= 0
        return out[:n].tobytes()

    def decode_with_indexes(self, encoded: bytes, indexes, cdfs, cdf_sizes, offsets) -> np.ndarray:
        lib = _c_lib()
        ix, cd, cs, of = _i32(indexes).ravel(), _i32(cdfs), _i32(cdf_sizes), _i32(offsets)
        buf = np.frombuffer(encoded, dtype=np.uint8)
        out = np.empty(ix.size, dtype=np.int32)
        lib.rans_port_decode(buf.ctypes.data, buf.size, ix.ctypes.data, ix.size, cd.ctypes.data, cd.shape[1],
                             cs.ctypes.data, of.ctypes.data, out.ctypes.data)
        return out


class RefCoder:
    """The reference's own compiled coder (oracle/_ref). Python lists across pybind11, as in
    entropy_models.py:227-235, 276-286."""

    name = "reference"

    def __init__(self):
        from . import build_ref

        self.ans, self.cxx = build_ref.import_ref_coder()
        self._enc = self.ans.RansEncoder()
        self._dec = self.ans.RansDecoder()

    def encode_with_indexes(self, symbols, indexes, cdfs, cdf_sizes, offsets) -> bytes:
        return self._enc.encode_with_indexes(_i32(symbols).ravel().tolist(), _i32(indexes).ravel().tolist(),
                                             _i32(cdfs).tolist(), _i32(cdf_sizes).tolist(), _i32(offsets).tolist())

    def decode_with_indexes(self, encoded: bytes, indexes, cdfs, cdf_sizes, offsets) -> np.ndarray:
        return np.asarray(self._dec.decode_with_indexes(encoded, _i32(indexes).ravel().tolist(), _i32(cdfs).tolist(),
                                                        _i32(cdf_sizes).tolist(), _i32(offsets).tolist()), dtype=np.int32)


_default = None


def default_coder():
    global _default
    if _default is None:
        _default = CPortCoder()
    return _default


def pmf_to_quantized_cdf(pmf: Sequence[float], precision: int = 16) -> np.ndarray:
    """ops.cpp:10-67 through the C port."""
    p = np.ascontiguousarray(np.asarray(pmf, dtype=np.float32))
    cdf = np.empty(p.size + 1, dtype=np.uint32)
    rc = _c_lib().pmf_to_quantized_cdf_port(p.ctypes.data, p.size, precision, cdf.ctypes.data)
    assert rc == 0
    return cdf.astype(np.int64).astype(np.int32)


def pmf_to_cdf_table(pmf: Tensor, tail_mass: Tensor, pmf_length: Tensor, max_length: int) -> Tensor:
    """EntropyModel._pmf_to_cdf — entropy_models.py:172-180."""
    cdf = torch.zeros((len(pmf_length), max_length + 2), dtype=torch.int32)
    for i in range(len(pmf_length)):
        prob = torch.cat((pmf[i, : int(pmf_length[i])], tail_mass[i].reshape(1)))
        row = pmf_to_quantized_cdf(prob.tolist())
        cdf[i, : row.size] = torch.from_numpy(row)
    return cdf


# ----------------------------------------------------------------------------------------------
# quantile mask (numpy restatement of torch.quantile's linear interpolation)
# ----------------------------------------------------------------------------------------------

def quantile_threshold_np(values: np.ndarray, q: float) -> np.float32:
    """torch.quantile(v, q) for 1-D fp32 v: sort, rank = fp32(q)*(n-1) in fp32, lerp between the two
    order statistics with torch's lerp formula (w<0.5 ? a+w(b-a) : b-(b-a)(1-w))."""
    v = np.sort(np.asarray(values, dtype=np.float32).ravel())
    n = v.size
    rank = np.float32(q) * np.float32(n - 1)
    lo = np.floor(rank)
    w = np.float32(rank - lo)
    lo_i = int(lo)
    hi_i = min(lo_i + 1, n - 1)
    a, b = v[lo_i], v[hi_i]
    d = np.float32(b - a)
    if w < np.float32(0.5):
        return np.float32(a + np.float32(w * d))
    return np.float32(b - np.float32(d * np.float32(np.float32(1) - w)))


def point_based_std_mask_np(scale: np.ndarray, pr: float) -> np.ndarray:
    """ChannelMask 'point-based-std' — layers/masking.py:205-223, scale [B,C,H,W]."""
    s = np.asarray(scale, dtype=np.float32)
    if pr >= 10:
        return np.ones_like(s)
    if pr == 0:
        return np.zeros_like(s)
    q = 1.0 - pr * 0.1
    out = np.zeros_like(s)
    for j in range(s.shape[0]):
        out[j] = (s[j] >= quantile_threshold_np(s[j], q)).astype(np.float32)
    return out


# ----------------------------------------------------------------------------------------------
# GaussianConditional
# ----------------------------------------------------------------------------------------------

def get_scale_table(lo: float = 0.11, hi: float = 256.0, levels: int = 64) -> Tensor:
    """models/cnn.py:14-20 — the table the model actually uses."""
    return torch.exp(torch.linspace(math.log(lo), math.log(hi), levels))


def _std_cumulative(x: Tensor) -> Tensor:
    """entropy_models.py:575-579."""
    return 0.5 * torch.erfc(float(-(2 ** -0.5)) * x)


@dataclass
class GaussianTables:
    scale_table: Tensor
    cdf: Tensor  # int32 [T, Lmax]
    cdf_length: Tensor
    offset: Tensor
    scale_bound: float = 0.11

    @staticmethod
    def from_state_dict(sd: Dict[str, Tensor], prefix: str) -> "GaussianTables":
        return GaussianTables(sd[prefix + ".scale_table"].float(), sd[prefix + "._quantized_cdf"].int(),
                              sd[prefix + "._cdf_length"].int(), sd[prefix + "._offset"].int(),
                              float(sd[prefix + ".scale_bound"].item()) if (prefix + ".scale_bound") in sd else 0.11)

    @staticmethod
    def build(scale_table: Optional[Tensor] = None, tail_mass: float = 1e-9) -> "GaussianTables":
        """GaussianConditional.update — entropy_models.py:599-624."""
        import scipy.stats

        st = get_scale_table() if scale_table is None else scale_table.float()
        multiplier = -scipy.stats.norm.ppf(tail_mass / 2)
        center = torch.ceil(st * multiplier).int()
        length = 2 * center + 1
        max_len = int(length.max())
        samples = torch.abs(torch.arange(max_len).int() - center[:, None]).float()
        sc = st.unsqueeze(1).float()
        upper = _std_cumulative((0.5 - samples) / sc)
        lower = _std_cumulative((-0.5 - samples) / sc)
        pmf = upper - lower
        tail = 2 * lower[:, :1]
        cdf = pmf_to_cdf_table(pmf, tail, length, max_len)
        return GaussianTables(st, cdf, (length + 2).int(), (-center).int())

    def build_indexes(self, scales: Tensor) -> Tensor:
        """entropy_models.py:661-666."""
        s = torch.clamp_min(scales, self.scale_bound)
        idx = torch.full(s.shape, len(self.scale_table) - 1, dtype=torch.int32)
        for t in self.scale_table[:-1]:
            idx -= (s <= t).int()
        return idx

    def likelihood(self, inputs: Tensor, scales: Tensor, means: Optional[Tensor]) -> Tensor:
        """GaussianConditional.forward(training=False) — entropy_models.py:626-659 (+ bound 1e-9)."""
        if means is not None:
            outputs = torch.round(inputs - means) + means
            values = outputs - means
        else:
            values = torch.round(inputs)
        s = torch.clamp_min(scales, self.scale_bound)
        values = torch.abs(values)
        lik = _std_cumulative((0.5 - values) / s) - _std_cumulative((-0.5 - values) / s)
        return torch.clamp_min(lik, 1e-9)

    def encode(self, symbols: Tensor, indexes: Tensor, coder) -> List[bytes]:
        """EntropyModel.compress loop — entropy_models.py:225-236 (one stream per batch item)."""
        cd, cs, of = self.cdf.numpy(), self.cdf_length.numpy(), self.offset.numpy()
        return [coder.encode_with_indexes(symbols[b].reshape(-1).numpy(), indexes[b].reshape(-1).numpy(), cd, cs, of)
                for b in range(symbols.shape[0])]

    def decode(self, strings: Sequence[bytes], indexes: Tensor, coder) -> Tensor:
        """EntropyModel.decompress loop — entropy_models.py:272-287."""
        cd, cs, of = self.cdf.numpy(), self.cdf_length.numpy(), self.offset.numpy()
        out = torch.empty(indexes.shape, dtype=torch.int32)
        for b, s in enumerate(strings):
            v = coder.decode_with_indexes(s, indexes[b].reshape(-1).numpy(), cd, cs, of)
            out[b] = torch.from_numpy(np.asarray(v, dtype=np.int32)).reshape(indexes[b].shape)
        return out


# ----------------------------------------------------------------------------------------------
# EntropyBottleneck
# ----------------------------------------------------------------------------------------------

@dataclass
class BottleneckTables:
    matrices: List[Tensor]
    biases: List[Tensor]
    factors: List[Tensor]
    quantiles: Tensor  # [C,1,3]
    cdf: Tensor
    cdf_length: Tensor
    offset: Tensor

    @staticmethod
    def from_state_dict(sd: Dict[str, Tensor], prefix: str) -> "BottleneckTables":
        n = 0
        while f"{prefix}._matrix{n}" in sd:
            n += 1
        return BottleneckTables([sd[f"{prefix}._matrix{i}"].float() for i in range(n)],
                                [sd[f"{prefix}._bias{i}"].float() for i in range(n)],
                                [sd[f"{prefix}._factor{i}"].float() for i in range(n - 1)],
                                sd[prefix + ".quantiles"].float(), sd[prefix + "._quantized_cdf"].int(),
                                sd[prefix + "._cdf_length"].int(), sd[prefix + "._offset"].int())

    @property
    def medians(self) -> Tensor:
        """_get_medians — entropy_models.py:350-352, as a [C] vector."""
        return self.quantiles[:, 0, 1]

    def logits_cumulative(self, v: Tensor) -> Tensor:
        """entropy_models.py:400-419; v is [C,1,K]."""
        logits = v
        for i, (m, b) in enumerate(zip(self.matrices, self.biases)):
            logits = torch.matmul(F.softplus(m), logits) + b
            if i < len(self.factors):
                logits = logits + torch.tanh(self.factors[i]) * torch.tanh(logits)
        return logits

    def likelihood(self, z_hat: Tensor) -> Tensor:
        """EntropyBottleneck.forward(training=False)'s likelihood — entropy_models.py:421-433, 446-489.

        forward() re-quantises its input: round(z - med) + med; z_hat already has that form and the
        operation is idempotent in fp32 for the value ranges involved, but we replay it literally."""
        B, C = z_hat.shape[:2]
        v = z_hat.permute(1, 0, 2, 3).reshape(C, 1, -1)
        med = self.quantiles[:, :, 1:2]
        v = torch.round(v - med) + med
        lower = self.logits_cumulative(v - 0.5)
        upper = self.logits_cumulative(v + 0.5)
        sign = -torch.sign(lower + upper)
        lik = torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower))
        lik = torch.clamp_min(lik, 1e-9)
        return lik.reshape(C, B, *z_hat.shape[2:]).permute(1, 0, 2, 3).contiguous()

    def rebuild(self) -> None:
        """EntropyBottleneck.update — entropy_models.py:354-393."""
        q = self.quantiles
        med = q[:, 0, 1]
        minima = torch.clamp(torch.ceil(med - q[:, 0, 0]).int(), min=0)
        maxima = torch.clamp(torch.ceil(q[:, 0, 2] - med).int(), min=0)
        self.offset = -minima
        pmf_start = med - minima
        pmf_length = maxima + minima + 1
        max_len = int(pmf_length.max())
        samples = torch.arange(max_len)[None, :] + pmf_start[:, None, None]
        lower = self.logits_cumulative(samples - 0.5)
        upper = self.logits_cumulative(samples + 0.5)
        sign = -torch.sign(lower + upper)
        pmf = torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower))[:, 0, :]
        tail = torch.sigmoid(lower[:, 0, :1]) + torch.sigmoid(-upper[:, 0, -1:])
        self.cdf = pmf_to_cdf_table(pmf, tail, pmf_length, max_len)
        self.cdf_length = (pmf_length + 2).int()

    def symbols(self, z: Tensor) -> Tensor:
        """entropy_models.py:508-515 + 126-150: round(z - median).int()."""
        return torch.round(z - self.medians.reshape(1, -1, 1, 1)).int()

    def dequantize(self, sym: Tensor) -> Tensor:
        """entropy_models.py:152-165, 517-522."""
        return sym.float() + self.medians.reshape(1, -1, 1, 1)

    def _indexes(self, shape) -> Tensor:
        """_build_indexes — entropy_models.py:491-502: channel id."""
        B, C, H, W = shape
        return torch.arange(C, dtype=torch.int32).reshape(1, C, 1, 1).expand(B, C, H, W)

    def encode(self, sym: Tensor, coder) -> List[bytes]:
        idx = self._indexes(sym.shape)
        cd, cs, of = self.cdf.numpy(), self.cdf_length.numpy(), self.offset.numpy()
        return [coder.encode_with_indexes(sym[b].reshape(-1).numpy(), idx[b].reshape(-1).numpy(), cd, cs, of)
                for b in range(sym.shape[0])]

    def decode(self, strings: Sequence[bytes], size: Tuple[int, int], coder) -> Tensor:
        C = self.cdf.shape[0]
        shape = (len(strings), C, size[0], size[1])
        idx = self._indexes(shape)
        cd, cs, of = self.cdf.numpy(), self.cdf_length.numpy(), self.offset.numpy()
        out = torch.empty(shape, dtype=torch.int32)
        for b, s in enumerate(strings):
            v = coder.decode_with_indexes(s, idx[b].reshape(-1).numpy(), cd, cs, of)
            out[b] = torch.from_numpy(np.asarray(v, dtype=np.int32)).reshape(shape[1:])
        return out
