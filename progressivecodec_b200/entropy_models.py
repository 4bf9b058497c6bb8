"""Entropy models with the reference's API on the B200 kernels.

Mirrors compress/entropy_models/entropy_models.py of the reference: ``EntropyModel``,
``EntropyBottleneck`` and ``GaussianConditional`` keep the same constructor arguments, parameters,
buffers (``_offset``, ``_quantized_cdf``, ``_cdf_length``, ``scale_table`` ... — the state-dict
contract), method names and error behaviour.  Quantisation, index lookup, likelihoods and the rANS
coder run in hand-written CUDA through the C-ABI (include/pcodec_b200.h); table construction
(``update()``) is set-up work done once on the host with the library's pmf_to_quantized_cdf.

The tensor-level methods here accept the reference's NCHW tensors; the model's internal hot path uses
the NHWC-native ``*_nhwc`` helpers in ``engine.py`` and never goes through python lists.
"""
from __future__ import annotations

import ctypes as C
import warnings
from typing import Any, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import Tensor

from . import _lib as L
from . import ans as _ans
from . import tables as _tables


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _require_cuda(t: Tensor, what: str) -> None:
    if not t.is_cuda:
        raise L.PcodecError(f"{what}: expected a CUDA tensor (the B200 path has no CPU fallback)")


def pmf_to_quantized_cdf(pmf: Tensor, precision: int = 16) -> Tensor:
    """compressai._CXX.pmf_to_quantized_cdf (reference cpp_exts/ops/ops.cpp:10-67) via the C-ABI."""
    p = np.ascontiguousarray(pmf.detach().cpu().numpy().astype(np.float32))
    out = np.empty(p.size + 1, dtype=np.uint32)
    L.check(L.lib().pcodec_pmf_to_quantized_cdf(p.ctypes.data, p.size, precision, out.ctypes.data),
            "pmf_to_quantized_cdf")
    return torch.from_numpy(out.astype(np.int64)).to(torch.int32)


class LowerBound(nn.Module):
    """max(x, bound) — reference ops/bound_ops.py:44-65 (inference: plain max; carries the `bound` buffer)."""

    bound: Tensor

    def __init__(self, bound: float):
        super().__init__()
        self.register_buffer("bound", torch.Tensor([float(bound)]))

    def forward(self, x: Tensor) -> Tensor:
        return torch.max(x, self.bound)


class EntropyModel(nn.Module):
    """reference entropy_models.py:69-290."""

    def __init__(self, likelihood_bound: float = 1e-9, entropy_coder: Optional[str] = None,
                 entropy_coder_precision: int = 16):
        super().__init__()
        if entropy_coder not in (None, "ans"):
            raise ValueError(f'Unknown entropy coder "{entropy_coder}" (available: ans)')
        self.entropy_coder_precision = int(entropy_coder_precision)
        self.use_likelihood_bound = likelihood_bound > 0
        if self.use_likelihood_bound:
            self.likelihood_lower_bound = LowerBound(likelihood_bound)
        self.register_buffer("_offset", torch.IntTensor())
        self.register_buffer("_quantized_cdf", torch.IntTensor())
        self.register_buffer("_cdf_length", torch.IntTensor())
        self._tables_cache = None

    # read-only views of the table buffers (names of the reference's properties, entropy_models.py:102-120)
    offset = property(lambda self: self._offset)
    quantized_cdf = property(lambda self: self._quantized_cdf)
    cdf_length = property(lambda self: self._cdf_length)

    # -- quantisation (entropy_models.py:126-165) ------------------------------------------------------
    _MODES = ("noise", "dequantize", "symbols")

    def quantize(self, inputs: Tensor, mode: str, means: Optional[Tensor] = None) -> Tensor:
        """"symbols": int32 round(x - mean); "dequantize": round(x - mean) + mean; "noise": x + U(-1/2, 1/2) (training).
        Host-tensor API of the reference; the model path does this inside pcodec_slice_quantize."""
        if mode not in self._MODES:
            raise ValueError(f'Invalid quantization mode: "{mode}"')
        if mode == "noise":
            return inputs + torch.empty_like(inputs).uniform_(-0.5, 0.5)
        centred = inputs if means is None else inputs - means
        rounded = torch.round(centred)
        if mode == "symbols":
            return rounded.int()
        return rounded if means is None else rounded + means

    @staticmethod
    def dequantize(inputs: Tensor, means: Optional[Tensor] = None) -> Tensor:
        return inputs.float() if means is None else inputs.type_as(means) + means

    def _quantize(self, inputs, mode, means=None):
        warnings.warn("_quantize is deprecated. Use quantize instead.")
        return self.quantize(inputs, mode, means)

    @classmethod
    def _dequantize(cls, inputs, means=None):
        warnings.warn("_dequantize. Use dequantize instead.")
        return cls.dequantize(inputs, means)

    # -- tables ---------------------------------------------------------------------------------------
    def _pmf_to_cdf(self, pmf, tail_mass, pmf_length, max_length):
        """API of the reference (entropy_models.py:172-180); the work is tables.quantise_rows."""
        rows = _tables.quantise_rows(pmf, tail_mass, pmf_length, self.entropy_coder_precision, pmf_to_quantized_cdf)
        assert rows.shape[1] == max_length + 2
        return rows

    def _require_tables(self):
        """The three table buffers must be filled and well-formed before coding (the reference's _check_cdf_size /
        _check_cdf_length / _check_offsets_size, same messages: they are part of the error contract)."""
        spec = ((self._quantized_cdf, 2, "Uninitialized CDFs. Run update() first", "Invalid CDF size"),
                (self._cdf_length, 1, "Uninitialized CDF lengths. Run update() first", "Invalid offsets size"),
                (self._offset, 1, "Uninitialized offsets. Run update() first", "Invalid offsets size"))
        for buf, rank, empty_msg, shape_msg in spec:
            if buf.numel() == 0:
                raise ValueError(empty_msg)
            if buf.dim() != rank:
                raise ValueError(f"{shape_msg} {buf.size()}")

    _check_cdf_size = _check_cdf_length = _check_offsets_size = _require_tables

    def device_tables(self, device) -> _ans.CdfTables:
        """CDF tables resident on `device` (cached until update()/load_state_dict replaces the buffers)."""
        self._require_tables()
        key = (self._quantized_cdf.data_ptr(), self._quantized_cdf._version, str(device))
        if self._tables_cache is None or self._tables_cache[0] != key:
            self._tables_cache = (key, _ans.CdfTables(self._quantized_cdf, self._cdf_length.reshape(-1),
                                                      self._offset.reshape(-1), device))
        return self._tables_cache[1]

    # -- coder (entropy_models.py:203-290) --------------------------------------------------------------
    def compress(self, inputs: Tensor, indexes: Tensor, means: Optional[Tensor] = None, flag=1) -> List[bytes]:
        if len(inputs.size()) < 2:
            raise ValueError("Invalid `inputs` size. Expected a tensor with at least 2 dimensions.")
        if inputs.size() != indexes.size():
            raise ValueError("`inputs` and `indexes` should have the same size.")
        _require_cuda(inputs, "EntropyModel.compress")
        tables = self.device_tables(inputs.device)
        symbols = self.quantize(inputs, "symbols", means)
        B = symbols.size(0)
        data, offs = _ans.encode_batch(symbols.reshape(B, -1).contiguous(),
                                       indexes.int().reshape(B, -1).contiguous(), tables)
        return _ans.split_streams(data, offs)

    def decompress(self, strings, indexes: Tensor, means: Optional[Tensor] = None, flag=1) -> Tensor:
        if not isinstance(strings, (tuple, list)):
            raise ValueError("Invalid `strings` parameter type.")
        if not len(strings) == indexes.size(0):
            raise ValueError("Invalid strings or indexes parameters")
        if len(indexes.size()) < 2:
            raise ValueError("Invalid `indexes` size. Expected a tensor with at least 2 dimensions.")
        if means is not None:  # same leading [N, C]; every further dimension either matches or is 1 (broadcast)
            if means.size()[:2] != indexes.size()[:2]:
                raise ValueError("Invalid means or indexes parameters")
            if means.size() != indexes.size() and any(means.size(d) != 1 for d in range(2, indexes.dim())):
                raise ValueError("Invalid means parameters")
        _require_cuda(indexes, "EntropyModel.decompress")
        tables = self.device_tables(indexes.device)
        B = indexes.size(0)
        blob, offs = _ans.pack_streams(list(strings), indexes.device)
        sym = _ans.decode_batch(blob, offs, indexes.int().reshape(B, -1).contiguous(), tables)
        return self.dequantize(sym.reshape(indexes.size()), means)


class EntropyBottleneck(EntropyModel):
    """reference entropy_models.py:293-522."""

    _offset: Tensor

    def __init__(self, channels: int, *args: Any, tail_mass: float = 1e-9, init_scale: float = 10,
                 filters: Tuple[int, ...] = (3, 3, 3, 3), **kwargs: Any):
        super().__init__(*args, **kwargs)
        self.channels = int(channels)
        self.filters = tuple(int(f) for f in filters)
        self.init_scale = float(init_scale)
        self.tail_mass = float(tail_mass)
        # Parameters of the per-channel cumulative network (names, shapes and initial values are the state-dict
        # contract, entropy_models.py:316-348): layer i maps widths[i] -> widths[i + 1] features.
        widths = (1,) + self.filters + (1,)
        n_layers = len(widths) - 1
        per_layer = self.init_scale ** (1.0 / n_layers)
        C = self.channels

        def fresh(*shape, fill=None, uniform=None):
            t = torch.Tensor(*shape)
            if fill is not None:
                t.data.fill_(fill)
            elif uniform is not None:
                nn.init.uniform_(t, -uniform, uniform)
            return nn.Parameter(t)

        for i in range(n_layers):
            fan_out, fan_in = widths[i + 1], widths[i]
            self.register_parameter(f"_matrix{i:d}", fresh(C, fan_out, fan_in, fill=np.log(np.expm1(1 / per_layer / fan_out))))
            self.register_parameter(f"_bias{i:d}", fresh(C, fan_out, 1, uniform=0.5))
            if i < n_layers - 1:
                self.register_parameter(f"_factor{i:d}", fresh(C, fan_out, 1, fill=0.0))
        self.quantiles = nn.Parameter(torch.Tensor([-self.init_scale, 0, self.init_scale]).repeat(C, 1, 1))
        logit_tail = np.log(2 / self.tail_mass - 1)
        self.register_buffer("target", torch.Tensor([-logit_tail, 0, logit_tail]))
        self._lik_params_cache = None

    def _get_medians(self) -> Tensor:
        return self.quantiles[:, :, 1:2]

    def update(self, force: bool = False) -> bool:
        """entropy_models.py:354-393 (host set-up)."""
        if self._offset.numel() > 0 and not force:
            return False
        with torch.no_grad():
            t = _tables.bottleneck_tables(self.quantiles, lambda v: self._logits_cumulative(v, stop_gradient=True),
                                          self.entropy_coder_precision, pmf_to_quantized_cdf)
        self._quantized_cdf, self._cdf_length, self._offset = t.cdf, t.length, t.offset
        self._tables_cache = None
        return True

    def loss(self) -> Tensor:
        logits = self._logits_cumulative(self.quantiles, stop_gradient=True)
        return torch.abs(logits - self.target).sum()

    def _logits_cumulative(self, inputs: Tensor, stop_gradient: bool) -> Tensor:
        """The per-channel cumulative network of entropy_models.py:400-419: layer k is h <- softplus(M_k) @ h + b_k,
        followed (all layers but the last) by the gated non-linearity h <- h + tanh(a_k) * tanh(h).  Torch, set-up /
        training only — the hot path evaluates the same network in pcodec_bottleneck_likelihood."""
        n_hidden = len(self.filters)

        def param(kind: str, k: int) -> Tensor:
            p = getattr(self, f"_{kind}{k:d}")
            return p.detach() if stop_gradient else p

        h = inputs
        for k in range(n_hidden + 1):
            h = torch.matmul(F.softplus(param("matrix", k)), h) + param("bias", k)
            if k != n_hidden:
                h = h + torch.tanh(param("factor", k)) * torch.tanh(h)
        return h

    def likelihood_params(self, device) -> Tensor:
        """Per-channel packed parameters for pcodec_bottleneck_likelihood: [C, 58] fp32."""
        if self.filters != (3, 3, 3, 3):
            raise L.PcodecError("bottleneck likelihood kernel supports filters=(3,3,3,3)")
        key = tuple(getattr(self, f"_matrix{i}")._version for i in range(5)) + (str(device),)
        if self._lik_params_cache is None or self._lik_params_cache[0] != key:
            with torch.no_grad():
                parts = [F.softplus(getattr(self, f"_matrix{i}")).reshape(self.channels, -1) for i in range(5)]
                parts += [getattr(self, f"_bias{i}").reshape(self.channels, -1) for i in range(5)]
                parts += [torch.tanh(getattr(self, f"_factor{i}")).reshape(self.channels, -1) for i in range(4)]
                packed = torch.cat(parts, dim=1).float().contiguous().to(device)
            assert packed.shape[1] == 58
            self._lik_params_cache = (key, packed)
        return self._lik_params_cache[1]

    def forward(self, x: Tensor, training: Optional[bool] = None) -> Tuple[Tensor, Tensor]:
        """entropy_models.py:446-489 on NCHW input.  Eval mode runs on the CUDA kernels."""
        if training is None:
            training = self.training
        if training:
            raise L.PcodecError("EntropyBottleneck.forward(training=True) is outside the B200 inference hot path")
        _require_cuda(x, "EntropyBottleneck.forward")
        B, Cn = x.shape[:2]
        hw = int(np.prod(x.shape[2:]))
        xc = x.contiguous().float()
        med = self._get_medians().detach().reshape(-1).float().contiguous()
        lib = L.lib()
        # NCHW viewed as [B*C] images of hw pixels x 1 channel is not expressible with per-channel medians, so
        # transpose to NHWC first (tiny tensor).
        nhwc = torch.empty((B, hw, Cn), dtype=torch.float32, device=x.device)
        L.check(lib.pcodec_nchw_to_nhwc(xc.data_ptr(), nhwc.data_ptr(), B, Cn, hw, Cn, Cn, _stream()), "nchw_to_nhwc")
        z_hat = torch.empty_like(nhwc)
        L.check(lib.pcodec_bottleneck_quantize(nhwc.data_ptr(), Cn, med.data_ptr(), B, hw, Cn, None, None,
                                               z_hat.data_ptr(), Cn, _stream()), "bottleneck_quantize")
        lik = torch.empty_like(xc)
        L.check(lib.pcodec_bottleneck_likelihood(z_hat.data_ptr(), Cn, self.likelihood_params(x.device).data_ptr(), B,
                                                 hw, Cn, lik.data_ptr(), _stream()), "bottleneck_likelihood")
        out = torch.empty_like(xc)
        L.check(lib.pcodec_nhwc_to_nchw(z_hat.data_ptr(), Cn, out.data_ptr(), B, Cn, hw, _stream()), "nhwc_to_nchw")
        return out, lik

    @staticmethod
    def _build_indexes(size):
        """Channel number of every element of an [N, C, *spatial] tensor, int32 (entropy_models.py:491-502)."""
        n, ch, spatial = size[0], size[1], tuple(size[2:])
        per_channel = torch.arange(ch, dtype=torch.int32).reshape(1, ch, *([1] * len(spatial)))
        return per_channel.repeat(n, 1, *spatial)

    @staticmethod
    def _extend_ndims(tensor, n):
        return tensor.reshape(-1, *([1] * n)) if n > 0 else tensor.reshape(-1)

    def _medians_for(self, batch: int, n_spatial: int) -> Tensor:
        """Per-channel medians broadcast (as a view) to [batch, C, 1 x n_spatial]."""
        med = self._extend_ndims(self._get_medians().detach(), n_spatial)
        return med.expand(batch, *([-1] * (n_spatial + 1)))

    def compress(self, x):
        """entropy_models.py:508-515: code round(x - median) with one table per channel."""
        idx = self._build_indexes(x.size()).to(x.device)
        return super().compress(x, idx, self._medians_for(x.size(0), x.dim() - 2), 0)

    def decompress(self, strings, size):
        """entropy_models.py:517-522: `size` is the spatial shape; one string per batch item."""
        shape = (len(strings), self._quantized_cdf.size(0), *size)
        idx = self._build_indexes(shape).to(self.quantiles.device)
        return super().decompress(strings, idx, self._medians_for(len(strings), len(size)), 0)


class GaussianConditional(EntropyModel):
    """reference entropy_models.py:525-666."""

    def __init__(self, scale_table: Optional[Union[List, Tuple]], *args: Any, scale_bound: float = 0.11,
                 tail_mass: float = 1e-9, **kwargs: Any):
        super().__init__(*args, **kwargs)
        self._validate_scale_table(scale_table)  # error contract of entropy_models.py:541-552
        self.tail_mass = float(tail_mass)
        if scale_bound is None and scale_table:
            scale_bound = self.scale_table[0]
        if scale_bound <= 0:
            raise ValueError("Invalid parameters")
        self.lower_bound_scale = LowerBound(scale_bound)
        self.register_buffer("scale_table", self._prepare_scale_table(scale_table) if scale_table else torch.Tensor())
        self.register_buffer("scale_bound", torch.Tensor([float(scale_bound)]) if scale_bound is not None else None)

    @staticmethod
    def _validate_scale_table(table) -> None:
        if table is None:
            return
        if not isinstance(table, (list, tuple)):
            raise ValueError(f'Invalid type for scale_table "{type(table)}"')
        if len(table) < 1:
            raise ValueError(f'Invalid scale_table length "{len(table)}"')
        if any(s <= 0 for s in table) or any(b < a for a, b in zip(table, table[1:])):
            raise ValueError(f'Invalid scale_table "({table})"')

    @staticmethod
    def _prepare_scale_table(scale_table):
        return torch.Tensor(tuple(float(s) for s in scale_table))

    def _standardized_cumulative(self, inputs: Tensor) -> Tensor:
        return 0.5 * torch.erfc(float(-(2 ** -0.5)) * inputs)

    @staticmethod
    def _standardized_quantile(quantile):
        import scipy.stats

        return scipy.stats.norm.ppf(quantile)

    def update_scale_table(self, scale_table, force=False):
        if self._offset.numel() > 0 and not force:
            return False
        device = self.scale_table.device
        self.scale_table = self._prepare_scale_table(scale_table).to(device)
        self.update()
        return True

    def update(self):
        """entropy_models.py:599-624 (host set-up)."""
        with torch.no_grad():
            t = _tables.gaussian_tables(self.scale_table, self.tail_mass, self._standardized_cumulative,
                                        self.entropy_coder_precision, pmf_to_quantized_cdf)
        self._quantized_cdf, self._cdf_length, self._offset = t.cdf, t.length, t.offset
        self._tables_cache = None

    def _flat_call(self, inputs: Optional[Tensor], scales: Tensor, means: Optional[Tensor], want):
        """Run pcodec_slice_quantize on arbitrary-shaped (NCHW) tensors by viewing each batch item as
        n pixels x 1 channel (the kernel's NCHW outputs then keep the caller's element order)."""
        _require_cuda(scales, "GaussianConditional")
        B = scales.shape[0]
        n = scales[0].numel()
        sc = scales.contiguous().float()
        x = inputs.contiguous().float() if inputs is not None else None
        mu = means.expand_as(scales).contiguous().float() if means is not None else None
        table = self.scale_table.to(scales.device).float().contiguous()
        dev = scales.device
        sym = torch.empty(scales.shape, dtype=torch.int32, device=dev) if "sym" in want else None
        idx = torch.empty(scales.shape, dtype=torch.int32, device=dev) if "idx" in want else None
        lik = torch.empty(scales.shape, dtype=torch.float32, device=dev) if "lik" in want else None
        yh = torch.empty(scales.shape, dtype=torch.float32, device=dev) if "y_hat" in want else None
        ptr = lambda t: t.data_ptr() if t is not None else None
        L.check(L.lib().pcodec_slice_quantize(ptr(x), 1, None, 0, ptr(mu), 1, sc.data_ptr(), 1, B, n, 1, L.MASK_ONES,
                                              None, table.data_ptr(), table.numel(), float(self.scale_bound.item()),
                                              ptr(sym), ptr(idx), None, ptr(lik), ptr(yh), 1, _stream()),
                "slice_quantize")
        return sym, idx, lik, yh

    def _likelihood(self, inputs: Tensor, scales: Tensor, means: Optional[Tensor] = None) -> Tensor:
        """Probability mass of the unit bin around |x - mu| under N(0, max(sigma, bound)) — entropy_models.py:626-643,
        in torch (kept for API completeness; forward() uses the fused kernel)."""
        dist = torch.abs(inputs if means is None else inputs - means)
        sigma = self.lower_bound_scale(scales)
        cdf = self._standardized_cumulative
        return cdf((0.5 - dist) / sigma) - cdf((-0.5 - dist) / sigma)

    def forward(self, inputs: Tensor, scales: Tensor, means: Optional[Tensor] = None,
                training: Optional[bool] = None) -> Tuple[Tensor, Tensor]:
        """entropy_models.py:645-659; eval mode: outputs = round(x - mu) + mu, likelihood lower-bounded."""
        if training is None:
            training = self.training
        if training:
            raise L.PcodecError("GaussianConditional.forward(training=True) is outside the B200 inference hot path")
        _sym, _idx, lik, y_hat = self._flat_call(inputs, scales, means, ("lik", "y_hat"))
        if not self.use_likelihood_bound:
            raise L.PcodecError("likelihood_bound <= 0 is not supported by the fused kernel")
        return y_hat, lik

    def build_indexes(self, scales: Tensor) -> Tensor:
        """entropy_models.py:661-666."""
        _sym, idx, _lik, _y = self._flat_call(None, scales, None, ("idx",))
        return idx
