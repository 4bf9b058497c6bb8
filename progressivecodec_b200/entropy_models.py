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

    @property
    def offset(self):
        return self._offset

    @property
    def quantized_cdf(self):
        return self._quantized_cdf

    @property
    def cdf_length(self):
        return self._cdf_length

    # -- quantisation (entropy_models.py:126-165) ------------------------------------------------------
    def quantize(self, inputs: Tensor, mode: str, means: Optional[Tensor] = None) -> Tensor:
        if mode not in ("noise", "dequantize", "symbols"):
            raise ValueError(f'Invalid quantization mode: "{mode}"')
        if mode == "noise":  # training-time path, not part of the inference hot path
            return inputs + torch.empty_like(inputs).uniform_(-0.5, 0.5)
        outputs = inputs.clone()
        if means is not None:
            outputs -= means
        outputs = torch.round(outputs)
        if mode == "dequantize":
            if means is not None:
                outputs += means
            return outputs
        return outputs.int()

    def _quantize(self, inputs, mode, means=None):
        warnings.warn("_quantize is deprecated. Use quantize instead.")
        return self.quantize(inputs, mode, means)

    @staticmethod
    def dequantize(inputs: Tensor, means: Optional[Tensor] = None) -> Tensor:
        if means is not None:
            outputs = inputs.type_as(means)
            outputs += means
        else:
            outputs = inputs.float()
        return outputs

    @classmethod
    def _dequantize(cls, inputs, means=None):
        warnings.warn("_dequantize. Use dequantize instead.")
        return cls.dequantize(inputs, means)

    # -- tables ---------------------------------------------------------------------------------------
    def _pmf_to_cdf(self, pmf, tail_mass, pmf_length, max_length):
        """entropy_models.py:172-180."""
        cdf = torch.zeros((len(pmf_length), max_length + 2), dtype=torch.int32, device=pmf.device)
        for i, p in enumerate(pmf):
            prob = torch.cat((p[: pmf_length[i]], tail_mass[i]), dim=0)
            _cdf = pmf_to_quantized_cdf(prob, self.entropy_coder_precision)
            cdf[i, : _cdf.size(0)] = _cdf
        return cdf

    def _check_cdf_size(self):
        if self._quantized_cdf.numel() == 0:
            raise ValueError("Uninitialized CDFs. Run update() first")
        if len(self._quantized_cdf.size()) != 2:
            raise ValueError(f"Invalid CDF size {self._quantized_cdf.size()}")

    def _check_offsets_size(self):
        if self._offset.numel() == 0:
            raise ValueError("Uninitialized offsets. Run update() first")
        if len(self._offset.size()) != 1:
            raise ValueError(f"Invalid offsets size {self._offset.size()}")

    def _check_cdf_length(self):
        if self._cdf_length.numel() == 0:
            raise ValueError("Uninitialized CDF lengths. Run update() first")
        if len(self._cdf_length.size()) != 1:
            raise ValueError(f"Invalid offsets size {self._cdf_length.size()}")

    def device_tables(self, device) -> _ans.CdfTables:
        """CDF tables resident on `device` (cached until update()/load_state_dict replaces the buffers)."""
        self._check_cdf_size()
        self._check_cdf_length()
        self._check_offsets_size()
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
        if means is not None:
            if means.size()[:2] != indexes.size()[:2]:
                raise ValueError("Invalid means or indexes parameters")
            if means.size() != indexes.size():
                for i in range(2, len(indexes.size())):
                    if means.size(i) != 1:
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
        filters = (1,) + self.filters + (1,)
        scale = self.init_scale ** (1 / (len(self.filters) + 1))
        channels = self.channels
        for i in range(len(self.filters) + 1):
            init = np.log(np.expm1(1 / scale / filters[i + 1]))
            matrix = torch.Tensor(channels, filters[i + 1], filters[i])
            matrix.data.fill_(init)
            self.register_parameter(f"_matrix{i:d}", nn.Parameter(matrix))
            bias = torch.Tensor(channels, filters[i + 1], 1)
            nn.init.uniform_(bias, -0.5, 0.5)
            self.register_parameter(f"_bias{i:d}", nn.Parameter(bias))
            if i < len(self.filters):
                factor = torch.Tensor(channels, filters[i + 1], 1)
                nn.init.zeros_(factor)
                self.register_parameter(f"_factor{i:d}", nn.Parameter(factor))
        self.quantiles = nn.Parameter(torch.Tensor(channels, 1, 3))
        init = torch.Tensor([-self.init_scale, 0, self.init_scale])
        self.quantiles.data = init.repeat(self.quantiles.size(0), 1, 1)
        target = np.log(2 / self.tail_mass - 1)
        self.register_buffer("target", torch.Tensor([-target, 0, target]))
        self._lik_params_cache = None

    def _get_medians(self) -> Tensor:
        return self.quantiles[:, :, 1:2]

    def update(self, force: bool = False) -> bool:
        """entropy_models.py:354-393 (host set-up)."""
        if self._offset.numel() > 0 and not force:
            return False
        with torch.no_grad():
            medians = self.quantiles[:, 0, 1]
            minima = torch.clamp(torch.ceil(medians - self.quantiles[:, 0, 0]).int(), min=0)
            maxima = torch.clamp(torch.ceil(self.quantiles[:, 0, 2] - medians).int(), min=0)
            self._offset = -minima
            pmf_start = medians - minima
            pmf_length = maxima + minima + 1
            max_length = pmf_length.max().item()
            samples = torch.arange(max_length, device=pmf_start.device)
            samples = samples[None, :] + pmf_start[:, None, None]
            lower = self._logits_cumulative(samples - 0.5, stop_gradient=True)
            upper = self._logits_cumulative(samples + 0.5, stop_gradient=True)
            sign = -torch.sign(lower + upper)
            pmf = torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower))
            pmf = pmf[:, 0, :]
            tail_mass = torch.sigmoid(lower[:, 0, :1]) + torch.sigmoid(-upper[:, 0, -1:])
            self._quantized_cdf = self._pmf_to_cdf(pmf, tail_mass, pmf_length, max_length)
            self._cdf_length = pmf_length + 2
        self._tables_cache = None
        return True

    def loss(self) -> Tensor:
        logits = self._logits_cumulative(self.quantiles, stop_gradient=True)
        return torch.abs(logits - self.target).sum()

    def _logits_cumulative(self, inputs: Tensor, stop_gradient: bool) -> Tensor:
        """entropy_models.py:400-419 (torch; set-up / training only — the hot path uses the CUDA kernel)."""
        logits = inputs
        for i in range(len(self.filters) + 1):
            matrix = getattr(self, f"_matrix{i:d}")
            bias = getattr(self, f"_bias{i:d}")
            if stop_gradient:
                matrix, bias = matrix.detach(), bias.detach()
            logits = torch.matmul(F.softplus(matrix), logits) + bias
            if i < len(self.filters):
                factor = getattr(self, f"_factor{i:d}")
                if stop_gradient:
                    factor = factor.detach()
                logits = logits + torch.tanh(factor) * torch.tanh(logits)
        return logits

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
        """entropy_models.py:491-502."""
        dims = len(size)
        N, Cn = size[0], size[1]
        view_dims = np.ones((dims,), dtype=np.int64)
        view_dims[1] = -1
        indexes = torch.arange(Cn).view(*view_dims).int()
        return indexes.repeat(N, 1, *size[2:])

    @staticmethod
    def _extend_ndims(tensor, n):
        return tensor.reshape(-1, *([1] * n)) if n > 0 else tensor.reshape(-1)

    def compress(self, x):
        """entropy_models.py:508-515."""
        indexes = self._build_indexes(x.size()).to(x.device)
        medians = self._get_medians().detach()
        spatial_dims = len(x.size()) - 2
        medians = self._extend_ndims(medians, spatial_dims)
        medians = medians.expand(x.size(0), *([-1] * (spatial_dims + 1)))
        return super().compress(x, indexes, medians, 0)

    def decompress(self, strings, size):
        """entropy_models.py:517-522."""
        output_size = (len(strings), self._quantized_cdf.size(0), *size)
        dev = self.quantiles.device
        indexes = self._build_indexes(output_size).to(dev)
        medians = self._extend_ndims(self._get_medians().detach(), len(size))
        medians = medians.expand(len(strings), *([-1] * (len(size) + 1)))
        return super().decompress(strings, indexes, medians, 0)


class GaussianConditional(EntropyModel):
    """reference entropy_models.py:525-666."""

    def __init__(self, scale_table: Optional[Union[List, Tuple]], *args: Any, scale_bound: float = 0.11,
                 tail_mass: float = 1e-9, **kwargs: Any):
        super().__init__(*args, **kwargs)
        if not isinstance(scale_table, (type(None), list, tuple)):
            raise ValueError(f'Invalid type for scale_table "{type(scale_table)}"')
        if isinstance(scale_table, (list, tuple)) and len(scale_table) < 1:
            raise ValueError(f'Invalid scale_table length "{len(scale_table)}"')
        if scale_table and (scale_table != sorted(scale_table) or any(s <= 0 for s in scale_table)):
            raise ValueError(f'Invalid scale_table "({scale_table})"')
        self.tail_mass = float(tail_mass)
        if scale_bound is None and scale_table:
            scale_bound = self.scale_table[0]
        if scale_bound <= 0:
            raise ValueError("Invalid parameters")
        self.lower_bound_scale = LowerBound(scale_bound)
        self.register_buffer("scale_table", self._prepare_scale_table(scale_table) if scale_table else torch.Tensor())
        self.register_buffer("scale_bound", torch.Tensor([float(scale_bound)]) if scale_bound is not None else None)

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
            multiplier = -self._standardized_quantile(self.tail_mass / 2)
            pmf_center = torch.ceil(self.scale_table * multiplier).int()
            pmf_length = 2 * pmf_center + 1
            max_length = torch.max(pmf_length).item()
            device = pmf_center.device
            samples = torch.abs(torch.arange(max_length, device=device).int() - pmf_center[:, None]).float()
            samples_scale = self.scale_table.unsqueeze(1).float()
            upper = self._standardized_cumulative((0.5 - samples) / samples_scale)
            lower = self._standardized_cumulative((-0.5 - samples) / samples_scale)
            pmf = upper - lower
            tail_mass = 2 * lower[:, :1]
            self._quantized_cdf = self._pmf_to_cdf(pmf, tail_mass, pmf_length, max_length)
            self._offset = -pmf_center
            self._cdf_length = pmf_length + 2
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
        """entropy_models.py:626-643 in torch (kept for API completeness; forward() uses the fused kernel)."""
        values = inputs - means if means is not None else inputs
        scales = self.lower_bound_scale(scales)
        values = torch.abs(values)
        upper = self._standardized_cumulative((0.5 - values) / scales)
        lower = self._standardized_cumulative((-0.5 - values) / scales)
        return upper - lower

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
