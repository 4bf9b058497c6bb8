"""Calibrated synthetic weights.

There is no network in the build or GPU environment, so no trained checkpoint exists.  With the
reference's own Kaiming init the entropy stage is degenerate (all sigma below the 0.11 floor, every
symbol 0 — SURVEY.md §7 hard part 2), which would make both the parity tests and the benchmark
meaningless.  ``apply_synthetic_weights`` fills every *parameter* of a model (ours or the
reference's — it only relies on ``named_parameters()``) from a generator keyed by the parameter
NAME, so both sides get bit-identical weights without shipping a 600 MB state dict, and rescales a
few layers so that sigma spans the scale table and streams carry a realistic symbol mix
(mostly-zero symbols, some large ones, occasional bypass escapes).

Buffers (relative_position_index, pedestals, CDF tables, ...) are left as constructed; call
``model.update(force=True)`` afterwards to rebuild the CDF tables.
"""
from __future__ import annotations

import math
import re
import zlib

import torch


def _gen(name: str, seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    return g


# output gain of the last layer of each family (tuned on the reference, see oracle/gen_golden.py --probe)
_GAINS = [
    (r"^g_a(\.\d)?\.7\.weight$", 1.0),
    (r"^h_a\.8\.weight$", 40.0),
    (r"^cc_scale_transforms(_prog)?\.\d+\.8\.weight$", 8.0),
    (r"^cc_mean_transforms(_prog)?\.\d+\.8\.weight$", 6.0),
    (r"^lrp_transforms(_prog)?\.\d+\.8\.weight$", 25.0),
    (r"^g_s(\.\d)?\.8\.weight$", 0.25),
]
_BIAS = [
    (r"^cc_scale_transforms(_prog)?\.\d+\.8\.bias$", 0.3),
    (r"^g_s(\.\d)?\.8\.bias$", 0.5),
]


def synthetic_tensor(name: str, p: torch.Tensor, seed: int, gains=None, biases=None) -> torch.Tensor:
    g = _gen(name, seed)
    shape = tuple(p.shape)
    gains = _GAINS if gains is None else gains
    biases = _BIAS if biases is None else biases
    leaf = name.rsplit(".", 1)[-1]
    if leaf == "beta":  # GDN beta parameter (reparametrised sqrt space): beta ~ 1 +- 0.2
        v = 1.0 + 0.4 * (torch.rand(shape, generator=g) - 0.5)
        return torch.sqrt(v + 2.0 ** -36)
    if leaf == "gamma":  # GDN gamma: 0.1*I + small positive off-diagonal coupling
        c = shape[0]
        v = 0.1 * torch.eye(c) + (0.05 / c) * torch.rand(shape, generator=g)
        return torch.sqrt(v + 2.0 ** -36)
    if leaf == "relative_position_bias_table":
        return 0.5 * torch.randn(shape, generator=g)
    if leaf == "quantiles":  # EntropyBottleneck quantiles [C,1,3]: median jitter, +-(6..12) support
        c = shape[0]
        med = torch.rand(c, generator=g) - 0.5
        half = 6.0 + 6.0 * torch.rand(c, generator=g)
        return torch.stack([med - half, med, med + half], dim=-1).reshape(shape)
    if re.match(r"^_matrix\d$", leaf):
        # keep the reference's constant init (entropy_models.py:327-331) with a little jitter
        filters = (1, 3, 3, 3, 3, 1)
        i = int(leaf[-1])
        scale = 10.0 ** (1 / 5)
        init = math.log(math.expm1(1 / scale / filters[i + 1]))
        return init + 0.1 * torch.randn(shape, generator=g)
    if re.match(r"^_bias\d$", leaf):
        return torch.rand(shape, generator=g) - 0.5
    if re.match(r"^_factor\d$", leaf):
        return 0.2 * torch.randn(shape, generator=g)
    if leaf == "bias":
        v = 0.02 * torch.randn(shape, generator=g)
        for pat, b in biases:
            if re.match(pat, name):
                v = v + b
        return v
    if leaf == "weight":
        if len(shape) == 4:
            # Conv2d [Cout,Cin,k,k]; ConvTranspose2d [Cin,Cout,k,k] (g_s deconvs, stride 2: each output
            # pixel sees ~k*k/4 taps)
            is_deconv = bool(re.match(r"^g_s(\.\d)?\.(1|3|6|8)\.weight$", name))
            fan_in = shape[0] * shape[2] * shape[3] / 4.0 if is_deconv else shape[1] * shape[2] * shape[3]
            std = 1.0 / math.sqrt(fan_in)
        elif len(shape) == 2:  # Linear [out,in]
            std = 1.0 / math.sqrt(shape[1])
        else:
            std = 0.02
        for pat, gain in gains:
            if re.match(pat, name):
                std *= gain
        return std * torch.randn(shape, generator=g)
    # unknown parameter kind: small noise around its constructed value
    return p.detach().clone() + 0.01 * torch.randn(shape, generator=g)


@torch.no_grad()
def apply_synthetic_weights(model: torch.nn.Module, seed: int = 0, gains=None, biases=None) -> None:
    for name, p in model.named_parameters():
        p.copy_(synthetic_tensor(name, p, seed, gains, biases).to(p.dtype))


def synthetic_image(shape, seed: int) -> torch.Tensor:
    """Seeded smooth noise in [0,1] of shape [B,3,H,W] (SURVEY.md §8d 'Synthetic inputs'): uniform noise under a 5x5
    box filter with reflected borders.  Same generator as the golden fixtures' inputs (tests pin the equality), kept
    here so that bench.py / tools never import the checker package for their inputs."""
    g = torch.Generator().manual_seed(1000 + seed)
    noise = torch.rand(*shape, generator=g)
    padded = torch.nn.functional.pad(noise, (2, 2, 2, 2), mode="reflect")
    return torch.nn.functional.avg_pool2d(padded, kernel_size=5, stride=1).contiguous()
