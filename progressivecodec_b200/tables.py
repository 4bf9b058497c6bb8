"""CDF table construction for the two entropy models (set-up time, host side).

What the reference does inside ``EntropyBottleneck.update`` / ``GaussianConditional.update`` and ``_pmf_to_cdf``
(entropy_models.py:172-180, 354-393, 599-624) is split here into three independent steps, so that the model
classes only decide WHICH densities to tabulate:

    support   integer grid of every row: ``offset[r] + j``, j = 0 .. length[r]-1 (rows padded to the longest)
    masses    probability of each grid point plus the mass left in the two tails (escape symbol)
    quantise  every row through the coder's ``pmf_to_quantized_cdf`` (C-ABI, csrc/host.cpp) into one int32 matrix

The arithmetic is the reference's, statement for statement where rounding could differ, because the tables ARE the
bit-stream contract: ``tests/test_oracle_golden.py`` pins them bit-identical to the reference's ``update()``.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable

import torch
from torch import Tensor


@dataclass
class TableSet:
    cdf: Tensor      # int32 [rows, longest + 2]
    length: Tensor   # int32 [rows]: entries of each row that are in use (= support length + 2)
    offset: Tensor   # int32 [rows]: symbol value of the first grid point


def quantise_rows(masses: Tensor, tails: Tensor, support_len: Tensor, precision: int, quantiser: Callable) -> Tensor:
    """Row r = quantiser(concat(masses[r, :support_len[r]], tails[r])) left-aligned in a zero matrix."""
    longest = int(support_len.max())
    out = torch.zeros((masses.shape[0], longest + 2), dtype=torch.int32, device=masses.device)
    for r in range(masses.shape[0]):
        row = quantiser(torch.cat((masses[r, : int(support_len[r])], tails[r]), dim=0), precision)
        out[r, : row.numel()] = row
    return out


def gaussian_tables(scale_table: Tensor, tail_mass: float, cumulative: Callable[[Tensor], Tensor], precision: int,
                    quantiser: Callable) -> TableSet:
    """One zero-mean Gaussian per scale level, tabulated on [-c, c] with c = ceil(scale * z) and z the two-sided
    tail quantile (GaussianConditional.update)."""
    import scipy.stats

    z = -scipy.stats.norm.ppf(tail_mass / 2)
    half = torch.ceil(scale_table * z).int()
    support_len = 2 * half + 1
    grid = torch.arange(int(support_len.max()), device=half.device).int() - half[:, None]
    dist = torch.abs(grid).float()                      # symmetric density: only |value| matters
    sigma = scale_table.unsqueeze(1).float()
    hi = cumulative((0.5 - dist) / sigma)
    lo = cumulative((-0.5 - dist) / sigma)
    masses = hi - lo
    tails = 2 * lo[:, :1]
    return TableSet(quantise_rows(masses, tails, support_len, precision, quantiser), support_len + 2, -half)


def bottleneck_tables(quantiles: Tensor, logits_cumulative: Callable[[Tensor], Tensor], precision: int,
                      quantiser: Callable) -> TableSet:
    """One factorised density per channel, tabulated between its learned lower / upper quantiles around the median
    (EntropyBottleneck.update).  ``quantiles`` is the [C, 1, 3] parameter (lower, median, upper)."""
    median = quantiles[:, 0, 1]
    below = torch.clamp(torch.ceil(median - quantiles[:, 0, 0]).int(), min=0)
    above = torch.clamp(torch.ceil(quantiles[:, 0, 2] - median).int(), min=0)
    support_len = above + below + 1
    first = median - below
    grid = torch.arange(int(support_len.max()), device=first.device)[None, :] + first[:, None, None]
    lo = logits_cumulative(grid - 0.5)
    hi = logits_cumulative(grid + 0.5)
    flip = -torch.sign(lo + hi)                          # evaluate both sigmoids on the side where they are accurate
    masses = torch.abs(torch.sigmoid(flip * hi) - torch.sigmoid(flip * lo))[:, 0, :]
    tails = torch.sigmoid(lo[:, 0, :1]) + torch.sigmoid(-hi[:, 0, -1:])
    return TableSet(quantise_rows(masses, tails, support_len, precision, quantiser), support_len + 2, -below)
