"""CPU ORACLE (test infrastructure — never imported by the product path).

A functional, state-dict-driven restatement in plain PyTorch-CPU fp32 of the reference's
``ChannelProgresssiveWACNN`` inference hot path (forward / forward_single_quality / compress /
decompress).  It exists to (a) be pinned against the *real* reference (tests/golden, generated
by oracle/gen_golden.py from /root/reference in the build container), (b) act as the checker
for the CUDA path on the GPU box where /root/reference is absent, and (c) be timed as the
``cpu_baseline`` ("port") in bench.py.

Every function cites the reference file:line it follows (paths under /root/reference/src/compress).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

from . import entropy_port as EP

Tensor = torch.Tensor


@dataclass
class CodecConfig:
    """Constructor flags of the reference model (models/CHProg_cnn.py:31-51)."""

    N: int = 192
    M: int = 640
    division_dimension: Sequence[int] = (320, 640)
    dim_chunk: int = 32
    multiple_decoder: bool = True
    multiple_encoder: bool = True
    multiple_hyperprior: bool = False
    mask_policy: str = "two-levels"
    lmbda_list: Sequence[float] = (0.005, 0.05)
    joiner_policy: str = "res"
    support_progressive_slices: int = 0
    delta_encode: bool = False
    residual_before_lrp: bool = False
    support_std: bool = False
    total_mu_rep: bool = False
    all_scalable: bool = False
    max_support_slices: int = 5  # models/cnn.py:30

    @staticmethod
    def authors(**kw) -> "CodecConfig":
        """Flags of the authors' runs (code name mdmh-mem5-de)."""
        base = dict(multiple_decoder=True, multiple_encoder=False, multiple_hyperprior=True,
                    delta_encode=True, support_progressive_slices=5, mask_policy="point-based-std")
        base.update(kw)
        return CodecConfig(**base)


# ----------------------------------------------------------------------------------------------
# building blocks
# ----------------------------------------------------------------------------------------------

def _nonneg(p: Tensor, minimum: float) -> Tensor:
    """NonNegativeParametrizer.forward — ops/parametrizers.py:32-49 (reparam_offset 2^-18)."""
    ped = (2.0 ** -18) ** 2
    bound = (minimum + ped) ** 0.5
    return torch.clamp_min(p, torch.tensor(bound, dtype=p.dtype)) ** 2 - torch.tensor(ped, dtype=p.dtype)


class OracleCodec:
    def __init__(self, state_dict: Dict[str, Tensor], cfg: CodecConfig):
        self.sd = {k: v.detach().cpu() for k, v in state_dict.items()}
        self.cfg = cfg
        self.ns0 = cfg.division_dimension[0] // cfg.dim_chunk
        self.ns1 = cfg.division_dimension[1] // cfg.dim_chunk
        self.num_slices = cfg.M // cfg.dim_chunk
        self.gc = EP.GaussianTables.from_state_dict(self.sd, "gaussian_conditional")
        self.eb = EP.BottleneckTables.from_state_dict(self.sd, "entropy_bottleneck")
        self._attn_mask_cache: Dict = {}

    # -- parameter access ------------------------------------------------------------------
    def p(self, name: str) -> Tensor:
        return self.sd[name]

    # -- layers ----------------------------------------------------------------------------
    def conv(self, x: Tensor, pre: str, stride: int = 1) -> Tensor:
        """nn.Conv2d with padding k//2 — models/utils.py:186-193, layers/layers.py:15-29."""
        w = self.p(pre + ".weight")
        return F.conv2d(x, w, self.p(pre + ".bias"), stride=stride, padding=w.shape[-1] // 2)

    def deconv(self, x: Tensor, pre: str) -> Tensor:
        """nn.ConvTranspose2d k5 s2 p2 op1 — models/utils.py:196-204."""
        w = self.p(pre + ".weight")
        return F.conv_transpose2d(x, w, self.p(pre + ".bias"), stride=2, padding=w.shape[-1] // 2, output_padding=1)

    def gdn(self, x: Tensor, pre: str, inverse: bool) -> Tensor:
        """GDN.forward — layers/gdn.py:50-63."""
        C = x.shape[1]
        beta = _nonneg(self.p(pre + ".beta"), 1e-6)
        gamma = _nonneg(self.p(pre + ".gamma"), 0.0).reshape(C, C, 1, 1)
        norm = F.conv2d(x * x, gamma, beta)
        norm = torch.sqrt(norm) if inverse else torch.rsqrt(norm)
        return x * norm

    def residual_unit(self, x: Tensor, pre: str) -> Tensor:
        """ResidualUnit — layers/layers.py:39-59 (1x1, GELU, 3x3, GELU, 1x1, +id, GELU)."""
        h = F.gelu(self.conv(x, pre + ".conv.0"))
        h = F.gelu(self.conv(h, pre + ".conv.2"))
        h = self.conv(h, pre + ".conv.4")
        return F.gelu(h + x)

    def _shift_mask(self, H: int, W: int, ws: int, shift: int) -> Tensor:
        """SW-MSA mask (0 / -100) — layers/win_attention.py:159-177."""
        key = (H, W, ws, shift)
        if key not in self._attn_mask_cache:
            region = torch.zeros(H, W)
            bounds_h = [(0, H - ws), (H - ws, H - shift), (H - shift, H)]
            bounds_w = [(0, W - ws), (W - ws, W - shift), (W - shift, W)]
            c = 0
            for h0, h1 in bounds_h:
                for w0, w1 in bounds_w:
                    region[h0:h1, w0:w1] = c
                    c += 1
            win = region.view(H // ws, ws, W // ws, ws).permute(0, 2, 1, 3).reshape(-1, ws * ws)
            diff = win[:, None, :] - win[:, :, None]
            self._attn_mask_cache[key] = torch.where(diff != 0, torch.tensor(-100.0), torch.tensor(0.0))
        return self._attn_mask_cache[key]

    def win_attention(self, x: Tensor, pre: str, heads: int, ws: int, shift: int) -> Tensor:
        """WinBasedAttention.forward + WindowAttention.forward — layers/win_attention.py:84-115,153-207."""
        B, C, H, W = x.shape
        t = x.permute(0, 2, 3, 1)
        if shift > 0:
            t = torch.roll(t, shifts=(-shift, -shift), dims=(1, 2))
        nh, nw = H // ws, W // ws
        t = t.reshape(B, nh, ws, nw, ws, C).permute(0, 1, 3, 2, 4, 5).reshape(B * nh * nw, ws * ws, C)
        qkv = F.linear(t, self.p(pre + ".attn.qkv.weight"), self.p(pre + ".attn.qkv.bias"))
        T = ws * ws
        hd = C // heads
        qkv = qkv.reshape(-1, T, 3, heads, hd).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0] * (hd ** -0.5), qkv[1], qkv[2]
        att = q @ k.transpose(-2, -1)
        table = self.p(pre + ".attn.relative_position_bias_table")
        ridx = self.p(pre + ".attn.relative_position_index").reshape(-1)
        att = att + table[ridx].reshape(T, T, heads).permute(2, 0, 1).unsqueeze(0)
        if shift > 0:
            m = self._shift_mask(H, W, ws, shift)
            att = (att.reshape(B, nh * nw, heads, T, T) + m[None, :, None]).reshape(-1, heads, T, T)
        att = torch.softmax(att, dim=-1)
        o = (att @ v).transpose(1, 2).reshape(-1, T, C)
        o = F.linear(o, self.p(pre + ".attn.proj.weight"), self.p(pre + ".attn.proj.bias"))
        o = o.reshape(B, nh, nw, ws, ws, C).permute(0, 1, 3, 2, 4, 5).reshape(B, H, W, C)
        if shift > 0:
            o = torch.roll(o, shifts=(shift, shift), dims=(1, 2))
        return x + o.permute(0, 3, 1, 2)

    def win_block(self, x: Tensor, pre: str, heads: int, ws: int, shift: int) -> Tensor:
        """Win_noShift_Attention.forward — layers/layers.py:69-75."""
        a = x
        for i in range(3):
            a = self.residual_unit(a, f"{pre}.conv_a.{i}")
        b = self.win_attention(x, f"{pre}.conv_b.0", heads, ws, shift)
        for i in (1, 2, 3):
            b = self.residual_unit(b, f"{pre}.conv_b.{i}")
        b = self.conv(b, f"{pre}.conv_b.4")
        return a * torch.sigmoid(b) + x

    # -- transforms --------------------------------------------------------------------------
    def g_a_one(self, x: Tensor, pre: str) -> Tensor:
        """Analysis transform — models/CHProg_cnn.py:131-144 / models/cnn.py:34-44."""
        h = self.conv(x, pre + ".0", 2)
        h = self.gdn(h, pre + ".1", False)
        h = self.conv(h, pre + ".2", 2)
        h = self.gdn(h, pre + ".3", False)
        h = self.win_block(h, pre + ".4", 8, 8, 4)
        h = self.conv(h, pre + ".5", 2)
        h = self.gdn(h, pre + ".6", False)
        h = self.conv(h, pre + ".7", 2)
        return self.win_block(h, pre + ".8", 8, 4, 2)

    def g_a(self, x: Tensor) -> Tensor:
        """models/CHProg_cnn.py:483-488."""
        if self.cfg.multiple_encoder:
            return torch.cat([self.g_a_one(x, "g_a.0"), self.g_a_one(x, "g_a.1")], dim=1)
        return self.g_a_one(x, "g_a")

    def g_s(self, y_hat: Tensor, which: int) -> Tensor:
        """Synthesis transform — models/CHProg_cnn.py:148-161."""
        pre = f"g_s.{which}" if self.cfg.multiple_decoder else "g_s"
        h = self.win_block(y_hat, pre + ".0", 8, 4, 2)
        h = self.deconv(h, pre + ".1")
        h = self.gdn(h, pre + ".2", True)
        h = self.deconv(h, pre + ".3")
        h = self.gdn(h, pre + ".4", True)
        h = self.win_block(h, pre + ".5", 8, 8, 4)
        h = self.deconv(h, pre + ".6")
        h = self.gdn(h, pre + ".7", True)
        return self.deconv(h, pre + ".8")

    def h_a(self, y: Tensor) -> Tensor:
        """models/cnn.py:57-67."""
        h = F.gelu(self.conv(y, "h_a.0"))
        h = F.gelu(self.conv(h, "h_a.2"))
        h = F.gelu(self.conv(h, "h_a.4", 2))
        h = F.gelu(self.conv(h, "h_a.6"))
        return self.conv(h, "h_a.8", 2)

    def h_s(self, z_hat: Tensor, pre: str) -> Tensor:
        """Hyper synthesis — models/CHProg_cnn.py:208-232 (subpel = conv3x3 + PixelShuffle(2), layers.py:20-24)."""
        h = F.gelu(self.conv(z_hat, pre + ".0"))
        h = F.gelu(F.pixel_shuffle(self.conv(h, pre + ".2.0"), 2))
        h = F.gelu(self.conv(h, pre + ".4"))
        h = F.gelu(F.pixel_shuffle(self.conv(h, pre + ".6.0"), 2))
        return self.conv(h, pre + ".8")

    def slice_net(self, x: Tensor, pre: str) -> Tensor:
        """cc_mean / cc_scale / lrp stack: 5x conv3x3 with GELU between — models/CHProg_cnn.py:165-203."""
        h = x
        for j in (0, 2, 4, 6):
            h = F.gelu(self.conv(h, f"{pre}.{j}"))
        return self.conv(h, f"{pre}.8")

    def hyper_latents(self, z_hat: Tensor, enhanced: bool):
        """Which hyper-synthesis nets run — models/CHProg_cnn.py:404-417, 705-715, 856-867."""
        c = self.cfg
        if not c.multiple_hyperprior:
            return self.h_s(z_hat, "h_mean_s"), self.h_s(z_hat, "h_scale_s")
        m0, s0 = self.h_s(z_hat, "h_mean_s.0"), self.h_s(z_hat, "h_scale_s.0")
        if not enhanced:
            return m0, s0
        m1, s1 = self.h_s(z_hat, "h_mean_s.1"), self.h_s(z_hat, "h_scale_s.1")
        return torch.cat([m0, m1], 1), torch.cat([s0, s1], 1)

    # -- masking -----------------------------------------------------------------------------
    def mask(self, scale: Tensor, pr: float, mask_pol: Optional[str], cust_map: Optional[Tensor] = None) -> Tensor:
        """ChannelMask.forward — layers/masking.py:163-226 (policies on the inference path)."""
        if cust_map is not None:  # masking.py:171-194: the importance map replaces sigma, whatever the policy
            if pr >= 10:
                return torch.ones_like(cust_map)
            if pr == 0:
                return torch.zeros_like(cust_map)
            q = 1.0 - pr * 0.1
            out = torch.zeros_like(cust_map)
            for j in range(cust_map.shape[0]):
                flat = cust_map[j].reshape(-1)
                out[j] = (flat >= torch.quantile(flat, q)).reshape(cust_map.shape[1:]).to(cust_map.dtype)
            return out
        if mask_pol is None:
            mask_pol = self.cfg.mask_policy
        if mask_pol is None:
            return torch.ones_like(scale)
        if mask_pol == "point-based-std":
            if pr >= 10:
                return torch.ones_like(scale)
            if pr == 0:
                return torch.zeros_like(scale)
            q = 1.0 - pr * 0.1
            out = torch.zeros_like(scale)
            for j in range(scale.shape[0]):
                flat = scale[j].reshape(-1)
                out[j] = (flat >= torch.quantile(flat, q)).reshape(scale.shape[1:]).float()
            return out
        if mask_pol == "two-levels":
            return torch.zeros_like(scale) if pr == 0 else torch.ones_like(scale)
        raise NotImplementedError(mask_pol)

    # -- support vectors -------------------------------------------------------------------
    def _determine_support(self, y_hat_base: List[Tensor], i: int, progressive: List[Tensor]) -> List[Tensor]:
        """models/CHProg_cnn.py:377-383."""
        sps = self.cfg.support_progressive_slices
        if i == 0 or sps == 0:
            return [y_hat_base[i]]
        k = min(sps, i)
        return [y_hat_base[i]] + progressive[i - k:i]

    def define_quality(self, quality):
        """models/CHProg_cnn.py:364-374."""
        if quality is None:
            return list(range(len(self.cfg.lmbda_list)))
        if isinstance(quality, list):
            return quality if quality[0] == 0 else [0] + quality
        return [quality]

    # -- z path ------------------------------------------------------------------------------
    def _z_hat_and_lik(self, z: Tensor):
        """compute_hyperprior's z part — models/CHProg_cnn.py:399-403 + entropy_models.py:446-489 (eval)."""
        med = self.eb.medians.reshape(1, -1, 1, 1)
        z_hat = torch.round(z - med) + med
        lik = self.eb.likelihood(z_hat)
        return z_hat, lik

    # ==========================================================================================
    # forward() — models/CHProg_cnn.py:478-682 (training=False semantics: dequantize, no noise)
    # ==========================================================================================
    def _base_pass(self, y_slices, latent_means, latent_scales, want_lik: bool):
        c = self.cfg
        d0 = c.division_dimension[0]
        y_hat_base, liks, mus, stds = [], [], [], []
        for i in range(self.ns0):
            sup = y_hat_base[:min(c.max_support_slices, i)]
            mean_support = torch.cat([latent_means[:, :d0]] + sup, 1)
            scale_support = torch.cat([latent_scales[:, :d0]] + sup, 1)
            mu = self.slice_net(mean_support, f"cc_mean_transforms.{i}")
            scale = self.slice_net(scale_support, f"cc_scale_transforms.{i}")
            mus.append(mu)
            stds.append(scale)
            if want_lik:
                liks.append(self.gc.likelihood(y_slices[i], scale, mu))
            y_hat = torch.round(y_slices[i] - mu) + mu
            lrp = self.slice_net(torch.cat([mean_support, y_hat], 1), f"lrp_transforms.{i}")
            y_hat = y_hat + 0.5 * torch.tanh(lrp)
            y_hat_base.append(y_hat)
        return y_hat_base, liks, mus, stds

    def _prog_pass(self, y_slices, y_hat_base, latent_means, latent_scales, q, mask_pol, want_lik: bool,
                   mu_total=None, std_total=None, mode="forward", residual_before_lrp=False):
        """One progressive pass at quality q.

        models/CHProg_cnn.py:576-642 (forward), :1091-1166 (forward_single_quality).
        `mu_total/std_total` persist across qualities in forward() (:563-564), per call elsewhere.
        `mode` selects which entry point's std_total/mu_total bookkeeping is replayed.
        """
        c = self.cfg
        d0 = c.division_dimension[0]
        mu_total = [] if mu_total is None else mu_total
        std_total = [] if std_total is None else std_total
        y_hat_q, liks, mus, stds, masks = [], [], [], [], []
        for i in range(self.ns1 - self.ns0):
            y_slice = y_slices[self.ns0 + i]
            if c.delta_encode:
                y_slice = y_slice - y_slices[i]
            sv_mean = mu_total if c.all_scalable else y_hat_q
            sv_std = std_total if c.all_scalable else y_hat_q
            mean_support = torch.cat([latent_means[:, d0:]] + self._determine_support(y_hat_base, i, sv_mean), 1)
            scale_support = torch.cat([latent_scales[:, d0:]] + self._determine_support(y_hat_base, i, sv_std), 1)
            mu = self.slice_net(mean_support, f"cc_mean_transforms_prog.{i}")
            mut = mu + y_hat_base[i] if c.total_mu_rep else mu
            scale = self.slice_net(scale_support, f"cc_scale_transforms_prog.{i}")
            if mode == "forward":            # :608-610
                std_total.append(scale)
                mu_total.append(mut)
            elif mode == "fsq":              # :1123-1129 (appends to std_total twice)
                std_total.append(scale if c.support_std else mut)
                mu_total.append(mut)
                std_total.append(scale)
            else:                            # compress/decompress :807-812, :949-953
                std_total.append(scale if c.support_std else mut)
                mu_total.append(mut)
            mus.append(mu)
            stds.append(scale)
            m = torch.round(self.mask(scale, q, mask_pol))
            masks.append(m)
            if want_lik:
                liks.append(self.gc.likelihood((y_slice - mu) * m, scale * m, None))
            y_hat = torch.round(y_slice - mu) * m + mu
            if residual_before_lrp:
                y_hat = y_hat + y_hat_base[i]
            lrp = self.slice_net(torch.cat([mean_support, y_hat], 1), f"lrp_transforms_prog.{i}")
            y_hat = y_hat + 0.5 * torch.tanh(lrp)
            if not residual_before_lrp:
                y_hat = y_hat + y_hat_base[i]  # merge "res", :385-387
            y_hat_q.append(y_hat)
        return y_hat_q, liks, mus, stds, masks

    @torch.no_grad()
    def forward(self, x: Tensor, quality=None, mask_pol: Optional[str] = None):
        c = self.cfg
        mask_pol = c.mask_policy if mask_pol is None else mask_pol
        qs = self.define_quality(quality)
        y = self.g_a(x)
        z = self.h_a(y)
        z_hat, z_lik = self._z_hat_and_lik(z)
        # compute_hyperprior(y, quality): a list or None compares != 0 -> enhanced nets run (:404-417)
        lm, ls = self.hyper_latents(z_hat, enhanced=not (quality == 0))
        y_slices = y.chunk(self.num_slices, 1)
        y_hat_base, lik_base, mu_base, std_base = self._base_pass(y_slices, lm, ls, True)
        y_hat_b = torch.cat(y_hat_base, 1)
        x_hats = [self.g_s(y_hat_b, 0).unsqueeze(0)]
        y_lik_b = torch.cat(lik_base, 1)
        lik_total, y_hat_total = [], [y_hat_b]
        mu_total: List[Tensor] = []
        std_total: List[Tensor] = []
        std_prog: List[Tensor] = []
        y_hat_enh = None
        for q in qs[1:]:
            y_hat_q, liks, _mus, stds, _ = self._prog_pass(y_slices, y_hat_base, lm, ls, q, mask_pol, True,
                                                           mu_total, std_total, mode="forward")
            std_prog += stds
            y_hat_enh = torch.cat(y_hat_q, 1)
            x_hats.append(self.g_s(y_hat_enh, 1).unsqueeze(0))
            lik_total.append(torch.cat(lik_base + liks, 1).unsqueeze(0))
            y_hat_total.append(y_hat_enh)
        y_prog_lik = torch.cat(lik_total, 0) if lik_total else torch.ones_like(y_lik_b)
        return {"x_hat": torch.cat(x_hats, 0),
                "likelihoods": {"y": y_lik_b, "y_prog": y_prog_lik, "z": z_lik},
                "y_hat": y_hat_total, "y_base": y_hat_b, "y_prog": y_hat_enh,
                "mu_base": mu_base, "mu_prog": [], "std_base": std_base, "std_prog": std_prog}

    # ==========================================================================================
    # forward_single_quality() — models/CHProg_cnn.py:1002-1198
    # ==========================================================================================
    @torch.no_grad()
    def forward_single_quality(self, x: Tensor, quality, mask_pol: Optional[str] = "point-based-std",
                               force_enhanced: bool = False):
        c = self.cfg
        mask_pol = c.mask_policy if mask_pol is None else mask_pol
        y = self.g_a(x)
        z = self.h_a(y)
        z_hat, z_lik = self._z_hat_and_lik(z)
        enhanced = force_enhanced or not (quality == 0)
        lm, ls = self.hyper_latents(z_hat, enhanced=enhanced)
        y_slices = y.chunk(self.num_slices, 1)
        y_hat_base, lik_base, mu_base, std_base = self._base_pass(y_slices, lm, ls, True)
        if quality == 0 and not force_enhanced:
            y_hat = torch.cat(y_hat_base, 1)
            x_hat = self.g_s(y_hat, 0).clamp_(0, 1)
            return {"x_hat": x_hat, "likelihoods": {"y": torch.cat(lik_base, 1), "z": z_lik},
                    "y_hat": y_hat, "y_base": y_hat, "y_prog": y_hat,
                    "mu": torch.cat(mu_base, 1), "mu_prog": [], "std": torch.cat(std_base, 1), "std_prog": []}
        y_hat_q, liks, mus, stds, _ = self._prog_pass(y_slices, y_hat_base, lm, ls, quality, mask_pol, True,
                                                      mode="fsq", residual_before_lrp=c.residual_before_lrp)
        y_hat_p = torch.cat(y_hat_q, 1)
        x_hat = self.g_s(y_hat_p, 1).clamp_(0, 1)
        return {"x_hat": x_hat, "likelihoods": {"y": torch.cat(lik_base + liks, 1), "z": z_lik},
                "y_hat": y_hat_p, "y_base": torch.cat(y_hat_base, 1), "y_prog": y_hat_p,
                "mu_base": torch.cat(mu_base, 1), "mu": torch.cat(mus, 1),
                "std_base": torch.cat(std_base, 1), "std": torch.cat(stds, 1)}

    # ==========================================================================================
    # compress() — models/CHProg_cnn.py:686-847
    # ==========================================================================================
    @torch.no_grad()
    def compress(self, x: Tensor, quality=0.0, mask_pol: Optional[str] = None, coder=None, debug: Optional[dict] = None,
                 cust_map: Optional[Tensor] = None):
        c = self.cfg
        coder = coder or EP.default_coder()
        mask_pol = c.mask_policy if mask_pol is None else mask_pol
        d0 = c.division_dimension[0]
        y = self.g_a(x)
        z = self.h_a(y)
        z_sym = self.eb.symbols(z)
        z_strings = self.eb.encode(z_sym, coder)
        z_hat = self.eb.dequantize(z_sym)
        lm, ls = self.hyper_latents(z_hat, enhanced=not (quality == 0))
        y_slices = y.chunk(self.num_slices, 1)
        y_hat_base: List[Tensor] = []
        y_strings = []
        dbg_sym, dbg_idx = [], []
        for i in range(self.ns0):
            sup = y_hat_base[:min(c.max_support_slices, i)]
            mean_support = torch.cat([lm[:, :d0]] + sup, 1)
            scale_support = torch.cat([ls[:, :d0]] + sup, 1)
            mu = self.slice_net(mean_support, f"cc_mean_transforms.{i}")
            scale = self.slice_net(scale_support, f"cc_scale_transforms.{i}")
            idx = self.gc.build_indexes(scale)
            sym = torch.round(y_slices[i] - mu).int()
            y_strings.append(self.gc.encode(sym, idx, coder))
            dbg_sym.append(sym)
            dbg_idx.append(idx)
            y_hat = sym.float() + mu
            lrp = self.slice_net(torch.cat([mean_support, y_hat], 1), f"lrp_transforms.{i}")
            y_hat_base.append(y_hat + 0.5 * torch.tanh(lrp))
        masks: List[Tensor] = []
        if debug is not None:
            debug.update(y=y, z=z, z_sym=z_sym, lm=lm, ls=ls, y_hat_base=torch.cat(y_hat_base, 1))
        if quality <= 0:
            if debug is not None:
                debug.update(symbols=dbg_sym, indexes=dbg_idx)
            return {"strings": [y_strings, z_strings], "shape": z.shape[-2:], "masks": masks}
        y_hat_q: List[Tensor] = []
        mu_total: List[Tensor] = []
        std_total: List[Tensor] = []
        for i in range(self.ns1 - self.ns0):
            y_slice = y_slices[self.ns0 + i]
            if c.delta_encode:
                y_slice = y_slice - y_slices[i]
            sv_mean = mu_total if c.all_scalable else y_hat_q
            sv_std = std_total if c.all_scalable else y_hat_q
            mean_support = torch.cat([lm[:, d0:]] + self._determine_support(y_hat_base, i, sv_mean), 1)
            scale_support = torch.cat([ls[:, d0:]] + self._determine_support(y_hat_base, i, sv_std), 1)
            mu = self.slice_net(mean_support, f"cc_mean_transforms_prog.{i}")
            mut = mu + y_hat_base[i] if c.total_mu_rep else mu
            scale = self.slice_net(scale_support, f"cc_scale_transforms_prog.{i}")
            std_total.append(scale if c.support_std else mut)
            mu_total.append(mut)
            cm = cust_map.chunk(self.ns1 - self.ns0, 1)[i] if cust_map is not None else None  # CHProg_cnn.py:721
            m = self.mask(scale, quality, mask_pol, cm)
            masks.append(m)
            m = torch.round(m)
            idx = self.gc.build_indexes(scale * m)
            sym = torch.round((y_slice - mu) * m).int()
            y_strings.append(self.gc.encode(sym, idx, coder))
            dbg_sym.append(sym)
            dbg_idx.append(idx)
            y_hat = sym.float() + mu
            lrp = self.slice_net(torch.cat([mean_support, y_hat], 1), f"lrp_transforms_prog.{i}")
            y_hat = y_hat + 0.5 * torch.tanh(lrp)
            y_hat_q.append(y_hat + y_hat_base[i])
        if debug is not None:
            debug.update(symbols=dbg_sym, indexes=dbg_idx, y_hat_prog=torch.cat(y_hat_q, 1))
        return {"strings": [y_strings, z_strings], "shape": z.shape[-2:], "masks": masks}

    # ==========================================================================================
    # decompress() — models/CHProg_cnn.py:849-999
    # ==========================================================================================
    @torch.no_grad()
    def decompress(self, strings, shape, quality, mask_pol: Optional[str] = None, coder=None,
                   cust_map: Optional[Tensor] = None):
        c = self.cfg
        coder = coder or EP.default_coder()
        mask_pol = c.mask_policy if mask_pol is None else mask_pol
        d0 = c.division_dimension[0]
        z_sym = self.eb.decode(strings[1], tuple(shape), coder)
        z_hat = self.eb.dequantize(z_sym)
        lm, ls = self.hyper_latents(z_hat, enhanced=not (quality == 0))
        y_strings = strings[0]
        y_hat_base: List[Tensor] = []
        for i in range(self.ns0):
            sup = y_hat_base[:min(c.max_support_slices, i)]
            mean_support = torch.cat([lm[:, :d0]] + sup, 1)
            scale_support = torch.cat([ls[:, :d0]] + sup, 1)
            mu = self.slice_net(mean_support, f"cc_mean_transforms.{i}")
            scale = self.slice_net(scale_support, f"cc_scale_transforms.{i}")
            idx = self.gc.build_indexes(scale)
            sym = self.gc.decode(y_strings[i], idx, coder)
            y_hat = sym.float() + mu
            lrp = self.slice_net(torch.cat([mean_support, y_hat], 1), f"lrp_transforms.{i}")
            y_hat_base.append(y_hat + 0.5 * torch.tanh(lrp))
        if quality == 0:
            return {"x_hat": self.g_s(torch.cat(y_hat_base, 1), 0).clamp_(0, 1)}
        y_hat_q: List[Tensor] = []
        mu_total: List[Tensor] = []
        std_total: List[Tensor] = []
        for i in range(self.ns1 - self.ns0):
            sv_mean = mu_total if c.all_scalable else y_hat_q
            sv_std = std_total if c.all_scalable else y_hat_q
            mean_support = torch.cat([lm[:, d0:]] + self._determine_support(y_hat_base, i, sv_mean), 1)
            scale_support = torch.cat([ls[:, d0:]] + self._determine_support(y_hat_base, i, sv_std), 1)
            mu = self.slice_net(mean_support, f"cc_mean_transforms_prog.{i}")
            mut = mu + y_hat_base[i] if c.total_mu_rep else mu
            scale = self.slice_net(scale_support, f"cc_scale_transforms_prog.{i}")
            std_total.append(scale if c.support_std else mut)
            mu_total.append(mut)
            cm = cust_map.chunk(self.ns1 - self.ns0, 1)[i] if cust_map is not None else None  # CHProg_cnn.py:850
            m = self.mask(scale, quality, mask_pol, cm)
            idx = self.gc.build_indexes(scale * m)
            sym = self.gc.decode(y_strings[self.ns0 + i], idx, coder)
            y_hat = sym.float() + mu
            lrp = self.slice_net(torch.cat([mean_support, y_hat], 1), f"lrp_transforms_prog.{i}")
            y_hat = y_hat + 0.5 * torch.tanh(lrp)
            y_hat_q.append(y_hat + y_hat_base[i])
        return {"x_hat": self.g_s(torch.cat(y_hat_q, 1), 1).clamp_(0, 1)}


def psnr(a: Tensor, b: Tensor) -> float:
    """compute_psnr — training/step.py:13-15 (max_val 1)."""
    mse = torch.mean((a.double() - b.double()) ** 2).item()
    return float("inf") if mse == 0 else -10.0 * math.log10(mse)


def bpp_from_strings(strings, num_pixels: int) -> float:
    """training/step.py:360-365."""
    total = sum(len(s) for lst in strings[0] for s in lst) + sum(len(s) for s in strings[1])
    return 8.0 * total / num_pixels


def bpp_from_likelihoods(liks: Sequence[Tensor], num_pixels: int) -> float:
    """training/step.py:178."""
    return float(sum(torch.log(l.double()).sum() for l in liks) / (-math.log(2) * num_pixels))
