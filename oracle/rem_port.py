"""CPU restatement of the REM wrapper ``PostRateProcessedNetwork`` (reference: compress/models/CHProgREM.py:205-1126) on
top of oracle.codec_port.OracleCodec — test infrastructure only.

Inference paths with ``checkpoint_rep=None`` and ``escalation=False`` (what the evaluation code calls): compress()
(:673-887) and decompress() (:896-1126) are the base codec's loops with one extra step per progressive slice —
``apply_latent_enhancement`` (:375-431) refines sigma (and mu when ``mu_std``) with a per-slice
``LatentRateReduction`` net (:12-85) chosen by the quality interval [check level k, check level k+1), gated by the
difference of the variance-aware masks at the current quality and at the preceding check level.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F
from torch import Tensor

from . import entropy_port as EP
from .codec_port import CodecConfig, OracleCodec


class OracleREM(OracleCodec):
    def __init__(self, base_state: Dict[str, Tensor], post_state: Dict[str, Tensor], cfg: CodecConfig,
                 check_levels: Sequence[float] = (0.01, 0.25, 1.75), mu_std: bool = False, dimension: str = "big"):
        super().__init__(base_state, cfg)
        self.post = {k: v.detach().float() for k, v in post_state.items()}  # keys of rem.post_latent.state_dict()
        self.check_levels = list(check_levels)
        self.mu_std = mu_std
        self.n_rb = 3 if dimension == "big" else 2

    # -- LatentRateReduction ------------------------------------------------------------------------------
    def _rb(self, x: Tensor, pre: str) -> Tensor:
        """ResidualBlock (models/utils.py:59-87): conv3x3 - LeakyReLU - conv3x3 - LeakyReLU, + identity / 1x1 skip."""
        w = self.post
        out = F.leaky_relu(F.conv2d(x, w[pre + ".conv1.weight"], w[pre + ".conv1.bias"], padding=1), 0.01)
        out = F.leaky_relu(F.conv2d(out, w[pre + ".conv2.weight"], w[pre + ".conv2.bias"], padding=1), 0.01)
        ident = F.conv2d(x, w[pre + ".skip.weight"], w[pre + ".skip.bias"]) if (pre + ".skip.weight") in w else x
        return out + ident

    def _seq(self, x: Tensor, pre: str, n: int) -> Tensor:
        for j in range(n):
            x = self._rb(x, f"{pre}.{j}")
        return x

    def latent_rate_reduction(self, level: int, i: int, x_base, ep_base, ep_prog, att_mask) -> Tensor:
        pre = f"{level}.{i}"
        f_ent_prog = self._seq(ep_prog, pre + ".enc_enh_entropy_params", self.n_rb)
        f_latent = self._seq(x_base, pre + ".enc_base_rep", self.n_rb)
        f_ent_base = self._seq(ep_base, pre + ".enc_base_entropy_params", self.n_rb)
        ret = self._seq(torch.cat([f_latent, f_ent_base, f_ent_prog], 1), pre + ".enc", self.n_rb + 1)
        return ret * att_mask + ep_prog

    # -- CHProgREM.py:449-467 / 375-431 -----------------------------------------------------------------------
    def find_check_quality(self, quality):
        cl = self.check_levels
        if quality <= cl[0]:
            return 0, 0
        if len(cl) in (2, 3) and cl[0] < quality <= cl[1]:
            return cl[0], cl[1]
        if len(cl) == 2 and quality > cl[1]:
            return cl[1], 10
        if len(cl) == 3 and cl[1] < quality <= cl[2]:
            return cl[1], cl[-1]
        return cl[-1], 10

    def apply_latent_enhancement(self, i, quality, quality_bar, y_b_hat, mu_scale_base, mu_scale_enh, mu, scale, mask_pol):
        bar = self.mask(scale, quality_bar, mask_pol)
        star = self.mask(scale, quality, mask_pol)
        att = torch.round(star - bar)
        if self.mu_std:
            att = torch.cat([att, att], 1)
        cl = self.check_levels
        if quality <= cl[0]:
            return mu, scale
        if len(cl) == 1:
            level = 0
        elif len(cl) == 2:
            level = 0 if cl[0] < quality <= cl[1] else 1
        else:
            level = 0 if cl[0] < quality <= cl[1] else (1 if cl[1] < quality <= cl[2] else 2)
        enh = self.latent_rate_reduction(level, i, y_b_hat, mu_scale_base, mu_scale_enh, att)
        if self.mu_std:
            m, s = enh.chunk(2, 1)
            return m, s
        return mu, enh

    # -- shared progressive loop ----------------------------------------------------------------------------
    def _rem_prog(self, lm, ls, y_hat_base, mu_base, std_base, quality, mask_pol, code, checkpoint_rep=None):
        c = self.cfg
        d0 = c.division_dimension[0]
        y_hat_q: List[Tensor] = []
        mu_total: List[Tensor] = []
        std_total: List[Tensor] = []
        # CHProgREM.py:773 / :989: the refinement nets read the decoded check-level representation when one is given
        y_b_hats = list(checkpoint_rep.chunk(self.ns0, 1)) if checkpoint_rep is not None else y_hat_base
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
            mu_scale_base = torch.cat([mu_base[i], std_base[i]], 1)
            mu_scale_enh = torch.cat([mu, scale], 1) if self.mu_std else scale
            q_bar, _q_post = self.find_check_quality(quality)
            mu, scale = self.apply_latent_enhancement(i, quality, q_bar, y_b_hats[i], mu_scale_base, mu_scale_enh, mu,
                                                      scale, mask_pol)
            m = self.mask(scale, quality, mask_pol)
            y_hat = code(i, mu, scale, m)
            lrp = self.slice_net(torch.cat([mean_support, y_hat], 1), f"lrp_transforms_prog.{i}")
            y_hat = y_hat + 0.5 * torch.tanh(lrp)
            if not c.residual_before_lrp:
                y_hat = y_hat + y_hat_base[i]
            y_hat_q.append(y_hat)
        return y_hat_q

    def _rem_base(self, lm, ls, code):
        c = self.cfg
        d0 = c.division_dimension[0]
        y_hat_base, mu_base, std_base = [], [], []
        for i in range(self.ns0):
            sup = y_hat_base[:min(c.max_support_slices, i)]
            mean_support = torch.cat([lm[:, :d0]] + sup, 1)
            scale_support = torch.cat([ls[:, :d0]] + sup, 1)
            mu = self.slice_net(mean_support, f"cc_mean_transforms.{i}")
            scale = self.slice_net(scale_support, f"cc_scale_transforms.{i}")
            mu_base.append(mu)
            std_base.append(scale)
            y_hat = code(i, mu, scale)
            lrp = self.slice_net(torch.cat([mean_support, y_hat], 1), f"lrp_transforms.{i}")
            y_hat_base.append(y_hat + 0.5 * torch.tanh(lrp))
        return y_hat_base, mu_base, std_base

    @torch.no_grad()
    def compress(self, x: Tensor, quality=0.0, mask_pol: Optional[str] = "point-based-std", coder=None, debug=None,
                 checkpoint_rep: Optional[Tensor] = None):
        c = self.cfg
        coder = coder or EP.default_coder()
        y = self.g_a(x)
        z = self.h_a(y)
        z_sym = self.eb.symbols(z)
        z_strings = self.eb.encode(z_sym, coder)
        z_hat = self.eb.dequantize(z_sym)
        lm, ls = self.hyper_latents(z_hat, enhanced=quality > 0)
        y_slices = y.chunk(self.num_slices, 1)
        y_strings, dbg_sym, dbg_idx, masks = [], [], [], []

        def code_base(i, mu, scale):
            idx = self.gc.build_indexes(scale)
            sym = torch.round(y_slices[i] - mu).int()
            y_strings.append(self.gc.encode(sym, idx, coder))
            dbg_sym.append(sym)
            dbg_idx.append(idx)
            return sym.float() + mu

        y_hat_base, mu_base, std_base = self._rem_base(lm, ls, code_base)
        if quality <= 0:
            if debug is not None:
                debug.update(symbols=dbg_sym, indexes=dbg_idx, z_sym=z_sym)
            return {"strings": [y_strings, z_strings], "shape": z.shape[-2:], "masks": masks,
                    "y_hat": torch.cat(y_hat_base, 1)}

        def code_prog(i, mu, scale, m):
            masks.append(m)
            m = torch.round(m)
            y_slice = y_slices[self.ns0 + i]
            if c.delta_encode:
                y_slice = y_slice - y_slices[i]
            idx = self.gc.build_indexes(scale * m)
            sym = torch.round((y_slice - mu) * m).int()
            y_strings.append(self.gc.encode(sym, idx, coder))
            dbg_sym.append(sym)
            dbg_idx.append(idx)
            return sym.float() + mu

        y_hat_q = self._rem_prog(lm, ls, y_hat_base, mu_base, std_base, quality, mask_pol, code_prog, checkpoint_rep)
        if debug is not None:
            debug.update(symbols=dbg_sym, indexes=dbg_idx, z_sym=z_sym)
        return {"strings": [y_strings, z_strings], "shape": z.shape[-2:], "masks": masks, "y_hat": torch.cat(y_hat_q, 1)}

    @torch.no_grad()
    def decompress(self, strings, shape, quality, mask_pol: Optional[str] = None, coder=None,
                   checkpoint_rep: Optional[Tensor] = None):
        c = self.cfg
        coder = coder or EP.default_coder()
        mask_pol = c.mask_policy if mask_pol is None else mask_pol
        z_sym = self.eb.decode(strings[1], tuple(shape), coder)
        z_hat = self.eb.dequantize(z_sym)
        lm, ls = self.hyper_latents(z_hat, enhanced=quality > 0)
        y_strings = strings[0]

        def code_base(i, mu, scale):
            return self.gc.decode(y_strings[i], self.gc.build_indexes(scale), coder).float() + mu

        y_hat_base, mu_base, std_base = self._rem_base(lm, ls, code_base)
        if quality == 0:
            return {"x_hat": self.g_s(torch.cat(y_hat_base, 1), 0).clamp_(0, 1), "y_hat": y_hat_base}

        def code_prog(i, mu, scale, m):
            idx = self.gc.build_indexes(scale * m)
            return self.gc.decode(y_strings[self.ns0 + i], idx, coder).float() + mu

        y_hat_q = self._rem_prog(lm, ls, y_hat_base, mu_base, std_base, quality, mask_pol, code_prog, checkpoint_rep)
        y_hat_en = torch.cat(y_hat_q, 1)
        return {"x_hat": self.g_s(y_hat_en, 1).clamp_(0, 1), "y_hat": y_hat_en}
