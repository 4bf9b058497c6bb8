"""REM wrapper ``PostRateProcessedNetwork`` on the B200 engine (reference: compress/models/CHProgREM.py:205-1126).

Same constructor, module tree (``base_net.*`` + ``post_latent.{level}.{slice}.*`` => identical state-dict keys) and
inference API as the reference: ``compress(x, quality, mask_pol)`` / ``decompress(strings, shape, quality, mask_pol)``.
The wrapper is the base progressive codec plus ONE extra step per progressive slice: ``apply_latent_enhancement``
(:375-431) refines sigma (and mu when ``mu_std``) with the slice's ``LatentRateReduction`` net (:12-85) of the quality
interval the level falls in, gated by the difference of the variance-aware masks at the current quality and at the
preceding check level.  It hooks into the base model's slice loops (``_base_slices(record=)``, ``_prog_slices(refine=)``)
so every conv runs through the same tap-GEMM kernels (LeakyReLU / LeakyReLU+add epilogues) and the gate is one small
fused kernel (``pcodec_masked_residual``).

``checkpoint_rep`` (the refinement nets read a previously decoded check-level representation instead of the base
slices, :773 / :989) and the ``escalation`` chain ``extract_chekpoint_representation_from_images`` (:336-372) are
supported; ``real_compress=False`` quantises without running the entropy coder (``strings`` is None, ``y_hat`` is the
latent a round trip would give).  Out of scope (raise): training-time ``forward``/``forward_latent``.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.nn as nn

from . import _lib as L
from .engine import Act, Engine, pack_conv2d
from .layers import ChannelMask, LatentRateReduction
from .models import ChannelProgresssiveWACNN


class PostRateProcessedNetwork(nn.Module):
    def __init__(self, base_net, check_levels=[0.01, 0.25, 1.75], mu_std=False, dimension="big", escalation=False):
        super().__init__()
        assert isinstance(base_net, ChannelProgresssiveWACNN)
        self.base_net = base_net
        self.mu_std = mu_std
        self.check_levels = list(check_levels)
        self.check_multiple = len(self.check_levels)
        self.escalation = escalation
        self.dimension = dimension
        self.post_latent = nn.ModuleList(
            nn.ModuleList(LatentRateReduction(dim_chunk=base_net.dim_chunk, mu_std=mu_std, dimension=dimension)
                          for _ in range(10)) for _ in range(self.check_multiple))
        self._packed = None

    # -- state ---------------------------------------------------------------------------------------------
    def load_state_dict(self, state_dict_base, state_dict_post=None, strict=False):
        """CHProgREM.py:327-335: the two parts are loaded separately."""
        self.base_net.load_state_dict(state_dict_base, strict=strict)
        if state_dict_post is not None:
            self.post_latent.load_state_dict(state_dict_post, strict=strict)
        self._packed = None

    def update(self, scale_table=None, force=False):
        return self.base_net.update(scale_table, force)

    def _apply(self, fn, *a, **k):
        self._packed = None
        return super()._apply(fn, *a, **k)

    def _prepare(self, dev):
        if self._packed is not None and self._packed["device"] == dev:
            return self._packed

        def rb(m, name):
            d = {"c1": pack_conv2d(m.conv1, dev, name + ".conv1").attach_tc(3),
                 "c2": pack_conv2d(m.conv2, dev, name + ".conv2").attach_tc(3),
                 "skip": pack_conv2d(m.skip, dev, name + ".skip").attach_tc(3) if m.skip is not None else None}
            return d

        nets = []
        for l, level in enumerate(self.post_latent):
            row = []
            for i, m in enumerate(level):
                pre = f"post_latent.{l}.{i}"
                row.append({k: [rb(b, f"{pre}.{k}.{j}") for j, b in enumerate(getattr(m, k))]
                            for k in ("enc_base_entropy_params", "enc_enh_entropy_params", "enc_base_rep", "enc")})
            nets.append(row)
        self._packed = {"device": dev, "nets": nets}
        return self._packed

    # -- CHProgREM.py:449-467 ----------------------------------------------------------------------------------
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

    # -- engine pieces -----------------------------------------------------------------------------------------
    @staticmethod
    def _rb(E: Engine, pk, segs: Sequence[Act]) -> Act:
        """ResidualBlock (models/utils.py:59-87)."""
        out = E.act(segs[0].B, segs[0].H, segs[0].W, pk["c2"].cout)
        with E.scope():
            t = E.conv_new(pk["c1"], segs, L.EPI_LEAKY)
            if pk["skip"] is not None:
                ident = E.conv_new(pk["skip"], segs)
            else:
                assert len(segs) == 1
                ident = segs[0]
            E.conv(pk["c2"], [t], out, L.EPI_LEAKY_ADD, r1=ident)
        return out

    def _seq(self, E: Engine, blocks, segs: Sequence[Act]) -> Act:
        h = self._rb(E, blocks[0], segs)
        for b in blocks[1:]:
            h = self._rb(E, b, [h])
        return h

    def _refine(self, E: Engine, quality, mask_pol, i: int, mu: Act, scale: Act, base_i: Act, record: list):
        """apply_latent_enhancement (CHProgREM.py:375-431) for progressive slice i; `record[i]` = the base slice's
        (mu, sigma).  Returns the (mu, sigma) the slice is coded with."""
        cl = self.check_levels
        if quality <= cl[0]:
            return mu, scale
        if len(cl) == 1:
            level = 0
        elif len(cl) == 2:
            level = 0 if cl[0] < quality <= cl[1] else 1
        else:
            level = 0 if cl[0] < quality <= cl[1] else (1 if cl[1] < quality <= cl[2] else 2)
        q_bar, _ = self.find_check_quality(quality)
        modes = {"ones": L.MASK_ONES, "zeros": L.MASK_ZEROS, "threshold": L.MASK_THRESHOLD}
        kind_s, qs = ChannelMask.mode_for(mask_pol, quality)
        kind_b, qb = ChannelMask.mode_for(mask_pol, q_bar)
        thr_s = E.quantile_threshold(scale, qs) if kind_s == "threshold" else None
        thr_b = E.quantile_threshold(scale, qb) if kind_b == "threshold" else None
        net = self._prepare(E.device)["nets"][level][i]
        mu_b, std_b = record[i]
        B, h, w = scale.B, scale.H, scale.W
        n_out = 64 if self.mu_std else 32
        out = E.act(B, h, w, n_out)
        with E.scope():
            f_prog = self._seq(E, net["enc_enh_entropy_params"], [mu, scale] if self.mu_std else [scale])
            f_lat = self._seq(E, net["enc_base_rep"], [base_i])
            f_base = self._seq(E, net["enc_base_entropy_params"], [mu_b, std_b])
            ret = self._seq(E, net["enc"], [f_lat, f_base, f_prog])
            if self.mu_std:
                E.masked_residual(ret.slice(0, 32), mu, scale, modes[kind_s], thr_s, modes[kind_b], thr_b, out.slice(0, 32))
                E.masked_residual(ret.slice(32, 32), scale, scale, modes[kind_s], thr_s, modes[kind_b], thr_b,
                                  out.slice(32, 32))
            else:
                E.masked_residual(ret, scale, scale, modes[kind_s], thr_s, modes[kind_b], thr_b, out)
        if self.mu_std:
            return out.slice(0, 32), out.slice(32, 32)
        return mu, out

    # -- public API ----------------------------------------------------------------------------------------------
    def forward(self, *a, **k):
        raise L.PcodecError("PostRateProcessedNetwork.forward / forward_latent are training-time paths outside the B200 "
                            "inference hot path; use compress() / decompress()")

    @torch.no_grad()
    def extract_chekpoint_representation_from_images(self, x, quality, rc=True):
        """CHProgREM.py:336-372 (the reference's spelling): the decoded latent at a check level; with ``escalation`` each
        level is coded on top of the representation of the level below."""
        if not self.escalation:
            return self.compress(x, quality=quality, mask_pol="point-based-std", real_compress=rc)["y_hat"]
        cl = self.check_levels
        out = self.compress(x, quality=cl[0], mask_pol="point-based-std", real_compress=rc)["y_hat"]
        if quality == cl[0]:
            return out
        out = self.compress(x, quality=cl[1], mask_pol="point-based-std", checkpoint_rep=out, real_compress=rc)["y_hat"]
        if quality == cl[1]:
            return out
        return self.compress(x, quality=cl[2], mask_pol="point-based-std", checkpoint_rep=out, real_compress=rc)["y_hat"]

    @torch.no_grad()
    def compress(self, x, quality=0.0, mask_pol="point-based-std", checkpoint_rep=None, real_compress=True, used_qual=None,
                 debug: Optional[dict] = None):
        """CHProgREM.py:673-887 -> {"strings", "shape", "masks", "y_hat"}."""
        self._prepare(self.base_net._device())  # pack the refinement nets on this thread, before any worker needs them
        return self.base_net.compress(x, quality=quality, mask_pol=mask_pol, debug=debug, _rem=self, _rem_ckpt=checkpoint_rep,
                                      _no_entropy=not real_compress)

    @torch.no_grad()
    def decompress(self, strings, shape, quality, mask_pol=None, checkpoint_rep=None, timing=False, used_qual=None):
        """CHProgREM.py:896-1126 -> {"x_hat", "y_hat", "time"} (y_hat: list of base slices at quality 0, else a tensor)."""
        import time as _time

        self._prepare(self.base_net._device())
        t0 = _time.time()
        out = self.base_net.decompress(strings, shape, quality, mask_pol=mask_pol, _rem=self, _rem_ckpt=checkpoint_rep)
        if timing:
            torch.cuda.synchronize()
        y_hat = list(out["y_hat"].chunk(self.base_net.ns0, 1)) if quality == 0 else out["y_hat"]
        return {"x_hat": out["x_hat"], "y_hat": y_hat, "time": (_time.time() - t0) if timing else 0}
