"""``ChannelProgresssiveWACNN`` — the reference's progressive codec model, executed on B200 kernels.

Drop-in for compress/models/CHProg_cnn.py:30 (+ its parent compress/models/cnn.py:23): same constructor
keyword arguments, the same module tree (hence identical ``state_dict`` keys), and the same public methods
with the same return dictionaries:

    update(scale_table=None, force=False)            cnn.py:137-142
    load_state_dict(sd)                              cnn.py:195-202, base.py:62-70
    forward(x, quality, mask_pol, training)          CHProg_cnn.py:478-682
    forward_single_quality(x, quality, ...)          CHProg_cnn.py:1002-1198
    compress(x, quality, mask_pol)                   CHProg_cnn.py:686-847
    decompress(strings, shape, quality, mask_pol)    CHProg_cnn.py:849-999

Inputs/outputs at this boundary are the reference's NCHW torch tensors and python ``bytes`` strings;
inside, activations are NHWC and every arithmetic step is a launch into libpcodec_b200.so (engine.py).
Training-time behaviour (noise quantisation, learnable masks, UNet post-filter, REM wrapper) is out of
scope: those arguments raise instead of silently running something else.
"""
from __future__ import annotations

import math
import threading
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn as nn
from torch import Tensor

from . import _lib as L
from . import ans as _ans
from .engine import (Act, Engine, PackedConv, new_act, pack_conv2d, pack_deconv_merged_image, pack_deconv_phases,
                     pack_first_conv_im2col, pack_gdn, pack_linear)
from .entropy_models import EntropyBottleneck, GaussianConditional
from .layers import (GDN, ChannelMask, Win_noShift_Attention, conv, conv3x3, deconv, subpel_conv3x3)

SCALES_MIN, SCALES_MAX, SCALES_LEVELS = 0.11, 256, 64  # cnn.py:14-16 (the table update() really uses)


def get_scale_table(min=SCALES_MIN, max=SCALES_MAX, levels=SCALES_LEVELS):
    """cnn.py:19-20."""
    return torch.exp(torch.linspace(math.log(min), math.log(max), levels))


def _slice_stack(cin: int) -> nn.Sequential:
    """cc_mean / cc_scale / lrp stack (CHProg_cnn.py:165-203)."""
    return nn.Sequential(conv(cin, 224, stride=1, kernel_size=3), nn.GELU(), conv(224, 176, stride=1, kernel_size=3),
                         nn.GELU(), conv(176, 128, stride=1, kernel_size=3), nn.GELU(),
                         conv(128, 64, stride=1, kernel_size=3), nn.GELU(), conv(64, 32, stride=1, kernel_size=3))


def _hyper_synthesis(n_in: int, n_out: int) -> nn.Sequential:
    """CHProg_cnn.py:208-219 / cnn.py:69-79."""
    return nn.Sequential(conv3x3(n_in, 192), nn.GELU(), subpel_conv3x3(192, 224, 2), nn.GELU(), conv3x3(224, 256),
                         nn.GELU(), subpel_conv3x3(256, 288, 2), nn.GELU(), conv3x3(288, n_out))


def _analysis(N: int, M: int) -> nn.Sequential:
    return nn.Sequential(conv(3, N, kernel_size=5, stride=2), GDN(N), conv(N, N, kernel_size=5, stride=2), GDN(N),
                         Win_noShift_Attention(dim=N, num_heads=8, window_size=8, shift_size=4),
                         conv(N, N, kernel_size=5, stride=2), GDN(N), conv(N, M, kernel_size=5, stride=2),
                         Win_noShift_Attention(dim=M, num_heads=8, window_size=4, shift_size=2))


def _synthesis(N: int, M: int) -> nn.Sequential:
    return nn.Sequential(Win_noShift_Attention(dim=M, num_heads=8, window_size=4, shift_size=2),
                         deconv(M, N, kernel_size=5, stride=2), GDN(N, inverse=True),
                         deconv(N, N, kernel_size=5, stride=2), GDN(N, inverse=True),
                         Win_noShift_Attention(dim=N, num_heads=8, window_size=8, shift_size=4),
                         deconv(N, N, kernel_size=5, stride=2), GDN(N, inverse=True), deconv(N, 3, kernel_size=5, stride=2))


class ChannelProgresssiveWACNN(nn.Module):
    def __init__(self, N=192, M=640, division_dimension=[320, 640], dim_chunk=32, multiple_decoder=True,
                 multiple_encoder=True, multiple_hyperprior=False, mask_policy="two-levels", lmbda_list=[0.005, 0.05],
                 joiner_policy="res", support_progressive_slices=0, delta_encode=False, residual_before_lrp=False,
                 double_dim=False, support_std=False, total_mu_rep=False, all_scalable=False, u_net_post=0, **kwargs):
        super().__init__()
        assert joiner_policy in ("res", "cond", "channel_cond", "channel_res")
        if joiner_policy != "res":
            raise NotImplementedError("only joiner_policy='res' is live in the reference (SURVEY.md §3.6)")
        if u_net_post != 0:
            raise NotImplementedError("u_net_post post-filter is outside the B200 hot path (SURVEY.md §2 row 11)")
        if mask_policy not in ChannelMask.SUPPORTED:
            raise NotImplementedError(f"mask policy {mask_policy!r} is a training-time ablation (SURVEY.md §2 row 4)")
        self.N, self.M = N, M
        self.dim_chunk = dim_chunk
        self.max_support_slices = 5  # cnn.py:30
        self.lmbda_list = lmbda_list
        self.multiple_encoder, self.multiple_decoder = multiple_encoder, multiple_decoder
        self.multiple_hyperprior = multiple_hyperprior
        self.mask_policy = mask_policy
        self.num_slices = int(M // dim_chunk)
        self.double_dim = double_dim
        self.division_channel = division_dimension[0]
        self.dimensions_M = self.division_dimension = list(division_dimension)
        self.num_slices_list = [self.division_channel // dim_chunk, (M - self.division_channel) // dim_chunk]
        self.num_slice_cumulative_list = [p // dim_chunk for p in self.dimensions_M]
        self.scalable_levels = len(lmbda_list)
        self.quality_list = list(range(self.scalable_levels))
        self.total_mu_rep, self.support_std, self.all_scalable = total_mu_rep, support_std, all_scalable
        self.joiner_policy = joiner_policy
        self.support_progressive_slices = support_progressive_slices
        self.u_net_post = u_net_post
        self.delta_encode, self.residual_before_lrp = delta_encode, residual_before_lrp
        self.ns0, self.ns1 = self.num_slice_cumulative_list
        assert dim_chunk == 32 and self.dimensions_M[1] == M, "slice layout of the reference (32-channel slices)"

        d0 = self.division_dimension[0]
        delta_dim = self.division_dimension[1] - d0
        edge = support_progressive_slices + 1

        self.masking = ChannelMask(mask_policy, self.scalable_levels, dim_chunk, num_levels=self.num_slices_list[1],
                                   double_dim=double_dim)
        # transforms (cnn.py:34-79, overridden by CHProg_cnn.py:131-232 when the `multiple_*` flags are set)
        self.g_a = nn.ModuleList(_analysis(N, d0) for _ in range(2)) if multiple_encoder else _analysis(N, M)
        self.g_s = nn.ModuleList(_synthesis(N, d0) for _ in range(2)) if multiple_decoder else _synthesis(N, M)
        self.h_a = nn.Sequential(conv3x3(M, 320), nn.GELU(), conv3x3(320, 288), nn.GELU(), conv3x3(288, 256, stride=2),
                                 nn.GELU(), conv3x3(256, 224), nn.GELU(), conv3x3(224, N, stride=2))
        if multiple_hyperprior:
            self.h_mean_s = nn.ModuleList(_hyper_synthesis(N, d0) for _ in range(2))
            self.h_scale_s = nn.ModuleList(_hyper_synthesis(N, d0) for _ in range(2))
        else:
            self.h_mean_s = _hyper_synthesis(N, M)
            self.h_scale_s = _hyper_synthesis(N, M)
        self.cc_mean_transforms = nn.ModuleList(_slice_stack(d0 + 32 * min(i, 5)) for i in range(self.ns0))
        self.cc_scale_transforms = nn.ModuleList(_slice_stack(d0 + 32 * min(i, 5)) for i in range(self.ns0))
        self.lrp_transforms = nn.ModuleList(_slice_stack(d0 + 32 * min(i + 1, 6)) for i in range(self.ns0))
        n_prog = self.ns1 - self.ns0
        self.cc_mean_transforms_prog = nn.ModuleList(_slice_stack(delta_dim + 32 * min(i + 1, edge)) for i in range(n_prog))
        self.cc_scale_transforms_prog = nn.ModuleList(_slice_stack(delta_dim + 32 * min(i + 1, edge)) for i in range(n_prog))
        self.lrp_transforms_prog = nn.ModuleList(_slice_stack(delta_dim + 32 * min(i + 2, edge + 1)) for i in range(self.ns0))
        self.entropy_bottleneck = EntropyBottleneck(N)
        self.gaussian_conditional = GaussianConditional(None)
        self._packed: Optional[dict] = None
        # execution knobs (not part of the reference API): tcgen05 path on/off, TF32 products per MAC in g_s
        self.tensor_cores = True
        self.synthesis_tf32_passes = 3
        self.batch_independent_slices = True  # decode base slices 5..9 (mutually independent) as one phase
        self.decode_groups = 0          # 0 = auto (B // 4 capped at 4): image groups decoded on separate CUDA streams
        self._streams = {}               # decode-group streams per worker (created under _streams_lock)
        self._streams_lock = threading.Lock()

    # ------------------------------------------------------------------------------------------------------
    # state handling
    # ------------------------------------------------------------------------------------------------------
    def update(self, scale_table=None, force=False):
        """cnn.py:137-142 + base.py:41-60."""
        if scale_table is None:
            scale_table = get_scale_table()
        updated = self.gaussian_conditional.update_scale_table(scale_table, force=force)
        updated |= self.entropy_bottleneck.update(force=force)
        self._packed = None
        return updated

    def aux_loss(self):
        return self.entropy_bottleneck.loss()

    def load_state_dict(self, state_dict, strict: bool = True):
        """cnn.py:195-202 + base.py:62-70: CDF buffers are resized to the checkpoint's before loading."""
        for mod, name, bufs in ((self.gaussian_conditional, "gaussian_conditional",
                                 ["_quantized_cdf", "_offset", "_cdf_length", "scale_table"]),
                                (self.entropy_bottleneck, "entropy_bottleneck",
                                 ["_quantized_cdf", "_offset", "_cdf_length"])):
            for b in bufs:
                key = f"{name}.{b}"
                if key in state_dict:
                    buf = getattr(mod, b)
                    if buf.numel() == 0:
                        buf.resize_(state_dict[key].size())
        out = super().load_state_dict(state_dict, strict=strict)
        self._packed = None
        return out

    def _apply(self, fn, *a, **k):
        self._packed = None
        return super()._apply(fn, *a, **k)

    def define_quality(self, quality):
        """CHProg_cnn.py:364-374."""
        if quality is None:
            return self.quality_list
        if isinstance(quality, list):
            return quality if quality[0] == 0 else [0] + quality
        return [quality]

    # ------------------------------------------------------------------------------------------------------
    # weight packing (once per device / state)
    # ------------------------------------------------------------------------------------------------------
    def _device(self):
        return self.h_a[0].weight.device

    def prepare(self, conv_impl: int = 0) -> dict:
        dev = self._device()
        if dev.type != "cuda":
            raise L.PcodecError("ChannelProgresssiveWACNN (B200) must live on a CUDA device: there is no CPU path")
        if self._packed is not None and self._packed["device"] == dev and self._packed["impl"] == conv_impl:
            return self._packed
        if self.gaussian_conditional.scale_table.numel() == 0:
            raise ValueError("Uninitialized CDFs. Run update() first")
        P: dict = {"device": dev, "impl": conv_impl, "eng": Engine(dev, conv_impl)}

        def ru(m, name):
            return [pack_conv2d(m.conv[j], dev, f"{name}.conv.{j}") for j in (0, 2, 4)]

        def win(m: Win_noShift_Attention, name):
            att = m.conv_b[0].attn
            return {"a": [ru(m.conv_a[i], f"{name}.conv_a.{i}") for i in range(3)],
                    "qkv": pack_linear(att.qkv, dev, f"{name}.qkv"), "proj": pack_linear(att.proj, dev, f"{name}.proj"),
                    "rel": att.bias_matrix().to(dev).float().contiguous(),
                    "b": [ru(m.conv_b[i], f"{name}.conv_b.{i}") for i in (1, 2, 3)],
                    "out": pack_conv2d(m.conv_b[4], dev, f"{name}.conv_b.4"),
                    "heads": m.num_heads, "ws": m.window_size, "shift": m.shift_size}

        def analysis(seq, name):
            return {"c0": pack_first_conv_im2col(seq[0], dev, 80, f"{name}.0"), "g1": pack_gdn(seq[1], dev, f"{name}.1"),
                    "c2": pack_conv2d(seq[2], dev, f"{name}.2"), "g3": pack_gdn(seq[3], dev, f"{name}.3"),
                    "w4": win(seq[4], f"{name}.4"), "c5": pack_conv2d(seq[5], dev, f"{name}.5"),
                    "g6": pack_gdn(seq[6], dev, f"{name}.6"), "c7": pack_conv2d(seq[7], dev, f"{name}.7"),
                    "w8": win(seq[8], f"{name}.8")}

        def synthesis(seq, name):
            return {"w0": win(seq[0], f"{name}.0"), "d1": pack_deconv_phases(seq[1], dev, f"{name}.1"),
                    "g2": pack_gdn(seq[2], dev, f"{name}.2"), "d3": pack_deconv_phases(seq[3], dev, f"{name}.3"),
                    "g4": pack_gdn(seq[4], dev, f"{name}.4"), "w5": win(seq[5], f"{name}.5"),
                    "d6": pack_deconv_phases(seq[6], dev, f"{name}.6"), "g7": pack_gdn(seq[7], dev, f"{name}.7"),
                    "d8": (pack_deconv_merged_image(seq[8], dev, f"{name}.8") if self.tensor_cores
                           else pack_deconv_phases(seq[8], dev, f"{name}.8"))}

        def hyper_s(seq, name):
            return [pack_conv2d(seq[0], dev, f"{name}.0"), pack_conv2d(seq[2][0], dev, f"{name}.2.0"),
                    pack_conv2d(seq[4], dev, f"{name}.4"), pack_conv2d(seq[6][0], dev, f"{name}.6.0"),
                    pack_conv2d(seq[8], dev, f"{name}.8")]

        def stack(seq, name):
            return [pack_conv2d(seq[j], dev, f"{name}.{j}") for j in (0, 2, 4, 6, 8)]

        P["g_a"] = ([analysis(self.g_a[i], f"g_a.{i}") for i in range(2)] if self.multiple_encoder
                    else [analysis(self.g_a, "g_a")])
        P["g_s"] = ([synthesis(self.g_s[i], f"g_s.{i}") for i in range(2)] if self.multiple_decoder
                    else [synthesis(self.g_s, "g_s")])
        P["h_a"] = [pack_conv2d(self.h_a[j], dev, f"h_a.{j}") for j in (0, 2, 4, 6, 8)]
        if self.multiple_hyperprior:
            P["h_mean_s"] = [hyper_s(self.h_mean_s[i], f"h_mean_s.{i}") for i in range(2)]
            P["h_scale_s"] = [hyper_s(self.h_scale_s[i], f"h_scale_s.{i}") for i in range(2)]
        else:
            P["h_mean_s"] = [hyper_s(self.h_mean_s, "h_mean_s")]
            P["h_scale_s"] = [hyper_s(self.h_scale_s, "h_scale_s")]
        for fam in ("cc_mean_transforms", "cc_scale_transforms", "lrp_transforms", "cc_mean_transforms_prog",
                    "cc_scale_transforms_prog", "lrp_transforms_prog"):
            P[fam] = [stack(m, f"{fam}.{i}") for i, m in enumerate(getattr(self, fam))]
        if self.tensor_cores:
            def walk(o, split):
                if isinstance(o, PackedConv):
                    o.attach_tc(split)
                elif isinstance(o, dict):
                    for v in o.values():
                        walk(v, split)
                elif isinstance(o, (list, tuple)):
                    for v in o:
                        walk(v, split)
            for key, val in P.items():
                # g_s only feeds the reconstruction (PSNR): plain TF32; everything that feeds round(), the sigma
                # thresholds or the quantile ranking uses the 3xTF32 split (fp32-class accuracy)
                walk(val, self.synthesis_tf32_passes if key == "g_s" else 3)
        gc, eb = self.gaussian_conditional, self.entropy_bottleneck
        P["scale_table"] = gc.scale_table.detach().to(dev).float().contiguous()
        P["scale_bound"] = float(gc.scale_bound.item())
        P["gc_tables"] = gc.device_tables(dev)
        P["eb_tables"] = eb.device_tables(dev)
        P["medians"] = eb._get_medians().detach().reshape(-1).to(dev).float().contiguous()
        P["eb_lik"] = eb.likelihood_params(dev)
        self._packed = P
        return P

    # ------------------------------------------------------------------------------------------------------
    # network pieces on the engine
    # ------------------------------------------------------------------------------------------------------
    @staticmethod
    def _ru(E: Engine, pk, x: Act, out: Optional[Act] = None, fmt: int = 3) -> Act:
        """ResidualUnit (layers.py:39-59): 1x1 -> GELU -> 3x3 -> GELU -> 1x1 -> (+x) -> GELU."""
        fmt = E._out_fmt(pk[2], fmt)
        out = out or E.act(x.B, x.H, x.W, x.C, fmt)
        with E.scope():
            h = E.conv_new(pk[0], [x], L.EPI_GELU, fmt=2)  # (intermediates feed convolutions only: split planes)
            h = E.conv_new(pk[1], [h], L.EPI_GELU, fmt=2)
            E.conv(pk[2], [h], out, L.EPI_ADD_GELU, r1=x, fmt=fmt if out.base == 0 else 3)
        return out

    def _win(self, E: Engine, pk, x: Act, out: Optional[Act] = None) -> Act:
        """Win_noShift_Attention (layers.py:69-75)."""
        out = out or E.act(x.B, x.H, x.W, x.C)
        with E.scope():
            # the units' outputs feed convolutions and residual adds only: split planes (the epilogues read the
            # residual operand from the planes as hi + lo * 2^-11)
            a = x
            for r in pk["a"]:
                a = self._ru(E, r, a, fmt=2)
            fb = E._out_fmt(pk["proj"], 2)
            b = E.act(x.B, x.H, x.W, x.C, fb)
            with E.scope():
                qkv = E.conv_new(pk["qkv"], [x], fmt=1)  # read by the attention kernel (fp32)
                att = E.window_attention(qkv, pk["rel"], pk["heads"], pk["ws"], pk["shift"])
                E.conv(pk["proj"], [att], b, L.EPI_ADD, r1=x, fmt=fb)
            for r in pk["b"]:
                b = self._ru(E, r, b, fmt=2)
            E.conv(pk["out"], [b], out, L.EPI_GATE, r1=x, r2=a)
        return out

    def _g_a_one(self, E: Engine, pk, x: Tensor, out: Act) -> Act:
        """CHProg_cnn.py:131-144."""
        with E.scope():
            # a conv feeding a GDN writes fp32 x (the GDN's multiplier) and the planes of x*x (its operand); a GDN
            # feeding a conv is read as planes only
            h = E.im2col_first(x, 5, 2, 2, 80)
            h = E.conv_new(pk["c0"], [h], square_planes=True)
            h = E.gdn_new(pk["g1"], h, False, fmt=2)
            h = E.conv_new(pk["c2"], [h], square_planes=True)
            h = E.gdn_new(pk["g3"], h, False)
            h = self._win(E, pk["w4"], h)
            h = E.conv_new(pk["c5"], [h], square_planes=True)
            h = E.gdn_new(pk["g6"], h, False, fmt=2)
            h = E.conv_new(pk["c7"], [h])
            self._win(E, pk["w8"], h, out)
        return out

    def _g_a(self, P, x: Tensor) -> Act:
        E = P["eng"]
        B, _, H, W = x.shape
        y = E.act(B, H // 16, W // 16, self.M)
        if self.multiple_encoder:
            d0 = self.dimensions_M[0]
            self._g_a_one(E, P["g_a"][0], x, y.slice(0, d0))
            self._g_a_one(E, P["g_a"][1], x, y.slice(d0, self.M - d0))
        else:
            self._g_a_one(E, P["g_a"][0], x, y)
        return y

    def _g_s(self, P, y_hat: Act, which: int, clamp: bool) -> Tensor:
        """CHProg_cnn.py:148-161 (+ clamp_(0,1) of :909/:988 fused into the last deconv)."""
        E = P["eng"]
        pk = P["g_s"][which if self.multiple_decoder else 0]
        with E.scope():
            h = self._win(E, pk["w0"], y_hat)
            h = E.deconv_new(pk["d1"], h, square_planes=True)
            h = E.gdn_new(pk["g2"], h, True, fmt=2)
            h = E.deconv_new(pk["d3"], h, square_planes=True)
            h = E.gdn_new(pk["g4"], h, True)
            h = self._win(E, pk["w5"], h)
            h = E.deconv_new(pk["d6"], h, square_planes=True)
            h = E.gdn_new(pk["g7"], h, True, fmt=2)
            epi = L.EPI_CLAMP01 if clamp else L.EPI_LINEAR
            if isinstance(pk["d8"], PackedConv):  # merged sub-pixel phases, written straight to the NCHW image
                return E.deconv_image(pk["d8"], h, epi)
            h = E.deconv_new(pk["d8"], h, epi)
            return E.to_nchw(h)

    def _h_a(self, P, y: Act) -> Act:
        E = P["eng"]
        out = E.act(y.B, y.H // 4, y.W // 4, P["h_a"][4].cout)
        with E.scope():
            h = y
            for j, pc in enumerate(P["h_a"][:4]):
                h = E.conv_new(pc, [h], L.EPI_GELU, fmt=2)
            E.conv(P["h_a"][4], [h], out)
        return out

    @staticmethod
    def _h_s(E: Engine, pk, z_hat: Act, out: Act) -> Act:
        with E.scope():
            h = E.conv_new(pk[0], [z_hat], L.EPI_GELU, fmt=2)
            h = E.conv_shuffle_new(pk[1], h, L.EPI_GELU, fmt=2)
            h = E.conv_new(pk[2], [h], L.EPI_GELU, fmt=2)
            h = E.conv_shuffle_new(pk[3], h, L.EPI_GELU, fmt=2)
            E.conv(pk[4], [h], out)
        return out

    def _latents(self, P, z_hat: Act, enhanced: bool):
        """Which hyper-synthesis nets run: CHProg_cnn.py:404-417 / 705-715 / 856-867."""
        E = P["eng"]
        B, h, w = z_hat.B, z_hat.H * 4, z_hat.W * 4
        d0 = self.dimensions_M[0]
        if not self.multiple_hyperprior:
            lm, ls = E.act(B, h, w, self.M), E.act(B, h, w, self.M)
            self._h_s(E, P["h_mean_s"][0], z_hat, lm)
            self._h_s(E, P["h_scale_s"][0], z_hat, ls)
            return lm, ls
        ctot = 2 * d0 if enhanced else d0
        lm, ls = E.act(B, h, w, ctot), E.act(B, h, w, ctot)
        self._h_s(E, P["h_mean_s"][0], z_hat, lm.slice(0, d0))
        self._h_s(E, P["h_scale_s"][0], z_hat, ls.slice(0, d0))
        if enhanced:
            self._h_s(E, P["h_mean_s"][1], z_hat, lm.slice(d0, d0))
            self._h_s(E, P["h_scale_s"][1], z_hat, ls.slice(d0, d0))
        return lm, ls

    @staticmethod
    def _stack(E: Engine, pk, segs: Sequence[Act], out: Act, epi=L.EPI_LINEAR, r1=None, r2=None) -> Act:
        with E.scope():
            h = E.conv_new(pk[0], segs, L.EPI_GELU, fmt=2)
            for j in (1, 2, 3):
                h = E.conv_new(pk[j], [h], L.EPI_GELU, fmt=2)
            E.conv(pk[4], [h], out, epi, r1, r2)
        return out

    @staticmethod
    def _merge_segments(E: Engine, acts: Sequence[Act]) -> List[Act]:
        """Adjacent channel ranges of one tensor become one segment; fall back to a materialised concat when
        more than 4 segments remain (only the `all_scalable` + forward_single_quality bookkeeping quirk)."""
        out: List[Act] = []
        for a in acts:
            p = out[-1] if out else None
            if (p is not None and p._root is a._root and p.ps == a.ps and (p.B, p.H, p.W) == (a.B, a.H, a.W)
                    and p.c0 + p.C == a.c0):
                out[-1] = p.slice(0, p.C + a.C)  # same buffer, adjacent channel window
            else:
                out.append(a)
        if len(out) > L.MAX_SEGMENTS:
            cat = torch.cat([a.dense() for a in out], dim=-1).contiguous()
            out = [Act(cat)]
        return out

    # -- the two slice loops -----------------------------------------------------------------------------------
    def _base_slices(self, P, lm: Act, ls: Act, code, code_many=None, record: Optional[list] = None):
        """Base loop (CHProg_cnn.py:507-544 / 729-764 / 874-904).  `code(i, mu, scale, y_pre)` performs the
        quantise-or-decode step and must fill y_pre (= symbols + mu).

        Slice i is conditioned on the FIRST min(5, i) decoded slices only (`y_hat_slices[:indice]`, :731-735), so
        from slice 5 on the entropy parameters no longer depend on the previous slice: when `code_many` is given
        (decoder), slices 5..9 are coded as ONE phase — their parameter nets run back to back and all their
        streams are entropy-decoded by a single launch — which removes 4 of the 10 serial stream-decode steps."""
        E = P["eng"]
        d0 = self.dimensions_M[0]
        B, h, w = lm.B, lm.H, lm.W
        y_hat_base = E.act(B, h, w, d0)
        lm0, ls0 = lm.slice(0, d0), ls.slice(0, d0)
        mss = self.max_support_slices

        def support(i):
            k = min(mss, i)
            return [y_hat_base.slice(0, 32 * k)] if k > 0 else []

        def params(i):
            sup = support(i)
            mu, scale = E.act(B, h, w, 32), E.act(B, h, w, 32)
            self._stack(E, P["cc_mean_transforms"][i], [lm0] + sup, mu)
            self._stack(E, P["cc_scale_transforms"][i], [ls0] + sup, scale)
            if record is not None:  # REM wrapper: the base slices' (mu, sigma) feed the refinement nets
                record.append((mu, scale))
            return mu, scale

        def refine(i, y_pre):
            self._stack(E, P["lrp_transforms"][i], [lm0] + support(i) + [y_pre], y_hat_base.slice(32 * i, 32),
                        L.EPI_LRP, r1=y_pre)

        i = 0
        while i < self.ns0:
            if code_many is not None and mss >= 0 and i >= mss and self.ns0 - i > 1:
                idxs = list(range(i, self.ns0))
                ps = [params(j) for j in idxs]
                y_pres = [E.act(B, h, w, 32) for _ in idxs]
                code_many(idxs, [p[0] for p in ps], [p[1] for p in ps], y_pres)
                for j, y_pre in zip(idxs, y_pres):
                    refine(j, y_pre)
                break
            mu, scale = params(i)
            y_pre = E.act(B, h, w, 32)
            code(i, mu, scale, y_pre)
            refine(i, y_pre)
            i += 1
        return y_hat_base

    def _prog_slices(self, P, lm: Act, ls: Act, y_hat_base: Act, quality, mask_pol, code, mode: str,
                     state: Optional[dict] = None, residual_before_lrp: bool = False, deferred: Optional[list] = None,
                     cust_map: Optional[Act] = None, refine=None):
        """Progressive loop (CHProg_cnn.py:576-642 / 775-845 / 921-983 / 1091-1166).
        `code(i, mu, scale, mask_mode, thr, y_pre)`; `mode` in {"forward","fsq","codec"} selects the
        mu_total / std_total bookkeeping of that entry point (only observable with all_scalable)."""
        E = P["eng"]
        d0 = self.dimensions_M[0]
        B, h, w = lm.B, lm.H, lm.W
        n_prog = self.ns1 - self.ns0
        y_hat_q = E.act(B, h, w, 32 * n_prog)
        lm1, ls1 = lm.slice(d0, lm.C - d0), ls.slice(d0, ls.C - d0)
        state = state if state is not None else {}
        mu_total: List[Act] = state.setdefault("mu_total", [])
        std_total: List[Act] = state.setdefault("std_total", [])
        if self.all_scalable:  # per-pass pools so that consecutive list entries are adjacent channel ranges
            mu_pool = E.act(B, h, w, 32 * n_prog)
            sc_pool = E.act(B, h, w, 32 * n_prog)
        # a custom importance map replaces sigma in the mask whatever the policy (masking.py:171-194)
        kind, q = ChannelMask.mode_for("point-based-std" if cust_map is not None else mask_pol, quality)
        mask_mode = {"ones": L.MASK_ONES, "zeros": L.MASK_ZEROS, "threshold": L.MASK_THRESHOLD}[kind]
        sps = self.support_progressive_slices
        for i in range(n_prog):
            base_i = y_hat_base.slice(32 * i, 32)

            def support(vec_is_yhat: bool, vec: List[Act]) -> List[Act]:
                if i == 0 or sps == 0:
                    return [base_i]
                k = min(sps, i)
                if vec_is_yhat:
                    return [base_i, y_hat_q.slice(32 * (i - k), 32 * k)]
                return [base_i] + vec[i - k:i]

            mean_sup = self._merge_segments(E, [lm1] + support(not self.all_scalable, mu_total))
            scale_sup = self._merge_segments(E, [ls1] + support(not self.all_scalable, std_total))
            if self.all_scalable:
                mu, scale = mu_pool.slice(32 * i, 32), sc_pool.slice(32 * i, 32)
            else:
                mu, scale = E.act(B, h, w, 32), E.act(B, h, w, 32)
            self._stack(E, P["cc_mean_transforms_prog"][i], mean_sup, mu)
            self._stack(E, P["cc_scale_transforms_prog"][i], scale_sup, scale)
            if self.all_scalable:  # bookkeeping only matters when the supports read these lists
                mut = mu
                if self.total_mu_rep:
                    mut = Act((mu.dense() + base_i.dense()).contiguous())
                if mode == "forward":
                    std_total.append(scale)
                    mu_total.append(mut)
                elif mode == "fsq":
                    std_total.append(scale if self.support_std else mut)
                    mu_total.append(mut)
                    std_total.append(scale)
                else:
                    std_total.append(scale if self.support_std else mut)
                    mu_total.append(mut)
            if refine is not None:  # REM wrapper: apply_latent_enhancement (CHProgREM.py:375-431), after the bookkeeping
                mu, scale = refine(i, mu, scale, base_i)
            cm_i = cust_map.slice(32 * i, 32) if cust_map is not None else None  # cust_map.chunk(10, 1)[i], CHProg_cnn.py:721
            thr = E.quantile_threshold(cm_i if cm_i is not None else scale, q) if mask_mode == L.MASK_THRESHOLD else None
            y_pre = E.act(B, h, w, 32)
            if cm_i is not None:
                code(i, mu, scale, mask_mode, thr, y_pre, mask_src=cm_i)
            else:
                code(i, mu, scale, mask_mode, thr, y_pre)
            out_i = y_hat_q.slice(32 * i, 32)
            if residual_before_lrp:  # forward_single_quality only (CHProg_cnn.py:1153-1164)
                y_pre = Act((y_pre.dense() + base_i.dense()).contiguous())
                self._stack(E, P["lrp_transforms_prog"][i], self._merge_segments(E, mean_sup + [y_pre]), out_i,
                            L.EPI_LRP, r1=y_pre)
            elif deferred is not None:
                # all_scalable: nothing below depends on this slice's symbols, so the caller may entropy-decode every
                # slice in one launch and run the LRP stacks afterwards (progressive container, container.py)
                deferred.append((i, mean_sup + [y_pre], out_i, y_pre, base_i))
            else:
                self._stack(E, P["lrp_transforms_prog"][i], self._merge_segments(E, mean_sup + [y_pre]), out_i,
                            L.EPI_LRP, r1=y_pre, r2=base_i)
        return y_hat_q

    def _run_deferred_lrp(self, P, deferred: list) -> None:
        E = P["eng"]
        for i, segs, out_i, y_pre, base_i in deferred:  # segments are merged NOW: a materialised concat must see y_pre filled
            self._stack(E, P["lrp_transforms_prog"][i], self._merge_segments(E, segs), out_i, L.EPI_LRP, r1=y_pre, r2=base_i)

    # ------------------------------------------------------------------------------------------------------
    # public API
    # ------------------------------------------------------------------------------------------------------
    def _check_input(self, x: Tensor) -> Tensor:
        if not x.is_cuda:
            raise L.PcodecError("input must be a CUDA tensor (the B200 path has no CPU fallback)")
        if x.dim() != 4 or x.shape[1] != 3 or x.shape[2] % 64 or x.shape[3] % 64:
            raise ValueError("expected [B,3,H,W] with H and W multiples of 64 (pad as training/step.py:317-319 does)")
        return x.contiguous().float()

    def _encoder_front(self, P, x: Tensor, enhanced: bool, want_z_lik: bool, slot: int = 0):
        E = P["eng"]
        E.begin(slot)
        y = self._g_a(P, x)
        z = self._h_a(P, y)
        B, hw_z = z.B, z.H * z.W
        z_sym = torch.empty((B, z.C * hw_z), dtype=torch.int32, device=E.device)
        z_idx = torch.empty((B, z.C * hw_z), dtype=torch.int32, device=E.device)
        z_hat = E.act(B, z.H, z.W, z.C)
        E.bottleneck_quantize(z, P["medians"], z_sym, z_idx, z_hat)
        z_lik = E.bottleneck_likelihood(z_hat, P["eb_lik"]) if want_z_lik else None
        lm, ls = self._latents(P, z_hat, enhanced)
        return y, z, z_sym, z_idx, z_lik, lm, ls

    @torch.no_grad()
    def forward(self, x, quality=None, mask_pol=None, training=True):
        """CHProg_cnn.py:478-682 with training=False semantics (the reference default training=True adds
        uniform noise — a training-time behaviour outside this inference path)."""
        if training:
            raise L.PcodecError("forward(training=True) (noise quantisation) is outside the B200 inference hot path; "
                                "call with training=False as the evaluation code does")
        mask_pol = self.mask_policy if mask_pol is None else mask_pol
        qs = self.define_quality(quality)
        x = self._check_input(x)
        P = self.prepare()
        E: Engine = P["eng"]
        y, z, _zs, _zi, z_lik, lm, ls = self._encoder_front(P, x, enhanced=not (quality == 0), want_z_lik=True)
        B, h, w = y.B, y.H, y.W
        table, bound = P["scale_table"], P["scale_bound"]
        lik_base: List[Tensor] = []

        def code_base(i, mu, scale, y_pre):
            lik = torch.empty((B, 32, h, w), dtype=torch.float32, device=E.device)
            E.slice_quantize(y.slice(32 * i, 32), None, mu, scale, L.MASK_ONES, None, table, bound, None, None, None,
                             lik, y_pre)
            lik_base.append(lik)

        y_hat_base = self._base_slices(P, lm, ls, code_base)
        x_hats = [self._g_s(P, y_hat_base, 0, clamp=False).unsqueeze(0)]
        y_lik_b = torch.cat(lik_base, 1)
        y_hat_b = E.to_nchw(y_hat_base)
        lik_total, y_hat_total = [], [y_hat_b]
        state: dict = {}
        y_hat_enh = y_hat_b if len(qs) == 1 else None
        for q in qs[1:]:
            liks: List[Tensor] = []

            def code_prog(i, mu, scale, mask_mode, thr, y_pre):
                lik = torch.empty((B, 32, h, w), dtype=torch.float32, device=E.device)
                y_sub = y.slice(32 * i, 32) if self.delta_encode else None
                E.slice_quantize(y.slice(32 * (self.ns0 + i), 32), y_sub, mu, scale, mask_mode, thr, table, bound,
                                 None, None, None, lik, y_pre)
                liks.append(lik)

            y_hat_q = self._prog_slices(P, lm, ls, y_hat_base, q, mask_pol, code_prog, "forward", state)
            x_hats.append(self._g_s(P, y_hat_q, 1, clamp=False).unsqueeze(0))
            lik_total.append(torch.cat(lik_base + liks, 1).unsqueeze(0))
            y_hat_enh = E.to_nchw(y_hat_q)
            y_hat_total.append(y_hat_enh)
        y_prog_lik = torch.cat(lik_total, 0) if lik_total else torch.ones_like(y_lik_b)
        return {"x_hat": torch.cat(x_hats, 0),
                "likelihoods": {"y": y_lik_b, "y_prog": y_prog_lik, "z": z_lik},
                "y_hat": y_hat_total, "y_base": y_hat_b, "y_prog": y_hat_enh,
                "mu_base": [], "mu_prog": [], "std_base": [], "std_prog": []}

    @torch.no_grad()
    def forward_single_quality(self, x, quality, mask_pol="point-based-std", force_enhanced=False, training=False):
        """CHProg_cnn.py:1002-1198."""
        if training:
            raise L.PcodecError("forward_single_quality(training=True) is outside the B200 inference hot path")
        mask_pol = self.mask_policy if mask_pol is None else mask_pol
        x = self._check_input(x)
        P = self.prepare()
        E: Engine = P["eng"]
        enhanced = bool(force_enhanced) or not (quality == 0)
        y, z, _zs, _zi, z_lik, lm, ls = self._encoder_front(P, x, enhanced=enhanced, want_z_lik=True)
        B, h, w = y.B, y.H, y.W
        table, bound = P["scale_table"], P["scale_bound"]
        lik_base: List[Tensor] = []
        mu_b: List[Act] = []
        std_b: List[Act] = []

        def code_base(i, mu, scale, y_pre):
            lik = torch.empty((B, 32, h, w), dtype=torch.float32, device=E.device)
            E.slice_quantize(y.slice(32 * i, 32), None, mu, scale, L.MASK_ONES, None, table, bound, None, None, None,
                             lik, y_pre)
            lik_base.append(lik)
            mu_b.append(mu)
            std_b.append(scale)

        y_hat_base = self._base_slices(P, lm, ls, code_base)
        cat_nchw = lambda acts: torch.cat([E.to_nchw(a) for a in acts], 1)
        if quality == 0 and not force_enhanced:
            y_hat = E.to_nchw(y_hat_base)
            return {"x_hat": self._g_s(P, y_hat_base, 0, clamp=True),
                    "likelihoods": {"y": torch.cat(lik_base, 1), "z": z_lik},
                    "y_hat": y_hat, "y_base": y_hat, "y_prog": y_hat,
                    "mu": cat_nchw(mu_b), "mu_prog": [], "std": cat_nchw(std_b), "std_prog": []}
        liks: List[Tensor] = []
        mu_p: List[Act] = []
        std_p: List[Act] = []

        def code_prog(i, mu, scale, mask_mode, thr, y_pre):
            lik = torch.empty((B, 32, h, w), dtype=torch.float32, device=E.device)
            y_sub = y.slice(32 * i, 32) if self.delta_encode else None
            E.slice_quantize(y.slice(32 * (self.ns0 + i), 32), y_sub, mu, scale, mask_mode, thr, table, bound, None,
                             None, None, lik, y_pre)
            liks.append(lik)
            mu_p.append(mu)
            std_p.append(scale)

        y_hat_q = self._prog_slices(P, lm, ls, y_hat_base, quality, mask_pol, code_prog, "fsq",
                                    residual_before_lrp=self.residual_before_lrp)
        y_hat_p = E.to_nchw(y_hat_q)
        return {"x_hat": self._g_s(P, y_hat_q, 1, clamp=True),
                "likelihoods": {"y": torch.cat(lik_base + liks, 1), "z": z_lik},
                "y_hat": y_hat_p, "y_base": E.to_nchw(y_hat_base), "y_prog": y_hat_p,
                "mu_base": cat_nchw(mu_b), "mu": cat_nchw(mu_p), "std_base": cat_nchw(std_b), "std": cat_nchw(std_p)}

    @torch.no_grad()
    def compress(self, x, quality=0.0, mask_pol=None, cust_map=None, return_device_streams: bool = False,
                 debug: Optional[dict] = None, _rem=None, _rem_ckpt=None, _no_entropy: bool = False,
                 _planes_only: bool = False, _slot: int = 0):
        """CHProg_cnn.py:686-847.  One batched rANS launch codes every (slice, image) stream.
        `debug` (tests only) receives the device symbol / index planes [n_slices, B, 32*h*w] and z symbols.
        `_planes_only` (graphs.py) stops before the entropy coder and returns the symbol / index planes: everything up to
        there is free of host synchronisation and can be captured in a CUDA graph; `_entropy_tail` finishes the call."""
        mask_pol = self.mask_policy if mask_pol is None else mask_pol
        x = self._check_input(x)
        P = self.prepare()
        E: Engine = P["eng"]
        # `_slot` (graphs.py): engine context of this call; concurrent encoder threads use distinct slots
        y, z, z_sym, z_idx, _zl, lm, ls = self._encoder_front(P, x, enhanced=not (quality == 0), want_z_lik=False,
                                                              slot=_slot)
        B, h, w = y.B, y.H, y.W
        n = 32 * h * w
        n_slices = self.ns0 if quality <= 0 else self.ns1
        table, bound = P["scale_table"], P["scale_bound"]
        sym = torch.empty((n_slices, B, n), dtype=torch.int32, device=E.device)
        idx = torch.empty((n_slices, B, n), dtype=torch.int32, device=E.device)

        def code_base(i, mu, scale, y_pre):
            E.slice_quantize(y.slice(32 * i, 32), None, mu, scale, L.MASK_ONES, None, table, bound, sym[i], idx[i], None,
                             None, y_pre)

        record: Optional[list] = [] if _rem is not None else None
        y_hat_base = self._base_slices(P, lm, ls, code_base, record=record)
        masks: List[Tensor] = []
        y_hat_out = y_hat_base
        if quality > 0:
            def code_prog(i, mu, scale, mask_mode, thr, y_pre, mask_src=None):
                m = torch.empty((B, 32, h, w), dtype=torch.float32, device=E.device)
                y_sub = y.slice(32 * i, 32) if self.delta_encode else None
                E.slice_quantize(y.slice(32 * (self.ns0 + i), 32), y_sub, mu, scale, mask_mode, thr, table, bound,
                                 sym[self.ns0 + i], idx[self.ns0 + i], m, None, y_pre, mask_src=mask_src)
                masks.append(m)

            refine = None
            if _rem is not None:
                ck = self._checkpoint_act(E, _rem_ckpt, B, h, w)  # CHProgREM.py:773: y_b_hats = checkpoint_rep.chunk(10, 1)
                refine = lambda i, mu, scale, base_i: _rem._refine(
                    E, quality, mask_pol, i, mu, scale, base_i if ck is None else ck.slice(32 * i, 32), record)
            y_hat_out = self._prog_slices(P, lm, ls, y_hat_base, quality, mask_pol, code_prog, "codec",
                                          cust_map=self._cust_map_act(E, cust_map, B, h, w), refine=refine)
        if debug is not None:
            debug.update(symbols=sym, indexes=idx, z_symbols=z_sym, y=E.to_nchw(y), y_hat_base=E.to_nchw(y_hat_base))
        if _no_entropy:  # REM real_compress=False (CHProgREM.py:857-860): quantise only; y_hat is what a round trip gives
            return {"strings": None, "shape": torch.Size([z.H, z.W]), "masks": masks, "y_hat": E.to_nchw(y_hat_out)}
        extra = {"y_hat": E.to_nchw(y_hat_out)} if _rem is not None else {}
        planes = {"sym": sym, "idx": idx, "z_sym": z_sym, "z_idx": z_idx, "masks": masks,
                  "shape": torch.Size([z.H, z.W]), "extra": extra}
        if _planes_only:
            return planes
        return self._entropy_tail(planes, return_device_streams)

    def _entropy_tail(self, planes: dict, return_device_streams: bool):
        """Entropy-code the planes of a compress() call (one host synchronisation per table set: the stream lengths)."""
        P = self.prepare()
        sym, idx, masks, shape, extra = planes["sym"], planes["idx"], planes["masks"], planes["shape"], planes["extra"]
        n_slices, B, n = sym.shape
        z_data, z_off = _ans.encode_batch(planes["z_sym"], planes["z_idx"], P["eb_tables"])
        y_data, y_off = _ans.encode_batch(sym.reshape(n_slices * B, n), idx.reshape(n_slices * B, n), P["gc_tables"])
        if return_device_streams:
            return {"streams": (y_data, y_off, z_data, z_off), "shape": shape, "masks": masks, "batch": B, **extra}
        flat = _ans.split_streams(y_data, y_off)
        y_strings = [flat[s * B:(s + 1) * B] for s in range(n_slices)]
        return {"strings": [y_strings, _ans.split_streams(z_data, z_off)], "shape": shape, "masks": masks, **extra}

    @torch.no_grad()
    def decompress(self, strings, shape, quality, mask_pol=None, cust_map=None, _worker: int = 0, _rem=None,
                   _rem_ckpt=None):
        """CHProg_cnn.py:849-999.

        `_worker` (not part of the reference API) gives concurrent decompress() calls from different host threads their
        own engine contexts and streams (pipeline.sweep with several decode workers).

        The decoder is a strictly serial chain per image (slice i's sigma needs the decoded slice i-1), and one
        rANS stream is a serial state chain, so a batch is decoded as `decode_groups` image groups, each on its
        own CUDA stream (one host thread per group): while one group's streams are being entropy-decoded by a
        handful of warps, the tensor cores run another group's parameter networks."""
        mask_pol = self.mask_policy if mask_pol is None else mask_pol
        P = self.prepare()
        E: Engine = P["eng"]
        dev = E.device
        if isinstance(strings, dict):  # device-resident streams from compress(return_device_streams=True)
            y_data, y_off, z_data, z_off = strings["streams"]
            B = strings["batch"]
        else:
            B = len(strings[1])
            z_data, z_off = _ans.pack_streams(list(strings[1]), dev)
            y_data, y_off = _ans.pack_streams([s for sl in strings[0] for s in sl], dev)
        y_off_dev = y_off.to(dev)
        z_off_dev = z_off.to(dev)
        groups = self.decode_groups if self.decode_groups else max(1, min(4, B // 4))
        groups = max(1, min(groups, B, 7))  # engine slots: worker w owns slots 8w+1 .. 8w+7
        if groups == 1:
            # slot 1, not 0: slot 0 belongs to the encoder-side entry points, which pipeline.sweep() runs concurrently
            out = self._decompress_group(P, y_data, y_off_dev, z_data, z_off_dev, B, 0, B, shape, quality, mask_pol,
                                         slot=1 + 8 * _worker, cust_map=cust_map, rem=_rem, rem_ckpt=_rem_ckpt)
            return {"x_hat": out[0], "y_hat": out[1]} if _rem is not None else {"x_hat": out}
        from .sharding import shard_bounds

        cur = torch.cuda.current_stream(dev)
        with self._streams_lock:  # several decode workers may get here at once (pipeline.sweep, small batches)
            if len(self._streams.get(_worker, ())) < groups:
                # high priority: a group's entropy-decode launch is a handful of CTAs on its critical path; it should
                # get the next free SM ahead of the wide convolution grids of other groups / of a concurrent compress()
                self._streams[_worker] = [torch.cuda.Stream(device=dev, priority=-1) for _ in range(groups)]
            streams = self._streams[_worker]
        outs: List[Optional[Tensor]] = [None] * groups
        errs: List[Optional[BaseException]] = [None] * groups

        def work(g):
            try:
                lo, hi = shard_bounds(B, g, groups)
                st = streams[g]
                st.wait_stream(cur)
                with torch.cuda.device(dev), torch.cuda.stream(st), torch.no_grad():
                    outs[g] = self._decompress_group(P, y_data, y_off_dev, z_data, z_off_dev, B, lo, hi, shape,
                                                     quality, mask_pol, slot=g + 1 + 8 * _worker,
                                                     cust_map=cust_map[lo:hi] if cust_map is not None else None,
                                                     rem=_rem,
                                                     rem_ckpt=_rem_ckpt[lo:hi] if _rem_ckpt is not None else None)
            except BaseException as e:  # noqa: BLE001 - re-raised on the caller's thread
                errs[g] = e

        threads = [threading.Thread(target=work, args=(g,)) for g in range(groups)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        for e in errs:
            if e is not None:
                raise e
        for g in range(groups):
            cur.wait_stream(streams[g])
            for t in (outs[g] if _rem is not None else (outs[g],)):
                t.record_stream(cur)
        if _rem is not None:
            return {"x_hat": torch.cat([o[0] for o in outs], 0), "y_hat": torch.cat([o[1] for o in outs], 0)}
        return {"x_hat": torch.cat(outs, 0)}

    def _decode_base(self, P, y_data, y_off_dev, z_data, z_off_dev, B_total, lo, hi, shape, enhanced: bool, slot: int,
                     record: Optional[list] = None):
        """z + the 10 base slices of images [lo, hi) of a batch whose streams are laid out slice-major
        (stream (s, b) = s*B_total + b).  Returns (lm, ls, y_hat_base, decode_slice)."""
        E: Engine = P["eng"]
        dev = E.device
        E.begin(slot)
        B = hi - lo
        hz, wz = int(shape[0]), int(shape[1])
        Cz = self.entropy_bottleneck._quantized_cdf.size(0)
        z_idx = E.bottleneck_indexes(B, hz * wz, Cz)
        z_sym = _ans.decode_batch(z_data, z_off_dev[lo:hi + 1], z_idx, P["eb_tables"])
        z_hat = E.act(B, hz, wz, Cz)
        E.bottleneck_dequantize(z_sym, P["medians"], z_hat)
        lm, ls = self._latents(P, z_hat, enhanced=enhanced)
        h, w = 4 * hz, 4 * wz
        n = 32 * h * w
        table, bound = P["scale_table"], P["scale_bound"]
        tables = P["gc_tables"]

        def decode_slice(s, scale, mask_mode, thr, mu, y_pre, mask_src=None):
            ind = torch.empty((B, n), dtype=torch.int32, device=dev)
            E.slice_quantize(None, None, None, scale, mask_mode, thr, table, bound, None, ind, None, None, None,
                             mask_src=mask_src)
            sy = _ans.decode_batch(y_data, y_off_dev[s * B_total + lo:s * B_total + hi + 1], ind, tables)
            E.slice_dequantize(sy, mu, y_pre)

        def decode_many(slices, mus, scales, y_pres):
            k = len(slices)
            ind = torch.empty((k * B, n), dtype=torch.int32, device=dev)
            for j, scale in enumerate(scales):
                E.slice_quantize(None, None, None, scale, L.MASK_ONES, None, table, bound, None, ind[j * B:(j + 1) * B],
                                 None, None, None)
            starts = torch.cat([y_off_dev[s * B_total + lo:s * B_total + hi] for s in slices])
            ends = torch.cat([y_off_dev[s * B_total + lo + 1:s * B_total + hi + 1] for s in slices])
            sy = _ans.decode_ranges(y_data, starts, ends, ind, tables)
            for j, (mu, y_pre) in enumerate(zip(mus, y_pres)):
                E.slice_dequantize(sy[j * B:(j + 1) * B], mu, y_pre)

        y_hat_base = self._base_slices(
            P, lm, ls, lambda i, mu, scale, y_pre: decode_slice(i, scale, L.MASK_ONES, None, mu, y_pre),
            code_many=decode_many if self.batch_independent_slices else None, record=record)
        return lm, ls, y_hat_base, decode_slice

    def _cust_map_act(self, E: Engine, cust_map, B: int, h: int, w: int) -> Optional[Act]:
        """cust_map [B, 32*n_prog, h, w] (NCHW, as the reference takes it) -> NHWC activation of the current call."""
        if cust_map is None:
            return None
        n_prog = self.ns1 - self.ns0
        if tuple(cust_map.shape) != (B, 32 * n_prog, h, w):
            raise ValueError(f"cust_map must have shape {(B, 32 * n_prog, h, w)}, got {tuple(cust_map.shape)}")
        if not cust_map.is_cuda:
            raise L.PcodecError("cust_map must be a CUDA tensor")
        return E.from_nchw(cust_map)

    def _checkpoint_act(self, E: Engine, rep, B: int, h: int, w: int) -> Optional[Act]:
        """REM ``checkpoint_rep`` [B, 32*ns0, h, w] (the ``y_hat`` of a compress()/decompress() at a check level) -> NHWC."""
        if rep is None:
            return None
        if tuple(rep.shape) != (B, 32 * self.ns0, h, w):
            raise ValueError(f"checkpoint_rep must have shape {(B, 32 * self.ns0, h, w)}, got {tuple(rep.shape)}")
        if not rep.is_cuda:
            raise L.PcodecError("checkpoint_rep must be a CUDA tensor")
        return E.from_nchw(rep.float().contiguous())

    def _decompress_group(self, P, y_data, y_off_dev, z_data, z_off_dev, B_total, lo, hi, shape, quality, mask_pol,
                          slot: int = 0, cust_map=None, rem=None, rem_ckpt=None):
        """Decode images [lo, hi) of a batch whose streams are laid out slice-major: stream (s, b) = s*B_total + b."""
        record: Optional[list] = [] if rem is not None else None
        lm, ls, y_hat_base, decode_slice = self._decode_base(P, y_data, y_off_dev, z_data, z_off_dev, B_total, lo, hi,
                                                             shape, enhanced=not (quality == 0), slot=slot, record=record)
        E = P["eng"]
        if quality == 0:
            x_hat = self._g_s(P, y_hat_base, 0, clamp=True)
            return (x_hat, E.to_nchw(y_hat_base)) if rem is not None else x_hat
        refine = None
        if rem is not None:
            ck = self._checkpoint_act(E, rem_ckpt, lm.B, lm.H, lm.W)  # CHProgREM.py:989
            refine = lambda i, mu, scale, base_i: rem._refine(
                E, quality, mask_pol, i, mu, scale, base_i if ck is None else ck.slice(32 * i, 32), record)
        y_hat_q = self._prog_slices(
            P, lm, ls, y_hat_base, quality, mask_pol,
            lambda i, mu, scale, mask_mode, thr, y_pre, mask_src=None: decode_slice(self.ns0 + i, scale, mask_mode, thr, mu,
                                                                                    y_pre, mask_src),
            "codec", cust_map=self._cust_map_act(E, cust_map, lm.B, lm.H, lm.W), refine=refine)
        x_hat = self._g_s(P, y_hat_q, 1, clamp=True)
        return (x_hat, E.to_nchw(y_hat_q)) if rem is not None else x_hat
