"""CUDA execution engine: NHWC activations, packed weights, calls into the C-ABI.

Everything numeric here runs in libpcodec_b200.so; torch is used for device memory, streams and the
one-off weight repacking at prepare() time.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn as nn
from torch import Tensor

from . import _lib as L


def _stream():
    return torch.cuda.current_stream().cuda_stream


class Act:
    """View of `C` channels starting at `c0` of an NHWC fp32 tensor [B, H, W, Ctot]."""

    __slots__ = ("t", "B", "H", "W", "C", "c0", "ps")

    def __init__(self, t: Tensor, c0: int = 0, channels: Optional[int] = None):
        assert t.dim() == 4 and t.dtype == torch.float32 and t.is_contiguous()
        self.t = t
        self.B, self.H, self.W, self.ps = t.shape
        self.c0 = c0
        self.C = self.ps - c0 if channels is None else channels
        assert 0 <= c0 and c0 + self.C <= self.ps

    @property
    def ptr(self) -> int:
        return self.t.data_ptr() + 4 * self.c0

    def slice(self, c0: int, channels: int) -> "Act":
        return Act(self.t, self.c0 + c0, channels)

    def dense(self) -> Tensor:
        return self.t[..., self.c0:self.c0 + self.C]


def new_act(B: int, H: int, W: int, channels: int, device) -> Act:
    return Act(torch.empty((B, H, W, channels), dtype=torch.float32, device=device))


class PackedConv:
    """Weights of one tap-GEMM: fp32 [T][Cin][Cout] + bias + tap offsets."""

    __slots__ = ("w", "bias", "taps", "cin", "cout", "in_step", "out_step", "off", "name", "tc", "tc_split")

    def __init__(self, w: Tensor, bias: Optional[Tensor], taps: Sequence[Tuple[int, int]], in_step=1, out_step=1,
                 off=(0, 0), name=""):
        assert w.dim() == 3 and w.shape[0] == len(taps)
        self.w = w.contiguous()
        self.bias = bias.contiguous() if bias is not None else None
        self.taps = list(taps)
        self.cin, self.cout = w.shape[1], w.shape[2]
        self.in_step, self.out_step, self.off = in_step, out_step, off
        self.name = name
        self.tc = None       # opaque handle from pcodec_conv_tc_prepare (tcgen05 path)
        self.tc_split = 3

    def attach_tc(self, split: int = 3) -> "PackedConv":
        """Build the TF32 hi/lo K-major weights + TMA maps for the tcgen05 kernel (no-op when unsupported,
        e.g. the 3-channel output layer, which stays on the fp32 SIMT kernel)."""
        if self.tc is None and self.w.is_cuda:
            h = C.c_void_p()
            rc = L.lib().pcodec_conv_tc_prepare(self.w.data_ptr(), len(self.taps), self.cin, self.cout, C.byref(h),
                                                _stream())
            if rc == L.OK:
                self.tc = h.value
            elif rc != L.ERR_UNSUPPORTED:
                L.check(rc, f"conv_tc_prepare[{self.name}]")
        self.tc_split = split
        return self

    def __del__(self):
        try:
            if self.tc is not None:
                L.lib().pcodec_conv_tc_release(self.tc)
                self.tc = None
        except Exception:
            pass


def pack_conv2d(m: nn.Conv2d, device, name="") -> PackedConv:
    """nn.Conv2d [Cout,Cin,k,k], padding k//2 -> taps (ky-p, kx-p), W[t][ci][co]."""
    k, p, s = m.kernel_size[0], m.padding[0], m.stride[0]
    w = m.weight.detach().to(device=device, dtype=torch.float32)
    packed = w.permute(2, 3, 1, 0).reshape(k * k, w.shape[1], w.shape[0])
    taps = [(ky - p, kx - p) for ky in range(k) for kx in range(k)]
    b = m.bias.detach().to(device=device, dtype=torch.float32) if m.bias is not None else None
    return PackedConv(packed, b, taps, in_step=s, name=name)


def pack_linear(m: nn.Linear, device, name="") -> PackedConv:
    w = m.weight.detach().to(device=device, dtype=torch.float32)  # [out, in]
    b = m.bias.detach().to(device=device, dtype=torch.float32) if m.bias is not None else None
    return PackedConv(w.t().reshape(1, w.shape[1], w.shape[0]), b, [(0, 0)], name=name)


def pack_first_conv_im2col(m: nn.Conv2d, device, k_pad: int, name="") -> PackedConv:
    """First analysis conv (Cin = 3) as a 1-tap GEMM over im2col patches: row (ky*k+kx)*Cin + c."""
    w = m.weight.detach().to(device=device, dtype=torch.float32)  # [Cout, Cin, k, k]
    rows = w.permute(2, 3, 1, 0).reshape(-1, w.shape[0])
    packed = torch.zeros((1, k_pad, w.shape[0]), dtype=torch.float32, device=device)
    packed[0, : rows.shape[0]] = rows
    b = m.bias.detach().to(device=device, dtype=torch.float32)
    return PackedConv(packed, b, [(0, 0)], name=name)


def pack_deconv_phases(m: nn.ConvTranspose2d, device, name="") -> List[PackedConv]:
    """ConvTranspose2d k5 s2 p2 op1 [Cin,Cout,5,5] as 4 sub-pixel phases (py,px): output (2h+py, 2w+px) gathers
    input (h+1-a, w+1-b) with kernel tap (py+2a, px+2b)."""
    assert m.kernel_size == (5, 5) and m.stride == (2, 2) and m.padding == (2, 2) and m.output_padding == (1, 1)
    w = m.weight.detach().to(device=device, dtype=torch.float32)  # [Cin, Cout, 5, 5]
    b = m.bias.detach().to(device=device, dtype=torch.float32) if m.bias is not None else None
    phases = []
    for py in (0, 1):
        for px in (0, 1):
            taps, mats = [], []
            for a in range(3 if py == 0 else 2):
                for bb in range(3 if px == 0 else 2):
                    taps.append((1 - a, 1 - bb))
                    mats.append(w[:, :, py + 2 * a, px + 2 * bb])
            phases.append(PackedConv(torch.stack(mats, 0), b, taps, in_step=1, out_step=2, off=(py, px),
                                     name=f"{name}.phase{py}{px}"))
    return phases


def pack_gdn(g, device, name="") -> PackedConv:
    beta, gamma = g.effective()  # gamma [C_out, C_in]
    gamma = gamma.to(device=device, dtype=torch.float32)
    return PackedConv(gamma.t().reshape(1, gamma.shape[1], gamma.shape[0]),
                      beta.to(device=device, dtype=torch.float32), [(0, 0)], name=name)


class Engine:
    """Stateless helpers that launch kernels on the current stream."""

    def __init__(self, device, conv_impl: int = 0):
        self.device = device
        self.lib = L.lib()
        self.conv_impl = conv_impl

    # -- generic tap conv --------------------------------------------------------------------------------
    def conv(self, pc: PackedConv, segs: Sequence[Act], out: Act, epi: int = L.EPI_LINEAR, r1: Optional[Act] = None,
             r2: Optional[Act] = None, flags: int = 0) -> Act:
        a0 = segs[0]
        d = L.ConvDesc()
        cin = 0
        for i, s in enumerate(segs):
            assert (s.B, s.H, s.W) == (a0.B, a0.H, a0.W)
            d.seg[i].ptr = s.ptr
            d.seg[i].channels = s.C
            d.seg[i].pixel_stride = s.ps
            cin += s.C
        assert cin == pc.cin, (pc.name, cin, pc.cin)
        d.n_segments = len(segs)
        d.batch, d.in_h, d.in_w = a0.B, a0.H, a0.W
        d.n_taps = len(pc.taps)
        for t, (dy, dx) in enumerate(pc.taps):
            d.dy[t] = dy
            d.dx[t] = dx
        d.in_step = pc.in_step
        d.weight = pc.w.data_ptr()
        d.bias = pc.bias.data_ptr() if pc.bias is not None else None
        d.cin_total, d.cout = pc.cin, pc.cout
        shuffle = bool(flags & L.FLAG_PIXEL_SHUFFLE2)
        if pc.out_step == 1:
            gh, gw = a0.H // pc.in_step, a0.W // pc.in_step
            oh, ow = gh, gw
        else:
            gh, gw = a0.H, a0.W
            oh, ow = a0.H * pc.out_step, a0.W * pc.out_step
        d.grid_h, d.grid_w = gh, gw
        d.out_step, d.out_off_y, d.out_off_x = pc.out_step, pc.off[0], pc.off[1]
        if shuffle:
            assert (out.H, out.W, out.C) == (2 * oh, 2 * ow, pc.cout // 4), (pc.name, out.H, out.W, out.C)
            d.out_h, d.out_w = 2 * oh, 2 * ow   # kernel indexes the shuffled tensor with these
            # for the shuffled store the kernel computes (2*oh+si, 2*ow+sj) inside an [out_h, out_w] image
        else:
            assert (out.H, out.W, out.C) == (oh, ow, pc.cout), (pc.name, (out.H, out.W, out.C), (oh, ow, pc.cout))
            d.out_h, d.out_w = oh, ow
        assert out.B == a0.B
        d.out = out.ptr
        d.out_pixel_stride = out.ps
        d.epilogue, d.flags = epi, flags
        if r1 is not None:
            d.r1, d.r1_pixel_stride = r1.ptr, r1.ps
        if r2 is not None:
            d.r2, d.r2_pixel_stride = r2.ptr, r2.ps
        d.tc_weights, d.tc_split = pc.tc, pc.tc_split
        L.check(self.lib.pcodec_conv_taps(C.byref(d), self.conv_impl, _stream()), f"conv_taps[{pc.name}]")
        return out

    def conv_new(self, pc: PackedConv, segs: Sequence[Act], epi: int = L.EPI_LINEAR, r1=None, r2=None) -> Act:
        a0 = segs[0]
        out = new_act(a0.B, a0.H // pc.in_step, a0.W // pc.in_step, pc.cout, self.device)
        return self.conv(pc, segs, out, epi, r1, r2)

    def conv_shuffle_new(self, pc: PackedConv, x: Act, epi: int) -> Act:
        out = new_act(x.B, 2 * x.H, 2 * x.W, pc.cout // 4, self.device)
        return self.conv(pc, [x], out, epi, flags=L.FLAG_PIXEL_SHUFFLE2)

    def deconv_new(self, phases: List[PackedConv], x: Act, epi: int = L.EPI_LINEAR, out: Optional[Act] = None) -> Act:
        if out is None:
            out = new_act(x.B, 2 * x.H, 2 * x.W, phases[0].cout, self.device)
        for ph in phases:
            self.conv(ph, [x], out, epi)
        return out

    def gdn_new(self, pc: PackedConv, x: Act, inverse: bool) -> Act:
        out = new_act(x.B, x.H, x.W, x.C, self.device)
        return self.conv(pc, [x], out, L.EPI_IGDN if inverse else L.EPI_GDN, r1=x, flags=L.FLAG_SQUARE_INPUT)

    # -- attention -----------------------------------------------------------------------------------------
    def window_attention(self, qkv: Act, rel_bias: Tensor, heads: int, ws: int, shift: int) -> Act:
        Cn = qkv.C // 3
        out = new_act(qkv.B, qkv.H, qkv.W, Cn, self.device)
        L.check(self.lib.pcodec_window_attention(qkv.ptr, qkv.ps, out.ptr, out.ps, rel_bias.data_ptr(), qkv.B, qkv.H,
                                                 qkv.W, Cn, heads, ws, shift, _stream()), "window_attention")
        return out

    # -- layout ----------------------------------------------------------------------------------------------
    def im2col_first(self, x_nchw: Tensor, k: int, stride: int, pad: int, k_pad: int) -> Act:
        B, Cn, H, W = x_nchw.shape
        oh, ow = H // stride, W // stride
        out = new_act(B, oh, ow, k_pad, self.device)
        L.check(self.lib.pcodec_im2col_nchw(x_nchw.data_ptr(), out.ptr, B, Cn, H, W, k, stride, pad, oh, ow, k_pad,
                                            _stream()), "im2col_nchw")
        return out

    def to_nchw(self, a: Act) -> Tensor:
        out = torch.empty((a.B, a.C, a.H, a.W), dtype=torch.float32, device=self.device)
        L.check(self.lib.pcodec_nhwc_to_nchw(a.ptr, a.ps, out.data_ptr(), a.B, a.C, a.H * a.W, _stream()),
                "nhwc_to_nchw")
        return out

    def from_nchw(self, t: Tensor, c_pad: Optional[int] = None) -> Act:
        B, Cn, H, W = t.shape
        c_pad = c_pad or Cn
        out = new_act(B, H, W, c_pad, self.device)
        t = t.contiguous().float()
        L.check(self.lib.pcodec_nchw_to_nhwc(t.data_ptr(), out.ptr, B, Cn, H * W, c_pad, c_pad, _stream()),
                "nchw_to_nhwc")
        return out

    # -- entropy-side kernels ----------------------------------------------------------------------------
    def quantile_threshold(self, scale: Act, q: float) -> Tensor:
        thr = torch.empty((scale.B,), dtype=torch.float32, device=self.device)
        L.check(self.lib.pcodec_quantile_threshold(scale.ptr, scale.B, scale.H * scale.W, scale.C, scale.ps,
                                                   float(torch.tensor(q, dtype=torch.float32).item()), thr.data_ptr(),
                                                   None, _stream()), "quantile_threshold")
        return thr

    def slice_quantize(self, y: Optional[Act], y_sub: Optional[Act], mu: Optional[Act], scale: Act, mask_mode: int,
                       thr: Optional[Tensor], table: Tensor, bound: float, symbols: Optional[Tensor],
                       indexes: Optional[Tensor], mask_out: Optional[Tensor], lik: Optional[Tensor],
                       y_hat: Optional[Act]) -> None:
        p = lambda t: t.data_ptr() if t is not None else None
        ap = lambda a: a.ptr if a is not None else None
        aps = lambda a: a.ps if a is not None else 0
        L.check(self.lib.pcodec_slice_quantize(ap(y), aps(y), ap(y_sub), aps(y_sub), ap(mu), aps(mu), scale.ptr,
                                               scale.ps, scale.B, scale.H * scale.W, scale.C, mask_mode, p(thr),
                                               table.data_ptr(), table.numel(), bound, p(symbols), p(indexes),
                                               p(mask_out), p(lik), ap(y_hat), aps(y_hat), _stream()), "slice_quantize")

    def slice_dequantize(self, symbols: Tensor, mu: Act, y_hat: Act) -> None:
        L.check(self.lib.pcodec_slice_dequantize(symbols.data_ptr(), mu.ptr, mu.ps, mu.B, mu.H * mu.W, mu.C, y_hat.ptr,
                                                 y_hat.ps, _stream()), "slice_dequantize")

    def bottleneck_quantize(self, z: Act, medians: Tensor, symbols: Optional[Tensor], indexes: Optional[Tensor],
                            z_hat: Optional[Act]) -> None:
        p = lambda t: t.data_ptr() if t is not None else None
        L.check(self.lib.pcodec_bottleneck_quantize(z.ptr, z.ps, medians.data_ptr(), z.B, z.H * z.W, z.C, p(symbols),
                                                    p(indexes), z_hat.ptr if z_hat else None,
                                                    z_hat.ps if z_hat else 0, _stream()), "bottleneck_quantize")

    def bottleneck_dequantize(self, symbols: Tensor, medians: Tensor, z_hat: Act) -> None:
        L.check(self.lib.pcodec_bottleneck_dequantize(symbols.data_ptr(), medians.data_ptr(), z_hat.B,
                                                      z_hat.H * z_hat.W, z_hat.C, z_hat.ptr, z_hat.ps, _stream()),
                "bottleneck_dequantize")

    def bottleneck_indexes(self, B: int, hw: int, channels: int) -> Tensor:
        idx = torch.empty((B, channels * hw), dtype=torch.int32, device=self.device)
        L.check(self.lib.pcodec_bottleneck_indexes(B, hw, channels, idx.data_ptr(), _stream()), "bottleneck_indexes")
        return idx

    def bottleneck_likelihood(self, z_hat: Act, params: Tensor) -> Tensor:
        lik = torch.empty((z_hat.B, z_hat.C, z_hat.H, z_hat.W), dtype=torch.float32, device=self.device)
        L.check(self.lib.pcodec_bottleneck_likelihood(z_hat.ptr, z_hat.ps, params.data_ptr(), z_hat.B,
                                                      z_hat.H * z_hat.W, z_hat.C, lik.data_ptr(), _stream()),
                "bottleneck_likelihood")
        return lik
