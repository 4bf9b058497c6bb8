"""CUDA execution engine: NHWC activations, packed weights, calls into the C-ABI.

Everything numeric here runs in libpcodec_b200.so; torch is used for device memory, streams and the
one-off weight repacking at prepare() time.
"""
from __future__ import annotations

import ctypes as C
import threading
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn as nn
from torch import Tensor

from . import _lib as L


def _stream():
    return torch.cuda.current_stream().cuda_stream


class _Root:
    """Per-buffer state shared by all channel windows of an activation: the split-fp16 planes (pcodec_planes: two fp16
    tensors with the buffer's geometry) and which channel ranges of them hold current data."""

    __slots__ = ("hi", "lo", "valid", "keep", "sq")

    def __init__(self):
        self.hi = 0          # device pointers of channel 0 of the planes (0 = not allocated)
        self.lo = 0
        self.valid = []      # disjoint [c0, c1) ranges whose planes are up to date
        self.keep = None     # owner of the plane memory when it is not arena memory
        self.sq = False      # the planes hold (x * 2^-4)^2 (written by the producing conv for the GDN that follows)

    def covers(self, c0: int, c1: int) -> bool:
        for a, b in self.valid:
            if a <= c0 and c1 <= b:
                return True
        return False

    def mark(self, c0: int, c1: int) -> None:
        out = []
        for a, b in self.valid:
            if b < c0 or c1 < a:
                out.append((a, b))
            else:
                c0, c1 = min(a, c0), max(b, c1)
        out.append((c0, c1))
        self.valid = out


class Act:
    """View of `C` channels starting at `c0` of an NHWC fp32 buffer [B, H, W, ps].

    Backed either by a torch tensor (`t`) or by a raw range of an `Arena`; kernels only need `ptr`/`ps`, so
    arena-backed views never construct a torch tensor on the hot path (`t` is materialised lazily)."""

    __slots__ = ("_t", "base", "B", "H", "W", "C", "c0", "ps", "_owner", "_root")

    def __init__(self, t: Optional[Tensor] = None, c0: int = 0, channels: Optional[int] = None, *, base: int = 0,
                 shape: Optional[Tuple[int, int, int, int]] = None, owner=None):
        if t is not None:
            assert t.dim() == 4 and t.dtype == torch.float32 and t.is_contiguous()
            self._t = t
            self.base = t.data_ptr()
            self.B, self.H, self.W, self.ps = t.shape
            self._owner = None
        else:
            self._t = None
            self.base = base
            self.B, self.H, self.W, self.ps = shape
            self._owner = owner
        self.c0 = c0
        self.C = self.ps - c0 if channels is None else channels
        self._root = _Root()  # shared by every channel window (slice) of this buffer
        assert 0 <= c0 and c0 + self.C <= self.ps

    @property
    def ptr(self) -> int:
        return self.base + 4 * self.c0

    @property
    def t(self) -> Tensor:
        if self._t is None:
            self._t = self._owner.view(self.base, (self.B, self.H, self.W, self.ps))
        return self._t

    def slice(self, c0: int, channels: int) -> "Act":
        a = Act.__new__(Act)
        a._t, a.base, a._owner, a._root = self._t, self.base, self._owner, self._root
        a.B, a.H, a.W, a.ps = self.B, self.H, self.W, self.ps
        a.c0, a.C = self.c0 + c0, channels
        assert a.c0 + channels <= self.ps
        return a

    def dense(self) -> Tensor:
        return self.t[..., self.c0:self.c0 + self.C]


def new_act(B: int, H: int, W: int, channels: int, device) -> Act:
    return Act(torch.empty((B, H, W, channels), dtype=torch.float32, device=device))


class Arena:
    """Bump allocator over one device buffer.  An inference call performs the same sequence of allocations
    every time it sees the same shapes, so resetting the arena at the start of a call makes every activation
    land at the same address as last time — which lets the engine reuse fully built kernel descriptors."""

    def __init__(self, device, nbytes: int = 64 << 20):
        self.device = device
        self.buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
        self.off = 0
        self.generation = 0
        self.retired: List[Tensor] = []

    def reset(self) -> None:
        self.off = 0
        self.retired.clear()  # buffers replaced during the previous call are no longer referenced by kernels in flight
                              # on this stream only after the next sync; they stay alive until the following reset

    def alloc(self, nbytes: int) -> int:
        nbytes = (nbytes + 255) & ~255
        if self.off + nbytes > self.buf.numel():
            # grow: keep the old buffer alive (kernels in flight / live views), start a fresh, larger one
            self.retired.append(self.buf)
            # (the new buffer starts empty: it has to hold what is allocated from here on, not what came before)
            self.buf = torch.empty(max(min(2 * self.buf.numel(), self.buf.numel() + (6 << 30)), nbytes + (256 << 20)),
                                   dtype=torch.uint8, device=self.device)
            self.off = 0
            self.generation += 1
        p = self.buf.data_ptr() + self.off
        self.off += nbytes
        return p

    def mark(self):
        return (self.generation, self.off)

    def release(self, mark) -> None:
        """Give back everything allocated since `mark` (stack discipline).  Safe because every kernel that touches
        the released range was enqueued on this context's stream before any later kernel that reuses it."""
        if mark[0] == self.generation:
            self.off = mark[1]

    def view(self, ptr: int, shape) -> Tensor:
        for b in [self.buf] + self.retired:
            o = ptr - b.data_ptr()
            if 0 <= o < b.numel():
                n = 1
                for d in shape:
                    n *= d
                return b[o:o + 4 * n].view(torch.float32).view(*shape)
        raise RuntimeError("pointer does not belong to this arena")


class PackedConv:
    """Weights of one tap-GEMM: fp32 [T][Cin][Cout] + bias + tap offsets."""

    __slots__ = ("w", "bias", "taps", "cin", "cout", "in_step", "out_step", "off", "name", "tc", "tc_split",
                 "image_channels")

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
        self.image_channels = 0  # set by pack_deconv_merged_image

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


def pack_deconv_phases(m: nn.ConvTranspose2d, device, name="", pad_cout_to: int = 0) -> List[PackedConv]:
    """ConvTranspose2d k5 s2 p2 op1 [Cin,Cout,5,5] as 4 sub-pixel phases (py,px): output (2h+py, 2w+px) gathers
    input (h+1-a, w+1-b) with kernel tap (py+2a, px+2b).  `pad_cout_to` zero-pads the output channels (the
    3-channel image layer is padded to 16 so it can run on the tcgen05 kernel; callers read channels [0, Cout))."""
    assert m.kernel_size == (5, 5) and m.stride == (2, 2) and m.padding == (2, 2) and m.output_padding == (1, 1)
    w = m.weight.detach().to(device=device, dtype=torch.float32)  # [Cin, Cout, 5, 5]
    b = m.bias.detach().to(device=device, dtype=torch.float32) if m.bias is not None else None
    if pad_cout_to and w.shape[1] < pad_cout_to:
        extra = pad_cout_to - w.shape[1]
        w = torch.cat([w, torch.zeros((w.shape[0], extra, 5, 5), dtype=w.dtype, device=device)], 1)
        if b is not None:
            b = torch.cat([b, torch.zeros(extra, dtype=b.dtype, device=device)])
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


def pack_deconv_merged_image(m: nn.ConvTranspose2d, device, name="") -> PackedConv:
    """Image layer (ConvTranspose2d k5 s2 p2 op1, Cout <= 4) as ONE 9-tap GEMM over the 3x3 input neighbourhood with
    16 output columns: column (2*py + px) * Cout + c holds sub-pixel phase (py, px) of channel c (zero weights where a
    phase does not use a tap).  The four phase launches re-read the 192-channel input 25 times for a handful of
    output channels; merged, it is read 9 times and the result goes straight to the NCHW image."""
    assert m.kernel_size == (5, 5) and m.stride == (2, 2) and m.padding == (2, 2) and m.output_padding == (1, 1)
    w = m.weight.detach().to(device=device, dtype=torch.float32)  # [Cin, Cout, 5, 5]
    cin, cout = w.shape[0], w.shape[1]
    assert 4 * cout <= 16
    taps = [(dy, dx) for dy in (1, 0, -1) for dx in (1, 0, -1)]
    packed = torch.zeros((9, cin, 16), dtype=torch.float32, device=device)
    bias = torch.zeros(16, dtype=torch.float32, device=device)
    b = m.bias.detach().to(device=device, dtype=torch.float32) if m.bias is not None else None
    for py in (0, 1):
        for px in (0, 1):
            col = (2 * py + px) * cout
            if b is not None:
                bias[col:col + cout] = b
            for t, (dy, dx) in enumerate(taps):
                a, bb = 1 - dy, 1 - dx  # input (h + 1 - a, w + 1 - b) <-> kernel tap (py + 2a, px + 2b)
                if a < (3 if py == 0 else 2) and bb < (3 if px == 0 else 2):
                    packed[t, :, col:col + cout] = w[:, :, py + 2 * a, px + 2 * bb]
    pc = PackedConv(packed, bias, taps, in_step=1, out_step=1, name=f"{name}.merged")
    pc.image_channels = cout
    return pc


def pack_gdn(g, device, name="") -> PackedConv:
    beta, gamma = g.effective()  # gamma [C_out, C_in]
    gamma = gamma.to(device=device, dtype=torch.float32)
    return PackedConv(gamma.t().reshape(1, gamma.shape[1], gamma.shape[0]),
                      beta.to(device=device, dtype=torch.float32), [(0, 0)], name=name)


class _ArenaScope:
    __slots__ = ("arena", "m")

    def __init__(self, arena: Arena):
        self.arena = arena

    def __enter__(self):
        self.m = self.arena.mark()
        return self

    def __exit__(self, *exc):
        self.arena.release(self.m)
        return False


class Engine:
    """Stateless helpers that launch kernels on the current stream."""

    def __init__(self, device, conv_impl: int = 0):
        self.device = device
        self.lib = L.lib()
        self.conv_impl = conv_impl
        self._tls = threading.local()
        self._slots = {}
        self._slots_lock = threading.Lock()

    # -- per-slot call state: arena, stream handle, cached descriptors -----------------------------------
    class _Ctx:
        __slots__ = ("arena", "descs", "stream")

    def begin(self, slot: int = 0) -> None:
        """Start of an inference call: bind this thread to context `slot` (slot 0 = caller's thread, slot g+1 =
        decode group g), rewind its arena and latch the current CUDA stream."""
        with self._slots_lock:
            ctx = self._slots.get(slot)
            if ctx is None:
                ctx = Engine._Ctx()
                ctx.arena = Arena(self.device)
                ctx.descs = {}
                self._slots[slot] = ctx
        ctx.arena.reset()
        ctx.stream = torch.cuda.current_stream(self.device).cuda_stream
        self._tls.ctx = ctx

    def _state(self):
        ctx = getattr(self._tls, "ctx", None)
        if ctx is None:
            self.begin(0)
            ctx = self._tls.ctx
        return ctx

    def stream(self) -> int:
        return self._state().stream

    def scope(self):
        """`with E.scope():` — activations allocated inside are temporaries, released on exit."""
        return _ArenaScope(self._state().arena)

    def act(self, B: int, H: int, W: int, channels: int, fmt: int = 3) -> Act:
        """Arena-backed NHWC activation (valid until the next begin() on this thread).  fmt: which representations get
        memory — 1 = fp32 only (consumers are not convolutions: GDN's x, attention's qkv), 2 = split-fp16 planes only
        (the only consumers are convolutions), 3 = both.  The plane space is reserved with the buffer so that it lives
        exactly as long as the data, whatever scope first converts it."""
        ar = self._state().arena
        n = B * H * W * channels
        p = ar.alloc((4 * n if fmt & 1 else 0) + (4 * n if fmt & 2 else 0))
        a = Act(base=p if fmt & 1 else 0, shape=(B, H, W, channels), owner=ar)
        if fmt & 2:
            q = p + (4 * n if fmt & 1 else 0)
            a._root.hi, a._root.lo = q, q + 2 * n
        return a

    # -- split-fp16 planes (operand format of the fp16 tcgen05 kernel) -----------------------------------------
    def _alloc_planes(self, a: Act) -> None:
        """Arena activations get their plane space with the buffer itself (Engine.act), so it lives exactly as long as the
        fp32 data whatever scope first converts it; torch-backed activations get a torch buffer of their own."""
        r = a._root
        if r.hi == 0:
            n = a.B * a.H * a.W * a.ps
            r.keep = torch.empty(2 * n, dtype=torch.float16, device=self.device)
            r.hi = r.keep.data_ptr()
            r.lo = r.hi + 2 * n

    def planes(self, a: Act):
        """(hi pointer, lo pointer) of window `a` in its buffer's planes, converting from fp32 when the window is not
        current (inputs that no convolution epilogue produced: im2col patches, attention output, dequantised latents).
        A buffer created without plane space (fmt = 1) gets a temporary conversion in the current arena scope."""
        r = a._root
        c1 = a.c0 + a.C
        if r.hi and r.covers(a.c0, c1):
            return r.hi + 2 * a.c0, r.lo + 2 * a.c0
        if (a.c0 | a.C | a.ps) & 7 or a.base == 0:
            return None
        if r.hi == 0 and a._owner is not None:  # arena activation without plane space: uncached temporary
            n = a.B * a.H * a.W * a.ps
            q = self._state().arena.alloc(4 * n)
            hi, lo = q, q + 2 * n
        else:
            self._alloc_planes(a)
            hi, lo = r.hi, r.lo
            r.mark(a.c0, c1)
        L.check(self.lib.pcodec_split_planes(a.ptr, a.ps, a.B * a.H * a.W, a.C, hi + 2 * a.c0, lo + 2 * a.c0, a.ps, 0,
                                             self.stream()), "split_planes")
        return hi + 2 * a.c0, lo + 2 * a.c0

    def squared_planes(self, a: Act):
        """Temporary planes holding (x * 2^-4)^2 of window `a` (GDN operand), as a stand-alone Act-like triple."""
        if (a.C | a.ps) & 7:
            return None
        n = 2 * a.B * a.H * a.W * a.C
        base = self._state().arena.alloc(2 * n)
        L.check(self.lib.pcodec_split_planes(a.ptr, a.ps, a.B * a.H * a.W, a.C, base, base + n, a.C, 1, self.stream()),
                "split_planes(square)")
        return base, base + n, a.C

    # -- generic tap conv --------------------------------------------------------------------------------
    def conv(self, pc: PackedConv, segs: Sequence[Act], out: Act, epi: int = L.EPI_LINEAR, r1: Optional[Act] = None,
             r2: Optional[Act] = None, flags: int = 0, fmt: int = 3, square_planes: bool = False) -> Act:
        """fmt: what the epilogue writes — 1 = fp32 NHWC only, 2 = split-fp16 planes only (the consumer is another
        convolution), 3 = both.  square_planes: the planes get (v * 2^-4)^2, the operand of the GDN that follows."""
        tls = self._state()
        if square_planes and self.conv_impl in (0, 3) and pc.tc is not None and out._root.hi and not ((out.c0 | out.ps) & 3):
            flags |= L.FLAG_SQUARE_OUT_PLANES
            fmt = 3
        key = (id(pc), pc.w.data_ptr(), pc.tc, pc.tc_split, out.ptr, out.ps, epi, flags, fmt,
               (r1.ptr if r1.base else -(r1._root.hi + 2 * r1.c0)) if r1 is not None else 0,
               (r2.ptr if r2.base else -(r2._root.hi + 2 * r2.c0)) if r2 is not None else 0,
               segs[0].B, segs[0].H, segs[0].W) + tuple((s.ptr, s.C, s.ps) for s in segs)
        use16 = self.conv_impl in (0, 3) and pc.tc is not None
        seg_planes = None
        if use16:
            if flags & L.FLAG_SQUARE_INPUT:
                x0 = segs[0]
                if x0._root.sq:  # the producing conv already wrote the squares
                    sq = (x0._root.hi + 2 * x0.c0, x0._root.lo + 2 * x0.c0, x0.ps)
                else:
                    sq = self.squared_planes(x0)
                seg_planes = [sq] if sq is not None else None
            else:
                seg_planes = []
                for s_ in segs:
                    p = self.planes(s_)
                    if p is None:
                        seg_planes = None
                        break
                    seg_planes.append((p[0], p[1], s_.ps))
            if seg_planes is not None and (fmt & 2) and ((out.c0 | out.ps) & 3 or out._root.hi == 0):
                fmt = 1  # no plane space behind this output (or a window the 8-byte plane stores cannot address)
            if out.base == 0:
                if seg_planes is None or not (fmt & 2):
                    raise L.PcodecError(f"conv[{pc.name}]: planes-only output, but the fp16 kernel cannot run this launch")
                fmt = 2
            key = key + (tuple(p_[0] for p_ in seg_planes) if seg_planes else 0, out._root.hi, fmt)
        if not use16 or seg_planes is None:
            fmt = 1
            if out.base == 0 or any(s_.base == 0 for s_ in segs):
                raise L.PcodecError(f"conv[{pc.name}]: planes-only activation reached a launch the fp16 kernel cannot take")
        ent = tls.descs.get(key)
        if ent is None:
            ent = self._build_desc(pc, segs, out, epi, r1, r2, flags, seg_planes, fmt)
            if len(tls.descs) > 20000:
                tls.descs.clear()
            tls.descs[key] = ent
        d, _plan, has_plan = ent
        if use16 and not has_plan and self.conv_impl == 3:
            raise L.PcodecError(f"conv[{pc.name}]: the fp16 tensor-core kernel cannot take this launch")
        L.check(self.lib.pcodec_conv_taps(d, self.conv_impl if (has_plan or self.conv_impl != 3) else 0, tls.stream), pc.name)
        if has_plan and d.out_hi and (flags & L.FLAG_SQUARE_OUT_PLANES):
            out._root.sq = True
        elif has_plan and d.out_hi:
            out._root.mark(out.c0, out.c0 + (pc.cout // 4 if flags & L.FLAG_PIXEL_SHUFFLE2 else pc.cout))
        elif out._root.valid:  # fp32 rewritten without planes: whatever planes covered this window are stale now
            out._root.valid = [v for v in out._root.valid if v[1] <= out.c0 or v[0] >= out.c0 + out.C]
        return out

    def _build_desc(self, pc: PackedConv, segs: Sequence[Act], out: Act, epi: int, r1: Optional[Act],
                    r2: Optional[Act], flags: int, seg_planes=None, fmt: int = 1):
        a0 = segs[0]
        d = L.ConvDesc()
        cin = 0
        for i, s in enumerate(segs):
            assert (s.B, s.H, s.W) == (a0.B, a0.H, a0.W)
            d.seg[i].ptr = s.ptr if s.base else None  # planes-only activation: no fp32 data behind it
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
            d.out_h, d.out_w = 2 * oh, 2 * ow   # the kernel indexes the shuffled tensor with these
        else:
            assert (out.H, out.W, out.C) == (oh, ow, pc.cout), (pc.name, (out.H, out.W, out.C), (oh, ow, pc.cout))
            d.out_h, d.out_w = oh, ow
        assert out.B == a0.B
        d.out = out.ptr if out.base else None
        d.out_pixel_stride = out.ps
        d.epilogue, d.flags = epi, flags
        for r, name in ((r1, "r1"), (r2, "r2")):
            if r is None:
                continue
            if r.base:
                setattr(d, name, r.ptr)
                setattr(d, name + "_pixel_stride", r.ps)
            else:  # planes-only residual: the epilogue reconstructs hi + lo * 2^-11
                if not r._root.covers(r.c0, r.c0 + r.C) or r._root.sq:
                    raise L.PcodecError(f"conv[{pc.name}]: residual operand has neither fp32 data nor current planes")
                pl = getattr(d, name + "_16")
                pl.hi, pl.lo, pl.pixel_stride = r._root.hi + 2 * r.c0, r._root.lo + 2 * r.c0, r.ps
        d.tc_weights, d.tc_split = pc.tc, pc.tc_split
        plan, has_plan = None, False
        if seg_planes is not None:
            for i, (hi, lo, ps) in enumerate(seg_planes):
                d.seg16[i].hi, d.seg16[i].lo, d.seg16[i].pixel_stride = hi, lo, ps
            if (fmt & 2) and out._root.hi:
                d.out_hi, d.out_lo = out._root.hi + 2 * out.c0, out._root.lo + 2 * out.c0
                d.out_plane_stride = out.ps
                if not (fmt & 1):
                    d.flags = flags | L.FLAG_NO_F32_OUT
            plan = (C.c_ubyte * L.CONV_PLAN_BYTES)()
            d.plan = C.cast(plan, C.c_void_p)
            rc = self.lib.pcodec_conv_plan(d)
            if rc == L.OK:
                has_plan = True
            elif rc == L.ERR_UNSUPPORTED:
                if out.base == 0 or any(s.base == 0 for s in segs):
                    raise L.PcodecError(f"conv_plan[{pc.name}]: unsupported launch with planes-only activations")
                d.plan = None
                d.out_hi = d.out_lo = None
                d.flags = flags
            else:
                L.check(rc, f"conv_plan[{pc.name}]")
        return d, plan, has_plan

    def _out_fmt(self, pc: PackedConv, fmt: int) -> int:
        """Planes-only / planes outputs exist only on the fp16 tensor-core path."""
        return fmt if (self.conv_impl in (0, 3) and pc.tc is not None) else 1

    def conv_new(self, pc: PackedConv, segs: Sequence[Act], epi: int = L.EPI_LINEAR, r1=None, r2=None, fmt: int = 3,
                 square_planes: bool = False) -> Act:
        a0 = segs[0]
        fmt = self._out_fmt(pc, 3 if square_planes else fmt)
        out = self.act(a0.B, a0.H // pc.in_step, a0.W // pc.in_step, pc.cout, fmt)
        return self.conv(pc, segs, out, epi, r1, r2, fmt=fmt, square_planes=square_planes)

    def conv_shuffle_new(self, pc: PackedConv, x: Act, epi: int, fmt: int = 3) -> Act:
        fmt = self._out_fmt(pc, fmt)
        out = self.act(x.B, 2 * x.H, 2 * x.W, pc.cout // 4, fmt)
        return self.conv(pc, [x], out, epi, flags=L.FLAG_PIXEL_SHUFFLE2, fmt=fmt)

    def deconv_new(self, phases: List[PackedConv], x: Act, epi: int = L.EPI_LINEAR, out: Optional[Act] = None,
                   fmt: int = 3, square_planes: bool = False) -> Act:
        fmt = self._out_fmt(phases[0], 3 if square_planes else fmt)
        if out is None:
            out = self.act(x.B, 2 * x.H, 2 * x.W, phases[0].cout, fmt)
        for ph in phases:
            self.conv(ph, [x], out, epi, fmt=fmt, square_planes=square_planes)
        return out

    def deconv_image(self, pc: PackedConv, x: Act, epi: int) -> Tensor:
        """Merged-phase image layer (pack_deconv_merged_image): NHWC activation -> NCHW image [B, C, 2H, 2W]."""
        Cimg = pc.image_channels
        out = torch.empty((x.B, Cimg, 2 * x.H, 2 * x.W), dtype=torch.float32, device=self.device)
        tls = self._state()
        use16 = self.conv_impl in (0, 3) and pc.tc is not None
        pl = self.planes(x) if use16 else None
        key = ("img", id(pc), pc.tc, pc.tc_split, x.ptr, x.ps, x.B, x.H, x.W, epi, pl[0] if pl else 0)
        ent = tls.descs.get(key)
        if ent is None:
            d = L.ConvDesc()
            d.seg[0].ptr, d.seg[0].channels, d.seg[0].pixel_stride = x.ptr, x.C, x.ps
            d.n_segments = 1
            d.batch, d.in_h, d.in_w = x.B, x.H, x.W
            d.n_taps = len(pc.taps)
            for t, (dy, dx) in enumerate(pc.taps):
                d.dy[t], d.dx[t] = dy, dx
            d.in_step = 1
            d.weight, d.bias = pc.w.data_ptr(), pc.bias.data_ptr()
            d.cin_total, d.cout = pc.cin, pc.cout
            d.grid_h, d.grid_w = x.H, x.W
            d.out_step, d.out_off_y, d.out_off_x = 1, 0, 0
            d.out_h, d.out_w = 2 * x.H, 2 * x.W
            d.out_pixel_stride = Cimg
            d.epilogue, d.flags = epi, L.FLAG_SUBPIXEL_NCHW
            d.tc_weights, d.tc_split = pc.tc, pc.tc_split
            d.out = out.data_ptr()
            plan, has_plan = None, False
            if pl is not None:
                d.seg16[0].hi, d.seg16[0].lo, d.seg16[0].pixel_stride = pl[0], pl[1], x.ps
                plan = (C.c_ubyte * L.CONV_PLAN_BYTES)()
                d.plan = C.cast(plan, C.c_void_p)
                has_plan = self.lib.pcodec_conv_plan(d) == L.OK
                if not has_plan:
                    d.plan = None
            ent = (d, plan, has_plan)
            tls.descs[key] = ent
        d = ent[0]
        d.out = out.data_ptr()
        L.check(self.lib.pcodec_conv_taps(d, self.conv_impl if (ent[2] or self.conv_impl != 3) else 0, tls.stream), pc.name)
        return out

    def gdn_new(self, pc: PackedConv, x: Act, inverse: bool, fmt: int = 3) -> Act:
        fmt = self._out_fmt(pc, fmt)
        out = self.act(x.B, x.H, x.W, x.C, fmt)
        return self.conv(pc, [x], out, L.EPI_IGDN if inverse else L.EPI_GDN, r1=x, flags=L.FLAG_SQUARE_INPUT, fmt=fmt)

    # -- attention -----------------------------------------------------------------------------------------
    def window_attention(self, qkv: Act, rel_bias: Tensor, heads: int, ws: int, shift: int) -> Act:
        Cn = qkv.C // 3
        out = self.act(qkv.B, qkv.H, qkv.W, Cn)
        L.check(self.lib.pcodec_window_attention(qkv.ptr, qkv.ps, out.ptr, out.ps, rel_bias.data_ptr(), qkv.B, qkv.H,
                                                 qkv.W, Cn, heads, ws, shift, self.stream()), "window_attention")
        return out

    # -- layout ----------------------------------------------------------------------------------------------
    def im2col_first(self, x_nchw: Tensor, k: int, stride: int, pad: int, k_pad: int) -> Act:
        B, Cn, H, W = x_nchw.shape
        oh, ow = H // stride, W // stride
        out = self.act(B, oh, ow, k_pad)
        L.check(self.lib.pcodec_im2col_nchw(x_nchw.data_ptr(), out.ptr, B, Cn, H, W, k, stride, pad, oh, ow, k_pad,
                                            self.stream()), "im2col_nchw")
        return out

    def to_nchw(self, a: Act) -> Tensor:
        out = torch.empty((a.B, a.C, a.H, a.W), dtype=torch.float32, device=self.device)
        L.check(self.lib.pcodec_nhwc_to_nchw(a.ptr, a.ps, out.data_ptr(), a.B, a.C, a.H * a.W, self.stream()),
                "nhwc_to_nchw")
        return out

    def from_nchw(self, t: Tensor, c_pad: Optional[int] = None) -> Act:
        B, Cn, H, W = t.shape
        c_pad = c_pad or Cn
        out = self.act(B, H, W, c_pad)
        t = t.contiguous().float()
        L.check(self.lib.pcodec_nchw_to_nhwc(t.data_ptr(), out.ptr, B, Cn, H * W, c_pad, c_pad, self.stream()),
                "nchw_to_nhwc")
        return out

    # -- entropy-side kernels ----------------------------------------------------------------------------
    def quantile_threshold(self, scale: Act, q: float) -> Tensor:
        thr = torch.empty((scale.B,), dtype=torch.float32, device=self.device)
        L.check(self.lib.pcodec_quantile_threshold(scale.ptr, scale.B, scale.H * scale.W, scale.C, scale.ps,
                                                   float(torch.tensor(q, dtype=torch.float32).item()), thr.data_ptr(),
                                                   None, self.stream()), "quantile_threshold")
        return thr

    def slice_quantize(self, y: Optional[Act], y_sub: Optional[Act], mu: Optional[Act], scale: Act, mask_mode: int,
                       thr: Optional[Tensor], table: Tensor, bound: float, symbols: Optional[Tensor],
                       indexes: Optional[Tensor], mask_out: Optional[Tensor], lik: Optional[Tensor],
                       y_hat: Optional[Act], mask_src: Optional[Act] = None) -> None:
        p = lambda t: t.data_ptr() if t is not None else None
        ap = lambda a: a.ptr if a is not None else None
        aps = lambda a: a.ps if a is not None else 0
        L.check(self.lib.pcodec_slice_quantize_cust(ap(y), aps(y), ap(y_sub), aps(y_sub), ap(mu), aps(mu), scale.ptr,
                                                    scale.ps, scale.B, scale.H * scale.W, scale.C, mask_mode, p(thr),
                                                    table.data_ptr(), table.numel(), bound, p(symbols), p(indexes),
                                                    p(mask_out), p(lik), ap(y_hat), aps(y_hat), ap(mask_src), aps(mask_src),
                                                    self.stream()), "slice_quantize")

    def layer_partition(self, scale: Act, thresholds: Tensor, in_a: Optional[Tensor], in_b: Optional[Tensor],
                        out_a: Optional[Tensor], out_b: Optional[Tensor], counts: Optional[Tensor],
                        avail: Optional[Tensor] = None) -> None:
        """Stable partition of a slice's NCHW-order planes by progressive layer (gather), or its inverse (scatter,
        when `avail` is given).  thresholds: float32 [n_levels, B]; counts: int32 [B, 16]."""
        p = lambda t: t.data_ptr() if t is not None else None
        assert thresholds.dtype == torch.float32 and thresholds.is_contiguous() and thresholds.shape[1] == scale.B
        L.check(self.lib.pcodec_layer_partition(scale.ptr, scale.ps, scale.B, scale.H * scale.W, scale.C,
                                                thresholds.data_ptr(), thresholds.shape[0], p(in_a), p(in_b), p(out_a),
                                                p(out_b), p(counts), p(avail), 1 if avail is not None else 0,
                                                self.stream()), "layer_partition")

    def masked_residual(self, ret: Act, identity: Act, sigma: Act, mode_star: int, thr_star: Optional[Tensor],
                        mode_bar: int, thr_bar: Optional[Tensor], out: Act) -> None:
        """out = ret * round(star - bar) + identity (REM wrapper, CHProgREM.py:73-85, 375-400)."""
        p = lambda t: t.data_ptr() if t is not None else None
        assert ret.C == identity.C == out.C and sigma.C == 32
        L.check(self.lib.pcodec_masked_residual(ret.ptr, ret.ps, identity.ptr, identity.ps, sigma.ptr, sigma.ps, ret.B,
                                                ret.H * ret.W, ret.C, mode_star, p(thr_star), mode_bar, p(thr_bar),
                                                out.ptr, out.ps, self.stream()), "masked_residual")

    def slice_dequantize(self, symbols: Tensor, mu: Act, y_hat: Act) -> None:
        L.check(self.lib.pcodec_slice_dequantize(symbols.data_ptr(), mu.ptr, mu.ps, mu.B, mu.H * mu.W, mu.C, y_hat.ptr,
                                                 y_hat.ps, self.stream()), "slice_dequantize")

    def bottleneck_quantize(self, z: Act, medians: Tensor, symbols: Optional[Tensor], indexes: Optional[Tensor],
                            z_hat: Optional[Act]) -> None:
        p = lambda t: t.data_ptr() if t is not None else None
        L.check(self.lib.pcodec_bottleneck_quantize(z.ptr, z.ps, medians.data_ptr(), z.B, z.H * z.W, z.C, p(symbols),
                                                    p(indexes), z_hat.ptr if z_hat else None,
                                                    z_hat.ps if z_hat else 0, self.stream()), "bottleneck_quantize")

    def bottleneck_dequantize(self, symbols: Tensor, medians: Tensor, z_hat: Act) -> None:
        L.check(self.lib.pcodec_bottleneck_dequantize(symbols.data_ptr(), medians.data_ptr(), z_hat.B,
                                                      z_hat.H * z_hat.W, z_hat.C, z_hat.ptr, z_hat.ps, self.stream()),
                "bottleneck_dequantize")

    def bottleneck_indexes(self, B: int, hw: int, channels: int) -> Tensor:
        idx = torch.empty((B, channels * hw), dtype=torch.int32, device=self.device)
        L.check(self.lib.pcodec_bottleneck_indexes(B, hw, channels, idx.data_ptr(), self.stream()), "bottleneck_indexes")
        return idx

    def bottleneck_likelihood(self, z_hat: Act, params: Tensor) -> Tensor:
        lik = torch.empty((z_hat.B, z_hat.C, z_hat.H, z_hat.W), dtype=torch.float32, device=self.device)
        L.check(self.lib.pcodec_bottleneck_likelihood(z_hat.ptr, z_hat.ps, params.data_ptr(), z_hat.B,
                                                      z_hat.H * z_hat.W, z_hat.C, lik.data_ptr(), self.stream()),
                "bottleneck_likelihood")
        return lik
