"""Truncatable progressive container: ONE encode serves every quality level by byte-stream truncation.

The reference has no such format — ``compress(x, q)`` returns python lists for one quality at a time
(CHProg_cnn.py:686-847) and its evaluation re-encodes the image at every level (training/step.py:322-337).  With
``all_scalable=True`` the entropy parameters (mu, sigma) of a progressive slice are conditioned on the base slice and
on the *parameters* of earlier progressive slices (CHProg_cnn.py:596-597, 796-797), never on decoded progressive
symbols, so they do not depend on the quality level; and the variance-aware masks of increasing levels are nested
(``sigma >= quantile(sigma, 1 - 0.1*pr)``, masking.py:205-223).  Every latent element therefore ENTERS at exactly
one level with a level-independent symbol ``round(y - mu)``, which gives an embedded stream:

    header | z | base slice 0..9 | layer 0: slice 0..9 | layer 1: slice 0..9 | ...

Layer k of slice i is a stand-alone rANS stream (reference arithmetic, rans_interface.cpp) over the elements whose
sigma lies in [thr_k, thr_{k-1}), in the coder's NCHW order.  A prefix that ends after layer k reconstructs exactly
what the reference protocol ``decompress(compress(x, levels[k]), levels[k])`` reconstructs: same symbols inside the
mask, zeros outside (tests/test_gpu_container.py).  The decoder recomputes sigma and the thresholds, hence the layer
membership, before it touches a progressive byte — and because nothing depends on decoded progressive symbols, all
progressive streams of all slices and images are entropy-decoded by ONE launch.

Byte layout (little endian), one container per image:
    0   4s  magic "PCB2"      4  u8 version (1)   5  u8 n_levels   6  u8 n_base   7  u8 n_prog
    8   u32 H   12 u32 W  (padded image size the streams were coded at)   16 u16 zh   18 u16 zw
    20  f32 levels[n_levels]
    ..  u32 z_len | u32 base_len[n_base] | u32 layer_len[n_levels][n_prog]
    ..  payload in the order above (zero-length streams — empty layers — occupy no bytes)
"""
from __future__ import annotations

import struct
from typing import Dict, List, Optional, Sequence

import torch
from torch import Tensor

from . import _lib as L
from . import ans as _ans
from .layers import ChannelMask

MAGIC = b"PCB2"
VERSION = 1
DEFAULT_LEVELS = (0.05, 0.1, 0.25, 0.5, 0.6, 0.75, 1, 1.25, 2, 3, 5, 10)  # train.py:293 without the base level 0


def _require(net) -> None:
    if not getattr(net, "all_scalable", False):
        raise L.PcodecError("the progressive container needs a model built with all_scalable=True: only then are the "
                            "entropy parameters of the progressive slices independent of the quality level")


def _thresholds(E, scale, levels: Sequence[float], mask_pol) -> Tensor:
    """float32 [n_levels, B]: element enters at level k iff sigma >= thr[k] (and not at an earlier level)."""
    rows = []
    for lev in levels:
        kind, q = ChannelMask.mode_for(mask_pol, lev)
        if kind == "threshold":
            rows.append(E.quantile_threshold(scale, q))
        else:
            fill = float("-inf") if kind == "ones" else float("inf")
            rows.append(torch.full((scale.B,), fill, dtype=torch.float32, device=E.device))
    return torch.stack(rows, 0).contiguous()


class Header:
    def __init__(self, levels, H, W, zh, zw, n_base, n_prog, z_len, base_len, layer_len):
        self.levels, self.H, self.W, self.zh, self.zw = list(levels), H, W, zh, zw
        self.n_base, self.n_prog, self.z_len, self.base_len, self.layer_len = n_base, n_prog, z_len, base_len, layer_len

    @property
    def size(self) -> int:
        nl = len(self.levels)
        return 20 + 4 * nl + 4 * (1 + self.n_base + nl * self.n_prog)

    def pack(self) -> bytes:
        nl = len(self.levels)
        out = struct.pack("<4sBBBBIIHH", MAGIC, VERSION, nl, self.n_base, self.n_prog, self.H, self.W, self.zh, self.zw)
        out += struct.pack(f"<{nl}f", *self.levels)
        out += struct.pack(f"<{1 + self.n_base + nl * self.n_prog}I", self.z_len, *self.base_len,
                           *[v for row in self.layer_len for v in row])
        return out

    @staticmethod
    def parse(blob: bytes) -> "Header":
        if len(blob) < 20 or blob[:4] != MAGIC:
            raise L.PcodecError("not a PCB2 progressive container")
        _m, ver, nl, nb, npg, H, W, zh, zw = struct.unpack_from("<4sBBBBIIHH", blob, 0)
        if ver != VERSION:
            raise L.PcodecError(f"unsupported container version {ver}")
        need = 20 + 4 * nl + 4 * (1 + nb + nl * npg)
        if len(blob) < need:
            raise L.PcodecError("container truncated inside the header")
        levels = list(struct.unpack_from(f"<{nl}f", blob, 20))
        lens = struct.unpack_from(f"<{1 + nb + nl * npg}I", blob, 20 + 4 * nl)
        # untrusted input: the device decoder reads 32-bit words, so a stream whose length is not a multiple of 4
        # would shift every later stream to a misaligned address (a sticky GPU fault, not a python exception)
        if any(v % 4 or 0 < v < 8 for v in lens):
            raise L.PcodecError("corrupt container: rANS stream lengths must be multiples of 4 and at least 8 bytes")
        layer = [list(lens[1 + nb + k * npg:1 + nb + (k + 1) * npg]) for k in range(nl)]
        return Header(levels, H, W, zh, zw, nb, npg, lens[0], list(lens[1:1 + nb]), layer)

    def prefix_end(self, n_layers: int) -> int:
        """Byte offset just past the last stream of the first `n_layers` layers (0 = base only)."""
        return self.size + self.z_len + sum(self.base_len) + sum(sum(r) for r in self.layer_len[:n_layers])

    def layers_in(self, nbytes: int) -> int:
        """Number of COMPLETE layers present in a prefix of `nbytes` bytes (-1: not even the base)."""
        if nbytes < self.prefix_end(0):
            return -1
        k = 0
        while k < len(self.levels) and nbytes >= self.prefix_end(k + 1):
            k += 1
        return k


def truncate(blob: bytes, n_layers: int) -> bytes:
    """The shortest prefix that decodes at `n_layers` progressive layers (0 = base quality)."""
    return blob[:Header.parse(blob).prefix_end(n_layers)]


@torch.no_grad()
def encode_progressive(net, x: Tensor, levels: Sequence[float] = DEFAULT_LEVELS, mask_pol: Optional[str] = None) -> List[bytes]:
    """One encode of `x` [B,3,H,W] (H, W multiples of 64) -> one truncatable container per image."""
    _require(net)
    # the header stores the levels as f32 and the decoder recomputes the quantiles from what it reads there: round
    # them to f32 HERE so that both sides derive bit-identical thresholds for any caller-supplied level
    levels = [struct.unpack("<f", struct.pack("<f", float(v)))[0] for v in levels]
    if any(b <= a for a, b in zip(levels, levels[1:])) or not levels or levels[0] <= 0 or len(levels) > 15:
        raise ValueError("levels must be increasing, positive and at most 15")
    mask_pol = net.mask_policy if mask_pol is None else mask_pol
    x = net._check_input(x)
    P = net.prepare()
    E = P["eng"]
    dev = E.device
    y, z, z_sym, z_idx, _zl, lm, ls = net._encoder_front(P, x, enhanced=True, want_z_lik=False)
    B, h, w = y.B, y.H, y.W
    n = 32 * h * w
    ns0, n_prog, nl = net.ns0, net.ns1 - net.ns0, len(levels)
    table, bound = P["scale_table"], P["scale_bound"]
    sym_b = torch.empty((ns0, B, n), dtype=torch.int32, device=dev)
    idx_b = torch.empty((ns0, B, n), dtype=torch.int32, device=dev)

    def code_base(i, mu, scale, y_pre):
        E.slice_quantize(y.slice(32 * i, 32), None, mu, scale, L.MASK_ONES, None, table, bound, sym_b[i], idx_b[i], None,
                         None, y_pre)

    y_hat_base = net._base_slices(P, lm, ls, code_base)
    sym_p = torch.empty((B, n), dtype=torch.int32, device=dev)
    idx_p = torch.empty((B, n), dtype=torch.int32, device=dev)
    csym = torch.empty((n_prog, B, n), dtype=torch.int32, device=dev)   # layer-major compacted planes
    cidx = torch.empty((n_prog, B, n), dtype=torch.int32, device=dev)
    counts = torch.zeros((n_prog, B, 16), dtype=torch.int32, device=dev)

    def code_prog(i, mu, scale, _mask_mode, _thr, y_pre):
        y_sub = y.slice(32 * i, 32) if net.delta_encode else None
        E.slice_quantize(y.slice(32 * (ns0 + i), 32), y_sub, mu, scale, L.MASK_ONES, None, table, bound, sym_p, idx_p, None,
                         None, y_pre)
        E.layer_partition(scale, _thresholds(E, scale, levels, mask_pol), sym_p, idx_p, csym[i], cidx[i], counts[i])

    net._prog_slices(P, lm, ls, y_hat_base, 10, mask_pol, code_prog, "codec", deferred=[])
    # streams: z, base (slice-major), layers (level-major, then slice, then image)
    z_data, z_off = _ans.encode_batch(z_sym, z_idx, P["eb_tables"])
    b_data, b_off = _ans.encode_batch(sym_b.reshape(ns0 * B, n), idx_b.reshape(ns0 * B, n), P["gc_tables"])
    cnt = counts[:, :, :nl].to(torch.int64)                                   # [n_prog, B, nl]
    first = torch.cumsum(cnt, 2) - cnt                                        # exclusive prefix inside a (slice, image) row
    row0 = (torch.arange(n_prog * B, device=dev, dtype=torch.int64) * n).reshape(n_prog, B, 1)
    seg_start = (row0 + first).permute(2, 0, 1).contiguous().reshape(-1)      # [nl, n_prog, B]
    seg_count = cnt.permute(2, 0, 1).contiguous().reshape(-1).to(torch.int32)
    l_data, l_off = _ans.encode_segments(csym.reshape(-1), cidx.reshape(-1), seg_start, seg_count, P["gc_tables"], n)
    seg_count_h = seg_count.cpu().reshape(nl, n_prog, B)
    z_str = _ans.split_streams(z_data, z_off)
    b_str = _ans.split_streams(b_data, b_off)
    l_str = _ans.split_streams(l_data, l_off)
    blobs = []
    for b in range(B):
        base = [b_str[s * B + b] for s in range(ns0)]
        layers = [[l_str[(k * n_prog + i) * B + b] if int(seg_count_h[k, i, b]) > 0 else b"" for i in range(n_prog)]
                  for k in range(nl)]
        hdr = Header(levels, x.shape[2], x.shape[3], z.H, z.W, ns0, n_prog, len(z_str[b]), [len(s) for s in base],
                     [[len(s) for s in row] for row in layers])
        blobs.append(hdr.pack() + z_str[b] + b"".join(base) + b"".join(s for row in layers for s in row))
    return blobs


@torch.no_grad()
def decode_progressive(net, blobs: Sequence[bytes], n_layers: Optional[int] = None, mask_pol: Optional[str] = None,
                       debug: Optional[dict] = None) -> Dict:
    """Decode a batch of (possibly truncated) containers.  Uses the first `n_layers` progressive layers (default: as many
    complete layers as every blob still holds).  Returns {"x_hat": [B,3,H,W] in [0,1], "n_layers", "level"}."""
    _require(net)
    mask_pol = net.mask_policy if mask_pol is None else mask_pol
    hdrs = [Header.parse(b) for b in blobs]
    h0 = hdrs[0]
    for hd in hdrs[1:]:
        if (hd.levels, hd.H, hd.W, hd.zh, hd.zw, hd.n_base, hd.n_prog) != (h0.levels, h0.H, h0.W, h0.zh, h0.zw, h0.n_base,
                                                                          h0.n_prog):
            raise L.PcodecError("containers of one batch must share the image size and the level list")
    have = min(hd.layers_in(len(b)) for hd, b in zip(hdrs, blobs))
    if have < 0:
        raise L.PcodecError("container truncated before the end of the base streams")
    k_use = have if n_layers is None else n_layers
    if k_use > have:
        raise L.PcodecError(f"requested {k_use} layers but only {have} are complete in every container")
    P = net.prepare()
    E = P["eng"]
    dev = E.device
    B, ns0, n_prog, nl = len(blobs), h0.n_base, h0.n_prog, len(h0.levels)
    if (ns0, n_prog) != (net.ns0, net.ns1 - net.ns0):
        raise L.PcodecError("container slice layout does not match the model")
    # split the payloads
    z_str, base_str = [], [[None] * B for _ in range(ns0)]
    layer_str = [[[b""] * B for _ in range(n_prog)] for _ in range(k_use)]
    for b, (hd, blob) in enumerate(zip(hdrs, blobs)):
        o = hd.size
        z_str.append(blob[o:o + hd.z_len])
        o += hd.z_len
        for s in range(ns0):
            base_str[s][b] = blob[o:o + hd.base_len[s]]
            o += hd.base_len[s]
        for k in range(k_use):
            for i in range(n_prog):
                layer_str[k][i][b] = blob[o:o + hd.layer_len[k][i]]
                o += hd.layer_len[k][i]
    z_data, z_off = _ans.pack_streams(z_str, dev)
    y_data, y_off = _ans.pack_streams([s for sl in base_str for s in sl], dev)
    lm, ls, y_hat_base, _dec = net._decode_base(P, y_data, y_off.to(dev), z_data, z_off.to(dev), B, 0, B, (h0.zh, h0.zw),
                                                enhanced=True, slot=1)
    if k_use == 0:
        return {"x_hat": net._g_s(P, y_hat_base, 0, clamp=True), "n_layers": 0, "level": 0.0}
    levels = h0.levels[:k_use]
    h, w = 4 * h0.zh, 4 * h0.zw
    n = 32 * h * w
    table, bound = P["scale_table"], P["scale_bound"]
    idx_p = torch.empty((B, n), dtype=torch.int32, device=dev)
    cidx = torch.empty((n_prog, B, n), dtype=torch.int32, device=dev)
    counts = torch.zeros((n_prog, B, 16), dtype=torch.int32, device=dev)
    thr_all, scales, mus, y_pres = [], [], [], []

    def code_prog(i, mu, scale, _mask_mode, _thr, y_pre):
        E.slice_quantize(None, None, None, scale, L.MASK_ONES, None, table, bound, None, idx_p, None, None, None)
        thr = _thresholds(E, scale, levels, mask_pol)
        E.layer_partition(scale, thr, None, idx_p, None, cidx[i], counts[i])
        thr_all.append(thr)
        scales.append(scale)
        mus.append(mu)
        y_pres.append(y_pre)

    deferred: list = []
    y_hat_q = net._prog_slices(P, lm, ls, y_hat_base, 10, mask_pol, code_prog, "codec", deferred=deferred)
    # one launch decodes every (layer, slice, image) segment
    l_data, l_off = _ans.pack_streams([layer_str[k][i][b] for k in range(k_use) for i in range(n_prog) for b in range(B)], dev)
    l_off = l_off.to(dev)
    cnt = counts[:, :, :k_use].to(torch.int64)
    first = torch.cumsum(cnt, 2) - cnt
    row0 = (torch.arange(n_prog * B, device=dev, dtype=torch.int64) * n).reshape(n_prog, B, 1)
    seg_start = (row0 + first).permute(2, 0, 1).contiguous().reshape(-1)
    seg_count = cnt.permute(2, 0, 1).contiguous().reshape(-1).to(torch.int32)
    csym = torch.zeros((n_prog, B, n), dtype=torch.int32, device=dev)
    _ans.decode_segments(l_data, l_off[:-1].contiguous(), l_off[1:].contiguous(), seg_start, seg_count, cidx.reshape(-1),
                         csym.reshape(-1), P["gc_tables"])
    avail = torch.full((B,), k_use, dtype=torch.int32, device=dev)
    sym = torch.empty((B, n), dtype=torch.int32, device=dev)
    for i in range(n_prog):
        E.layer_partition(scales[i], thr_all[i], csym[i], None, sym, None, None, avail=avail)
        E.slice_dequantize(sym, mus[i], y_pres[i])
    if debug is not None:  # tests only
        debug["y_pre"] = [E.to_nchw(a) for a in y_pres]
    net._run_deferred_lrp(P, deferred)
    if debug is not None:
        debug["y_hat"] = E.to_nchw(y_hat_q)
    return {"x_hat": net._g_s(P, y_hat_q, 1, clamp=True), "n_layers": k_use, "level": h0.levels[k_use - 1]}
