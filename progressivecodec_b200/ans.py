"""Drop-in for ``compressai.ans`` on the B200 coder kernels.

Mirrors the pybind11 module of the reference (compress/cpp_exts/rans/rans_interface.cpp:352-372):
``RansEncoder``, ``RansDecoder``, ``BufferedRansEncoder`` with the same method names, argument order
(symbols, indexes, cdfs, cdfs_sizes, offsets) and return types (bytes / list[int]).  The python-list
API is kept for drop-in use and for parity tests that read like the reference's call sites
(entropy_models.py:227-235, 276-286); the model path uses ``encode_batch`` / ``decode_batch`` on device
tensors and never converts to lists.

Unlike the reference, streams shorter than 4 symbols are safe (reference bug: rans_interface.cpp:170).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np
import torch

from . import _lib as L


def _stream():
    return torch.cuda.current_stream().cuda_stream


class CdfTables:
    """Device-resident CDF table set (the `_quantized_cdf/_cdf_length/_offset` buffer contract,
    entropy_models.py:98-100)."""

    def __init__(self, cdfs: torch.Tensor, sizes: torch.Tensor, offsets: torch.Tensor, device=None):
        device = device or torch.device("cuda", torch.cuda.current_device())
        self.cdfs = cdfs.to(device=device, dtype=torch.int32).contiguous()
        self.sizes = sizes.to(device=device, dtype=torch.int32).contiguous()
        self.offsets = offsets.to(device=device, dtype=torch.int32).contiguous()
        assert self.cdfs.dim() == 2 and self.sizes.numel() == self.cdfs.shape[0] == self.offsets.numel()

    @staticmethod
    def from_lists(cdfs: Sequence[Sequence[int]], sizes: Sequence[int], offsets: Sequence[int]) -> "CdfTables":
        width = max(len(r) for r in cdfs)
        t = torch.zeros((len(cdfs), width), dtype=torch.int32)
        for i, r in enumerate(cdfs):
            t[i, : len(r)] = torch.tensor(list(r), dtype=torch.int32)
        return CdfTables(t, torch.tensor(list(sizes), dtype=torch.int32), torch.tensor(list(offsets), dtype=torch.int32))


def encode_batch(symbols: torch.Tensor, indexes: torch.Tensor, tables: CdfTables,
                 words_per_symbol: float = 0.5) -> Tuple[torch.Tensor, torch.Tensor]:
    """symbols/indexes: int32 CUDA tensors [S, N] (S independent streams).  Returns (bytes uint8 CUDA tensor,
    offsets int64 CPU tensor [S+1]).  One device->host sync (to learn the stream lengths)."""
    assert symbols.is_cuda and indexes.is_cuda and symbols.dtype == torch.int32 and indexes.dtype == torch.int32
    assert symbols.shape == indexes.shape and symbols.dim() == 2
    symbols, indexes = symbols.contiguous(), indexes.contiguous()
    S, N = symbols.shape
    dev = symbols.device
    lib = L.lib()
    while True:
        cap_words = int(N * words_per_symbol) + 16
        scratch = torch.empty((S, cap_words), dtype=torch.int32, device=dev)
        n_words = torch.empty((S,), dtype=torch.int32, device=dev)
        out_cap = S * cap_words * 4
        out = torch.empty((out_cap,), dtype=torch.uint8, device=dev)
        meta = torch.empty((S + 2,), dtype=torch.int64, device=dev)  # offsets[S+1] + status (int32 view)
        status = meta[S + 1:].view(torch.int32)
        L.check(lib.pcodec_rans_encode_batch(symbols.data_ptr(), indexes.data_ptr(), S, N, tables.cdfs.data_ptr(),
                                             tables.cdfs.shape[1], tables.sizes.data_ptr(), tables.offsets.data_ptr(),
                                             tables.cdfs.shape[0], scratch.data_ptr(), cap_words, n_words.data_ptr(),
                                             out.data_ptr(), out_cap, meta.data_ptr(), status.data_ptr(), _stream()),
                "rans_encode_batch")
        host = meta.cpu()
        if int(host[S + 1:].view(torch.int32)[0]) == L.ERR_OVERFLOW:
            if words_per_symbol >= 6.0:
                raise L.PcodecError("rans_encode_batch: stream exceeds 6 words/symbol")
            words_per_symbol = min(6.0, words_per_symbol * 4)
            continue
        offsets = host[: S + 1]
        return out[: int(offsets[S])], offsets


def decode_batch(data: torch.Tensor, offsets: torch.Tensor, indexes: torch.Tensor, tables: CdfTables) -> torch.Tensor:
    """data: uint8 CUDA tensor, offsets int64 [S+1] (any device), indexes int32 CUDA [S, N] -> int32 CUDA [S, N]."""
    assert data.is_cuda and indexes.is_cuda and indexes.dtype == torch.int32 and indexes.dim() == 2
    indexes = indexes.contiguous()
    S, N = indexes.shape
    offs = offsets.to(device=indexes.device, dtype=torch.int64).contiguous()
    out = torch.empty((S, N), dtype=torch.int32, device=indexes.device)
    L.check(L.lib().pcodec_rans_decode_batch(data.data_ptr(), offs.data_ptr(), S, N, indexes.data_ptr(),
                                             tables.cdfs.data_ptr(), tables.cdfs.shape[1], tables.sizes.data_ptr(),
                                             tables.offsets.data_ptr(), tables.cdfs.shape[0], out.data_ptr(), _stream()),
            "rans_decode_batch")
    return out


def decode_ranges(data: torch.Tensor, starts: torch.Tensor, ends: torch.Tensor, indexes: torch.Tensor,
                  tables: CdfTables) -> torch.Tensor:
    """Like decode_batch for streams at arbitrary byte ranges [starts[s], ends[s]) of `data` (int64 CUDA tensors)."""
    assert data.is_cuda and indexes.is_cuda and indexes.dtype == torch.int32 and indexes.dim() == 2
    indexes = indexes.contiguous()
    S, N = indexes.shape
    starts, ends = starts.contiguous(), ends.contiguous()
    assert starts.is_cuda and ends.is_cuda and starts.numel() == S and ends.numel() == S
    out = torch.empty((S, N), dtype=torch.int32, device=indexes.device)
    L.check(L.lib().pcodec_rans_decode_ranges(data.data_ptr(), starts.data_ptr(), ends.data_ptr(), S, N,
                                              indexes.data_ptr(), tables.cdfs.data_ptr(), tables.cdfs.shape[1],
                                              tables.sizes.data_ptr(), tables.offsets.data_ptr(), tables.cdfs.shape[0],
                                              out.data_ptr(), _stream()), "rans_decode_ranges")
    return out


def encode_segments(symbols: torch.Tensor, indexes: torch.Tensor, seg_start: torch.Tensor, seg_count: torch.Tensor,
                    tables: CdfTables, max_count: int, words_per_symbol: float = 0.5) -> Tuple[torch.Tensor, torch.Tensor]:
    """Variable-length streams: stream s codes elements [seg_start[s], seg_start[s]+seg_count[s]) of the flat int32
    CUDA arrays.  seg_start int64 / seg_count int32 CUDA tensors [S].  Returns (bytes, int64 CPU offsets [S+1])."""
    assert symbols.is_cuda and indexes.is_cuda and seg_start.is_cuda and seg_count.is_cuda
    assert seg_start.dtype == torch.int64 and seg_count.dtype == torch.int32
    symbols, indexes = symbols.contiguous(), indexes.contiguous()
    S = seg_start.numel()
    dev = symbols.device
    lib = L.lib()
    while True:
        cap_words = int(max_count * words_per_symbol) + 16
        scratch = torch.empty((S, cap_words), dtype=torch.int32, device=dev)
        n_words = torch.empty((S,), dtype=torch.int32, device=dev)
        out_cap = S * cap_words * 4
        out = torch.empty((out_cap,), dtype=torch.uint8, device=dev)
        meta = torch.empty((S + 2,), dtype=torch.int64, device=dev)
        status = meta[S + 1:].view(torch.int32)
        L.check(lib.pcodec_rans_encode_segments(symbols.data_ptr(), indexes.data_ptr(), seg_start.data_ptr(),
                                                seg_count.data_ptr(), S, tables.cdfs.data_ptr(), tables.cdfs.shape[1],
                                                tables.sizes.data_ptr(), tables.offsets.data_ptr(), tables.cdfs.shape[0],
                                                scratch.data_ptr(), cap_words, n_words.data_ptr(), out.data_ptr(), out_cap,
                                                meta.data_ptr(), status.data_ptr(), _stream()), "rans_encode_segments")
        host = meta.cpu()
        if int(host[S + 1:].view(torch.int32)[0]) == L.ERR_OVERFLOW:
            if words_per_symbol >= 6.0:
                raise L.PcodecError("rans_encode_segments: stream exceeds 6 words/symbol")
            words_per_symbol = min(6.0, words_per_symbol * 4)
            continue
        offsets = host[: S + 1]
        return out[: int(offsets[S])], offsets


def decode_segments(data: torch.Tensor, starts: torch.Tensor, ends: torch.Tensor, seg_start: torch.Tensor,
                    seg_count: torch.Tensor, indexes: torch.Tensor, out: torch.Tensor, tables: CdfTables) -> torch.Tensor:
    """Inverse of encode_segments: stream s (bytes [starts[s], ends[s]) of `data`) fills out[seg_start[s] : +count]."""
    assert data.is_cuda and indexes.is_cuda and out.is_cuda and out.dtype == torch.int32 and indexes.dtype == torch.int32
    S = seg_start.numel()
    L.check(L.lib().pcodec_rans_decode_segments(data.data_ptr(), starts.data_ptr(), ends.data_ptr(), S,
                                                seg_start.data_ptr(), seg_count.data_ptr(), indexes.data_ptr(),
                                                tables.cdfs.data_ptr(), tables.cdfs.shape[1], tables.sizes.data_ptr(),
                                                tables.offsets.data_ptr(), tables.cdfs.shape[0], out.data_ptr(), _stream()),
            "rans_decode_segments")
    return out


def pack_streams(strings: Sequence[bytes], device) -> Tuple[torch.Tensor, torch.Tensor]:
    """Host byte strings -> (uint8 CUDA blob with 8 bytes of zero slack, int64 CPU offsets).  The blob is staged in pinned
    host memory and copied on the current stream; the bulk copies are numpy's (they release the GIL, which the pipelined
    sweep's other host threads are launching kernels under)."""
    lens = [len(s) for s in strings]
    if any(n % 4 for n in lens):  # the device decoder reads aligned 32-bit words (the coder only ever emits whole words)
        raise L.PcodecError("rANS stream length is not a multiple of 4 bytes: not a stream of this coder / corrupted")
    if any(0 < n < 8 for n in lens):  # every stream ends with the 64-bit coder state (rans_interface.cpp:170-189)
        raise L.PcodecError("rANS stream shorter than the 8-byte final state: not a stream of this coder / corrupted")
    offs = torch.zeros(len(strings) + 1, dtype=torch.int64)
    offs[1:] = torch.cumsum(torch.tensor(lens, dtype=torch.int64), 0)
    total = int(offs[-1])
    host = torch.empty(total + 8, dtype=torch.uint8, pin_memory=True)
    buf = host.numpy()
    if total:
        buf[:total] = np.frombuffer(b"".join(strings), dtype=np.uint8)
    buf[total:] = 0
    return host.to(device, non_blocking=True), offs


def split_streams(data: torch.Tensor, offsets: torch.Tensor) -> List[bytes]:
    """uint8 CUDA blob + offsets -> python `bytes` per stream (one device->host copy into pinned memory)."""
    host = torch.empty(data.numel(), dtype=torch.uint8, pin_memory=True)
    host.copy_(data, non_blocking=True)
    torch.cuda.current_stream(data.device).synchronize()
    raw = host.numpy()
    o = offsets.tolist()
    return [raw[o[i]:o[i + 1]].tobytes() for i in range(len(o) - 1)]


def _dev():
    if not torch.cuda.is_available():
        raise L.PcodecError("progressivecodec_b200.ans needs a CUDA device (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


class RansEncoder:
    """rans_interface.cpp:193-204."""

    def encode_with_indexes(self, symbols, indexes, cdfs, cdfs_sizes, offsets) -> bytes:
        dev = _dev()
        tables = CdfTables.from_lists(cdfs, cdfs_sizes, offsets)
        sym = torch.tensor(list(symbols), dtype=torch.int32, device=dev).reshape(1, -1)
        idx = torch.tensor(list(indexes), dtype=torch.int32, device=dev).reshape(1, -1)
        data, offs = encode_batch(sym, idx, tables)
        return split_streams(data, offs)[0]


class BufferedRansEncoder:
    """rans_interface.cpp:99-191: buffer several encode_with_indexes calls, one stream at flush()."""

    def __init__(self):
        self._sym: List[int] = []
        self._idx: List[int] = []
        self._tables = None

    def encode_with_indexes(self, symbols, indexes, cdfs, cdfs_sizes, offsets) -> None:
        # The reference resolves (start, range) at call time, so each call may use different tables; we support the
        # (universal) case of one table set per flush and verify it.
        key = (tuple(cdfs_sizes), tuple(offsets))
        if self._tables is None:
            self._tables = (key, CdfTables.from_lists(cdfs, cdfs_sizes, offsets))
        elif self._tables[0] != key:
            raise L.PcodecError("BufferedRansEncoder: all calls before flush() must use the same CDF tables")
        self._sym += list(symbols)
        self._idx += list(indexes)

    def flush(self) -> bytes:
        dev = _dev()
        if self._tables is None:
            tables = CdfTables.from_lists([[0, 65536]], [2], [0])
        else:
            tables = self._tables[1]
        sym = torch.tensor(self._sym, dtype=torch.int32, device=dev).reshape(1, -1)
        idx = torch.tensor(self._idx, dtype=torch.int32, device=dev).reshape(1, -1)
        self._sym, self._idx, self._tables = [], [], None
        data, offs = encode_batch(sym, idx, tables)
        return split_streams(data, offs)[0]


class RansDecoder:
    """rans_interface.cpp:206-350."""

    def __init__(self):
        self._stream = None
        self._consumed: List[Tuple[List[int], tuple]] = []

    def decode_with_indexes(self, encoded: bytes, indexes, cdfs, cdfs_sizes, offsets) -> List[int]:
        dev = _dev()
        tables = CdfTables.from_lists(cdfs, cdfs_sizes, offsets)
        blob, offs = pack_streams([encoded], dev)
        idx = torch.tensor(list(indexes), dtype=torch.int32, device=dev).reshape(1, -1)
        return decode_batch(blob, offs, idx, tables)[0].tolist()

    def set_stream(self, encoded: bytes) -> None:
        self._stream = encoded
        self._done_idx: List[int] = []
        self._tables = None

    def decode_stream(self, indexes, cdfs, cdfs_sizes, offsets) -> List[int]:
        """Continue decoding where the previous decode_stream stopped (rans_interface.cpp:285-350).  The decoder
        state after k symbols is a function of the stream prefix, so we re-decode the prefix on the device
        (one launch) instead of keeping device state alive between python calls."""
        if self._stream is None:
            raise L.PcodecError("decode_stream before set_stream")
        key = (tuple(cdfs_sizes), tuple(offsets))
        if self._tables is None:
            self._tables = (key, CdfTables.from_lists(cdfs, cdfs_sizes, offsets))
        elif self._tables[0] != key:
            raise L.PcodecError("RansDecoder.decode_stream: all calls on one stream must use the same CDF tables")
        dev = _dev()
        start = len(self._done_idx)
        self._done_idx += list(indexes)
        blob, offs = pack_streams([self._stream], dev)
        idx = torch.tensor(self._done_idx, dtype=torch.int32, device=dev).reshape(1, -1)
        return decode_batch(blob, offs, idx, self._tables[1])[0, start:].tolist()
