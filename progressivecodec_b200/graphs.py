"""CUDA-graph replay of compress() / decompress() for small batches.

One 768x512 image launches ~750 kernels per compress() and ~750 per decompress(); each grid covers a handful of SMs and
runs for tens of microseconds, so the host (python + ctypes, ~25 us per launch, under the GIL that the encoder thread
and the decode workers share) is the bound, not the GPU: 13 compress() calls alone took 291 ms.  Everything between the
input and the entropy coder (compress) and between the packed streams and x_hat (decompress of one image group) is free
of host synchronisation, allocates from the engine's arena at addresses that repeat from call to call, and takes its
thresholds from device memory — so it is captured ONCE per (shape, quality) into a CUDA graph and replayed with one
launch.  Same kernels, same order, same results (tests/test_gpu_model.py).

Buffers a graph reads or writes are static: the input image, the packed streams with their offsets (copied in before a
replay) and the outputs (symbol planes / x_hat: valid until the next replay of the SAME graph — pipeline.sweep consumes
them before that).  The arena buffer the activations live in is pinned by the graph object.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from . import ans as _ans


class GraphedCompress:
    """net.compress(x, quality) for one input shape: graph of the network part + eager entropy-coding tail."""

    def __init__(self, net, x_shape, quality, mask_pol, stream: torch.cuda.Stream, slot: int = 0):
        dev = net._device()
        self.net, self.quality, self.mask_pol, self.stream = net, quality, mask_pol, stream
        self._state = net.prepare()  # the weights / tables / arenas the graph points into live as long as the graph
        self.x = torch.zeros(x_shape, dtype=torch.float32, device=dev)
        with torch.cuda.stream(stream), torch.no_grad():
            for _ in range(2):  # sizes the arena, builds the launch plans, sets the function attributes
                net.compress(self.x, quality=quality, mask_pol=mask_pol, _planes_only=True, _slot=slot)
            stream.synchronize()
            E = net.prepare()["eng"]
            self._pin = E._slots[slot].arena.buf  # the activations' memory must outlive the graph
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, stream=stream, capture_error_mode="thread_local"):
                self.planes = net.compress(self.x, quality=quality, mask_pol=mask_pol, _planes_only=True, _slot=slot)
            if E._slots[slot].arena.buf is not self._pin:
                raise RuntimeError("arena moved during graph capture")

    def __call__(self, x: torch.Tensor, return_device_streams: bool):
        if self.net.prepare() is not self._state:
            raise RuntimeError("the model was re-prepared after this graph was captured")
        self.x.copy_(x, non_blocking=True)
        self.graph.replay()
        return self.net._entropy_tail(self.planes, return_device_streams)


class GraphedDecompress:
    """net.decompress(...) of a whole small batch (one image group) at one quality, on one decode worker's slot."""

    def __init__(self, net, shape, quality, mask_pol, batch: int, n_slices: int, n_per_stream: int, worker: int,
                 stream: torch.cuda.Stream):
        dev = net._device()
        mask_pol = net.mask_policy if mask_pol is None else mask_pol  # (decompress() resolves the default the same way)
        self.net, self.quality, self.batch, self.n_slices = net, quality, batch, n_slices
        self.shape = torch.Size([int(shape[0]), int(shape[1])])
        Cz = net.entropy_bottleneck._quantized_cdf.size(0)
        nz = Cz * int(shape[0]) * int(shape[1])
        # capacities: the encoder refuses streams above 6 words per symbol
        self.y_data = torch.zeros(24 * n_slices * batch * n_per_stream + 64, dtype=torch.uint8, device=dev)
        self.z_data = torch.zeros(24 * batch * nz + 64, dtype=torch.uint8, device=dev)
        self.y_off = torch.zeros(n_slices * batch + 1, dtype=torch.int64, device=dev)
        self.z_off = torch.zeros(batch + 1, dtype=torch.int64, device=dev)
        self.slot = 1 + 8 * worker
        P = self._state = net.prepare()
        E = P["eng"]
        # warm-up needs DECODABLE input (the decoder of garbage is safe but data dependent in time, not in launches):
        # all-zero offsets decode zero-length streams, which the kernel treats as streams of zero words
        with torch.cuda.stream(stream), torch.no_grad():
            for _ in range(2):
                net._decompress_group(P, self.y_data, self.y_off, self.z_data, self.z_off, batch, 0, batch, self.shape,
                                      quality, mask_pol, slot=self.slot)
            stream.synchronize()
            self._pin = E._slots[self.slot].arena.buf
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, stream=stream, capture_error_mode="thread_local"):
                self.x_hat = net._decompress_group(P, self.y_data, self.y_off, self.z_data, self.z_off, batch, 0, batch,
                                                   self.shape, quality, mask_pol, slot=self.slot)
            if E._slots[self.slot].arena.buf is not self._pin:
                raise RuntimeError("arena moved during graph capture")

    def __call__(self, src) -> torch.Tensor:
        """src: the dict of compress(return_device_streams=True) or the reference-API `strings` list."""
        dev = self.y_data.device
        if self.net.prepare() is not self._state:
            raise RuntimeError("the model was re-prepared after this graph was captured")
        if isinstance(src, dict):
            y_data, y_off, z_data, z_off = src["streams"]
        else:
            z_data, z_off = _ans.pack_streams(list(src[1]), dev)
            y_data, y_off = _ans.pack_streams([s for sl in src[0] for s in sl], dev)
        ny, nzb = int(y_off[-1]), int(z_off[-1])
        if ny > self.y_data.numel() - 8 or nzb > self.z_data.numel() - 8 or y_off.numel() != self.y_off.numel():
            raise ValueError("streams do not fit the graph's static buffers")
        self.y_data[:ny].copy_(y_data[:ny], non_blocking=True)
        self.z_data[:nzb].copy_(z_data[:nzb], non_blocking=True)
        self.y_off.copy_(y_off, non_blocking=True)
        self.z_off.copy_(z_off, non_blocking=True)
        self.graph.replay()
        return self.x_hat


def cache(net) -> Dict[Tuple, object]:
    """Graphs of the model's CURRENT prepared state.  They are stored inside that state: a graph has the device addresses
    of the packed weights, the entropy tables and the engine's arenas baked in, and all of those are rebuilt (the old
    ones freed) whenever the model is re-prepared — after update(), load_state_dict() or any .to() / .cuda() call, even
    one that moves nothing.  A cache on the model itself outlived them and replayed graphs over freed memory."""
    return net.prepare().setdefault("_graph_cache", {})
