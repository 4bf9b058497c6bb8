"""Two-stage pipelining of the reference's evaluation sweep (training/step.py:322-337: for every quality level,
``compress`` then ``decompress``).

compress(q+1) does not depend on decompress(q), and the decoder is a serial chain of 16 entropy-decode phases that
leaves most of the GPU idle, so the sweep runs as two host threads on two CUDA streams: the caller's thread
compresses level after level, a worker thread decompresses each result as soon as it exists.  The tensor cores then
run the next level's analysis / parameter networks while the previous level's streams are being entropy-decoded.
Results are identical to calling compress()/decompress() back to back (same kernels, same order per image).
"""
from __future__ import annotations

import queue
import threading
from typing import Callable, List, Optional, Sequence

import torch

_streams_lock = threading.Lock()


def sweep(net, x: torch.Tensor, qualities: Sequence[float], mask_pol: Optional[str] = None, host_strings: bool = False,
          on_result: Optional[Callable[[float, dict, dict], None]] = None, keep: bool = True,
          decode_workers: Optional[int] = None,
          x_for_level: Optional[Callable[[float], torch.Tensor]] = None,
          graphs: Optional[bool] = None, encoders: Optional[int] = None) -> List[Optional[torch.Tensor]]:
    """compress + decompress `x` at every level of `qualities`; returns the reconstructions (``x_hat`` per level).

    host_strings=False keeps the rANS streams on the device between the two stages (compress(...,
    return_device_streams=True)); host_strings=True goes through python ``bytes`` exactly like the reference API.
    on_result(q, compressed, decompressed) is called on a worker thread after each level (its stream is
    synchronised at that point).  decode_workers: decompress() calls of different levels are independent too, so small
    batches (whose 16-phase decode chain leaves the GPU almost idle) run several of them concurrently; default
    max(1, min(6, 8 // batch)) — every worker owns the activation arenas of its image groups, so at 64 images of 768x512
    a second worker does not fit in 180 GB (measured: clean OutOfMemoryError at 171 GB).  x_for_level(q), when given, is called on the encoder stream before each level and
    returns that level's input (e.g. a fresh host->device upload); `x` then only fixes the device and batch size.
    With several decode workers the levels are coded in descending quality (the results keep the caller's order): the
    highest levels carry the most symbols, i.e. the longest serial decode chains, and those should start first instead of
    finishing the sweep alone.  encoders: concurrent compress() threads of a graphed sweep (default 4).
    graphs: replay the launch-bound network parts of both stages as CUDA graphs (graphs.py; default: batches of <= 2
    images, where the host — not the GPU — bounds the sweep).  The `masks` of a graphed compress() and the reconstructions
    handed to on_result alias static graph buffers: they are valid until the same level is coded again."""
    dev = x.device
    caller_stream = torch.cuda.current_stream(dev)
    use_graphs = (x.shape[0] <= 2) if graphs is None else bool(graphs)
    # (with graphs the host no longer limits how many levels are in flight: 8 workers measured 263 ms per 13-level sweep
    # of one 768x512 image against 279 ms with 6 and 371 ms with 4)
    n_workers = decode_workers if decode_workers else max(1, min(8 if use_graphs else 6, 8 // max(1, x.shape[0])))
    n_workers = max(1, min(8, n_workers))  # decompress() reserves 8 engine slots per worker (slots 8w+1 .. 8w+7)
    # One encoder stream and one stream per decode worker, created once per (model, device) and reused by every sweep:
    # torch hands out streams round-robin from a pool of 32, and every new stream gets its own caching-allocator pool,
    # so per-sweep streams grew the footprint by ~1.5 GB per sweep until the pool wrapped around.
    # graphs take the host out of the loop, and one small image leaves most SMs idle: code several levels at once
    n_enc = max(1, min(encoders if encoders else 4, 6, len(qualities))) if use_graphs else 1
    # processing order: position p of the sweep codes level order[p] (graphs, streams and workers are tied to p)
    order = list(range(len(qualities)))
    if n_workers > 1:
        order.sort(key=lambda i: -float(qualities[i]))
    with _streams_lock:
        cache = net.__dict__.setdefault("_pipeline_streams", {})
        have = cache.setdefault(dev, {"enc": [], "dec": []})
        while len(have["enc"]) < n_enc:
            have["enc"].append(torch.cuda.Stream(device=dev))
        while len(have["dec"]) < n_workers:
            have["dec"].append(torch.cuda.Stream(device=dev))
    enc_streams, dec_streams = have["enc"][:n_enc], have["dec"][:n_workers]
    enc_stream = enc_streams[0]
    for st in dec_streams + enc_streams:  # work of the caller (e.g. the upload of x) is ordered first
        st.wait_stream(caller_stream)
    g_enc, g_dec = {}, {}
    if use_graphs:
        # capture (once per model, shape, level and worker) on this thread, before the workers start
        from . import graphs as _graphs

        gc_ = _graphs.cache(net)
        B, H, W = x.shape[0], x.shape[2], x.shape[3]
        zshape, n_per = (H // 64, W // 64), 32 * (H // 16) * (W // 16)
        for p, i in enumerate(order):
            q = qualities[i]
            ek = p % n_enc
            k_enc = ("enc", dev, tuple(x.shape), float(q), mask_pol, ek)
            if k_enc not in gc_:  # encoder thread ek: engine slot 0 / 70 + ek, stream enc_streams[ek]
                gc_[k_enc] = _graphs.GraphedCompress(net, tuple(x.shape), q, mask_pol, enc_streams[ek],
                                                     slot=0 if ek == 0 else 70 + ek)
            g_enc[i] = gc_[k_enc]
            wk = p % n_workers
            k_dec = ("dec", dev, zshape, float(q), mask_pol, B, wk)
            if k_dec not in gc_:
                gc_[k_dec] = _graphs.GraphedDecompress(net, zshape, q, mask_pol, B, net.ns0 if q <= 0 else net.ns1, n_per,
                                                       wk, dec_streams[wk])
            g_dec[i] = gc_[k_dec]
    # with graphs a level belongs to a fixed worker (its graph lives on that worker's engine slot and stream)
    queues = [queue.Queue(maxsize=2) for _ in range(n_workers)] if use_graphs else None
    q_items: "queue.Queue" = queue.Queue(maxsize=n_workers + 1)
    outs: List[Optional[torch.Tensor]] = [None] * len(qualities)
    err: List[BaseException] = []

    def consumer(k: int):
        dec_stream = dec_streams[k]
        my_q = queues[k] if use_graphs else q_items
        try:
            with torch.cuda.device(dev), torch.cuda.stream(dec_stream), torch.no_grad():
                while True:
                    item = my_q.get()
                    if item is None:
                        if not use_graphs:
                            q_items.put(None)  # pass the end marker on to the other workers
                        return
                    i, q, c = item
                    src = c["strings"] if host_strings else c
                    if use_graphs:
                        r = {"x_hat": g_dec[i](src)}
                    else:
                        r = net.decompress(src, c["shape"], quality=q, mask_pol=mask_pol, _worker=k)
                    dec_stream.synchronize()  # `c` may be released (and its memory reused by the encoder stream) now
                    if on_result is not None:
                        on_result(q, c, r)
                    if keep:
                        outs[i] = r["x_hat"].clone() if use_graphs else r["x_hat"]
        except BaseException as e:  # noqa: BLE001 - re-raised on the caller's thread
            err.append(e)
            while my_q.get() is not None:  # drain so the producer never blocks on a dead consumer
                pass
            if not use_graphs:
                q_items.put(None)

    workers = [threading.Thread(target=consumer, args=(k,), name=f"pcodec-decompress-{k}") for k in range(n_workers)]
    for t in workers:
        t.start()
    def encode_levels(ek: int):
        """Encoder thread ek codes levels ek, ek + n_enc, ... in order and hands each to its decode worker."""
        try:
            with torch.cuda.device(dev), torch.cuda.stream(enc_streams[ek]), torch.no_grad():
                for p in range(ek, len(order), n_enc):
                    if err:
                        break
                    i = order[p]
                    q = qualities[i]
                    xq = x_for_level(q) if x_for_level is not None else x
                    if use_graphs:
                        c = g_enc[i](xq, return_device_streams=not host_strings)
                    else:
                        c = net.compress(xq, quality=q, mask_pol=mask_pol, return_device_streams=not host_strings)
                    enc_streams[ek].synchronize()  # compress() has already synchronised to learn the stream lengths
                    (queues[p % n_workers] if use_graphs else q_items).put((i, q, c))
        except BaseException as e:  # noqa: BLE001 - re-raised on the caller's thread
            err.append(e)

    encoders = [threading.Thread(target=encode_levels, args=(ek,), name=f"pcodec-compress-{ek}") for ek in range(1, n_enc)]
    try:
        for t in encoders:
            t.start()
        encode_levels(0)
        for t in encoders:
            t.join()
    finally:
        if use_graphs:
            for qq in queues:
                qq.put(None)
        else:
            q_items.put(None)
        for t in workers:
            t.join()
    if err:  # (first: after a device fault the stream calls below would raise a less informative error)
        raise err[0]
    for st in enc_streams + dec_streams:
        caller_stream.wait_stream(st)
    return outs
