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


def sweep(net, x: torch.Tensor, qualities: Sequence[float], mask_pol: Optional[str] = None, host_strings: bool = False,
          on_result: Optional[Callable[[float, dict, dict], None]] = None, keep: bool = True) -> List[Optional[torch.Tensor]]:
    """compress + decompress `x` at every level of `qualities`; returns the reconstructions (``x_hat`` per level).

    host_strings=False keeps the rANS streams on the device between the two stages (compress(...,
    return_device_streams=True)); host_strings=True goes through python ``bytes`` exactly like the reference API.
    on_result(q, compressed, decompressed) is called on the worker thread after each level (its stream is
    synchronised at that point)."""
    dev = x.device
    caller_stream = torch.cuda.current_stream(dev)
    enc_stream = torch.cuda.Stream(device=dev)
    dec_stream = torch.cuda.Stream(device=dev)
    enc_stream.wait_stream(caller_stream)
    q_items: "queue.Queue" = queue.Queue(maxsize=2)
    outs: List[Optional[torch.Tensor]] = [None] * len(qualities)
    err: List[BaseException] = []

    def consumer():
        try:
            with torch.cuda.device(dev), torch.cuda.stream(dec_stream), torch.no_grad():
                while True:
                    item = q_items.get()
                    if item is None:
                        return
                    i, q, c = item
                    src = c["strings"] if host_strings else c
                    r = net.decompress(src, c["shape"], quality=q, mask_pol=mask_pol)
                    dec_stream.synchronize()  # `c` may be released (and its memory reused by the encoder stream) now
                    if on_result is not None:
                        on_result(q, c, r)
                    if keep:
                        outs[i] = r["x_hat"]
        except BaseException as e:  # noqa: BLE001 - re-raised on the caller's thread
            err.append(e)
            while q_items.get() is not None:  # drain so the producer never blocks on a dead consumer
                pass

    t = threading.Thread(target=consumer, name="pcodec-decompress")
    t.start()
    try:
        with torch.cuda.stream(enc_stream), torch.no_grad():
            for i, q in enumerate(qualities):
                if err:
                    break
                c = net.compress(x, quality=q, mask_pol=mask_pol, return_device_streams=not host_strings)
                enc_stream.synchronize()  # compress() has already synchronised to learn the stream lengths
                q_items.put((i, q, c))
    finally:
        q_items.put(None)
        t.join()
    caller_stream.wait_stream(enc_stream)
    caller_stream.wait_stream(dec_stream)
    if err:
        raise err[0]
    return outs
