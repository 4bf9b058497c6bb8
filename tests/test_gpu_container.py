"""Truncatable progressive container (BASELINE config 4): one encode, every quality by truncation.

Parity definition: a prefix that ends after layer k must reconstruct what the reference protocol
``decompress(compress(x, q_k), q_k)`` reconstructs (training/step.py:322-337).  Checked (a) exactly against this
package's own per-quality path (itself pinned to the reference by tests/test_gpu_model.py), (b) against the CPU oracle's
per-quality round trip within the north-star PSNR tolerance, (c) each layer segment is a stand-alone stream the CPU
reference coder decodes, (d) rate of a prefix ~ rate of the per-quality streams."""
import numpy as np
import pytest
import torch

from conftest import CASE_KWARGS, build_pair
from oracle.codec_port import psnr
from oracle.gen_golden import synthetic_image

pytestmark = pytest.mark.gpu

LEVELS = (0.05, 0.5, 1.25, 5, 10)


def _total_bytes(strings):
    return sum(len(s) for sl in strings[0] for s in sl) + sum(len(s) for s in strings[1])


def test_prefixes_equal_per_quality_round_trips_small():
    from progressivecodec_b200 import container as C

    net, orc = build_pair("allscalable", "cuda")
    x = synthetic_image((2, 3, 128, 192), seed=21)
    blobs = C.encode_progressive(net, x.cuda(), LEVELS)
    assert len(blobs) == 2 and all(b[:4] == b"PCB2" for b in blobs)
    hdr = C.Header.parse(blobs[0])
    assert hdr.layers_in(len(blobs[0])) == len(LEVELS) and hdr.prefix_end(len(LEVELS)) == len(blobs[0])
    # base only
    base = C.decode_progressive(net, [C.truncate(b, 0) for b in blobs])
    c0 = net.compress(x.cuda(), quality=0)
    r0 = net.decompress(c0["strings"], c0["shape"], quality=0)["x_hat"]
    assert base["n_layers"] == 0 and torch.equal(base["x_hat"], r0)
    prev_len = len(C.truncate(blobs[0], 0))
    for k, q in enumerate(LEVELS, start=1):
        pre = [C.truncate(b, k) for b in blobs]
        assert len(pre[0]) >= prev_len
        prev_len = len(pre[0])
        # a few extra bytes of the next layer must not change anything (only COMPLETE layers are used)
        ragged = [blobs[i][:len(pre[i]) + (3 if k < len(LEVELS) else 0)] for i in range(2)]
        out = C.decode_progressive(net, ragged)
        assert out["n_layers"] == k and abs(out["level"] - q) < 1e-6
        c = net.compress(x.cuda(), quality=q)
        ref = net.decompress(c["strings"], c["shape"], quality=q)["x_hat"]
        assert torch.equal(out["x_hat"], ref), f"prefix {k} (q={q}) differs from the per-quality round trip"
        # CPU oracle of the reference protocol
        oc = orc.compress(x, quality=q)
        orec = orc.decompress(oc["strings"], oc["shape"], quality=q)["x_hat"]
        assert abs(psnr(out["x_hat"].cpu(), x) - psnr(orec, x)) <= 0.02
        # rate: the prefix codes the same in-mask symbols; it saves the (nearly free) masked zeros of the per-quality
        # streams and pays ~8 bytes of rANS flush per (layer, slice)
        per_q = _total_bytes(c["strings"]) / 2
        assert sum(len(p) for p in pre) / 2 <= 1.03 * per_q + 8 * 10 * k + hdr.size, (k, q)
    with pytest.raises(Exception):
        C.decode_progressive(net, [blobs[0][:hdr.prefix_end(0) - 1]])
    with pytest.raises(Exception):
        C.encode_progressive(build_pair("authors", "cuda")[0], x.cuda(), LEVELS)


def test_layer_segments_are_reference_decodable_streams():
    """Each (layer, slice) segment is a plain rANS stream: the CPU restatement of the reference decoder recovers the
    encoder's layer symbols given the layer's indexes."""
    from oracle.entropy_port import CPortCoder
    from progressivecodec_b200 import _lib as L
    from progressivecodec_b200 import ans, container as C

    net, orc = build_pair("allscalable", "cuda")
    torch.manual_seed(5)
    dev = torch.device("cuda", 0)
    B, hw, Cc, nl = 2, 24 * 16, 32, 3
    n = Cc * hw
    from progressivecodec_b200.engine import Act, Engine

    E = Engine(dev)
    sigma = Act(torch.rand(B, 24, 16, Cc, device=dev) * 4 + 0.05)
    thr = torch.tensor([[2.5, 2.0], [1.0, 1.5], [float("-inf")] * 2], dtype=torch.float32, device=dev)
    sym = torch.randint(-6, 7, (B, n), dtype=torch.int32, device=dev)
    idx = torch.randint(0, 64, (B, n), dtype=torch.int32, device=dev)
    csym, cidx = torch.empty_like(sym), torch.empty_like(idx)
    counts = torch.zeros((B, 16), dtype=torch.int32, device=dev)
    E.layer_partition(sigma, thr, sym, idx, csym, cidx, counts)
    # reference partition with torch (stable sort by layer)
    s_nchw = sigma.t.permute(0, 3, 1, 2).reshape(B, n)
    layer = (s_nchw.unsqueeze(0) < thr.unsqueeze(2)).sum(0)
    order = torch.argsort(layer, dim=1, stable=True)
    assert torch.equal(csym, torch.gather(sym, 1, order)) and torch.equal(cidx, torch.gather(idx, 1, order))
    assert torch.equal(counts[:, :nl + 1].long(), torch.stack([(layer == k).sum(1) for k in range(nl + 1)], 1))
    # scatter back with 2 of 3 layers available
    back = torch.empty_like(sym)
    E.layer_partition(sigma, thr, csym, None, back, None, None, avail=torch.full((B,), 2, dtype=torch.int32, device=dev))
    assert torch.equal(back, torch.where(layer < 2, sym, torch.zeros_like(sym)))
    # segments -> bytes -> CPU reference decoder
    t = orc.gc
    tables = net.gaussian_conditional.device_tables(dev)
    cnt = counts[:, :nl].long()
    first = torch.cumsum(cnt, 1) - cnt
    seg_start = (torch.arange(B, device=dev).unsqueeze(1) * n + first).t().contiguous().reshape(-1)
    seg_count = cnt.t().contiguous().reshape(-1).to(torch.int32)
    data, off = ans.encode_segments(csym.reshape(-1), cidx.reshape(-1), seg_start, seg_count, tables, n)
    streams = ans.split_streams(data, off)
    cd, cs, of = t.cdf.numpy(), t.cdf_length.numpy(), t.offset.numpy()
    port = CPortCoder()
    cs_h, ci_h = csym.cpu().numpy().reshape(-1), cidx.cpu().numpy().reshape(-1)
    for s in range(len(streams)):
        a, c = int(seg_start[s]), int(seg_count[s])
        assert streams[s] == port.encode_with_indexes(cs_h[a:a + c], ci_h[a:a + c], cd, cs, of)
    out = torch.zeros_like(csym).reshape(-1)
    offd = off.to(dev)
    ans.decode_segments(data, offd[:-1].contiguous(), offd[1:].contiguous(), seg_start, seg_count, cidx.reshape(-1), out, tables)
    keep = (torch.arange(n, device=dev).unsqueeze(0) < cnt.sum(1, keepdim=True)).reshape(-1)
    assert torch.equal(out[keep], csym.reshape(-1)[keep])


def test_config4_clic_size_single_encode_all_prefixes():
    """BASELINE config 4: one 2048x1365 image (padded to 2048x1408 as training/step.py:317-319 pads), ONE progressive
    encode, decode at every quality prefix by truncation: every prefix equals the per-quality round trip of the same
    model exactly; rate and PSNR grow with the prefix."""
    from progressivecodec_b200 import container as C

    net, _ = build_pair("allscalable", "cuda")
    x0 = synthetic_image((1, 3, 1365, 2048), seed=4)
    x = torch.nn.functional.pad(x0, (0, 0, 21, 22))  # -> 1408 x 2048
    assert x.shape[2:] == (1408, 2048)
    levels = C.DEFAULT_LEVELS
    blob = C.encode_progressive(net, x.cuda(), levels)[0]
    hdr = C.Header.parse(blob)
    sizes, psnrs = [], []
    for k in range(0, len(levels) + 1):
        out = C.decode_progressive(net, [C.truncate(blob, k)])
        xh = out["x_hat"][:, :, 21:21 + 1365, :].cpu()
        sizes.append(hdr.prefix_end(k))
        psnrs.append(psnr(xh, x0))
        if k in (0, 1, 6, len(levels)):
            q = 0 if k == 0 else levels[k - 1]
            c = net.compress(x.cuda(), quality=q)
            ref = net.decompress(c["strings"], c["shape"], quality=q)["x_hat"]
            assert torch.equal(out["x_hat"], ref), (k, q)
    assert all(b >= a for a, b in zip(sizes, sizes[1:]))
    assert sizes[-1] == len(blob)
    assert np.isfinite(psnrs).all()


def test_arbitrary_levels_round_trip_through_the_f32_header():
    """Levels that are not exactly representable in the header's f32 fields (0.3, 1.7, 4.2, 7.77): the encoder rounds
    them to f32 before deriving the mask quantiles, so encoder and decoder compute bit-identical thresholds and every
    prefix still equals the per-quality round trip at the level the header reports."""
    import struct

    from progressivecodec_b200 import container as C

    net, _ = build_pair("allscalable", "cuda")
    x = synthetic_image((1, 3, 128, 192), seed=33)
    levels = (0.3, 1.7, 4.2, 7.77)
    blobs = C.encode_progressive(net, x.cuda(), levels)
    hdr = C.Header.parse(blobs[0])
    f32 = [struct.unpack("<f", struct.pack("<f", v))[0] for v in levels]
    assert hdr.levels == f32
    for k in range(1, len(levels) + 1):
        out = C.decode_progressive(net, [C.truncate(blobs[0], k)])
        q = hdr.levels[k - 1]
        c = net.compress(x.cuda(), quality=q)
        ref = net.decompress(c["strings"], c["shape"], quality=q)["x_hat"]
        assert out["n_layers"] == k and torch.equal(out["x_hat"], ref), (k, q)
