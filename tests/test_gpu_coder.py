"""CUDA rANS coder vs the oracle (bit-exact, both directions), through the C-ABI."""
import numpy as np
import pytest
import torch

from oracle import build_ref, entropy_port as EP

pytestmark = pytest.mark.gpu

CDF4 = [[0, 16384, 32768, 49152, 65536]]
KATS = [
    ([-1, 0, 1, 0, 0, 1, 1, -1], "0000590a00800000"),
    ([-1, 0, 1, 0, 0, 5, -7, 0], "03040080000000006171d602"),
    ([0] * 8, "0040551500800000"),
    ([0, 0, 0, 40000, 0, 0, 0, 0], "45553500800000007c380100"),
    ([0, 1, 1, 0, -1, 3, 0, 0], "21401a0314008000"),
]


@pytest.fixture(scope="module")
def tables():
    from progressivecodec_b200 import ans

    t = EP.GaussianTables.build()
    return t, ans.CdfTables(t.cdf, t.cdf_length, t.offset)


def _stream(rng, t, n, heavy):
    idx = rng.integers(0, t.cdf.shape[0], size=n).astype(np.int32)
    sigma = t.scale_table.numpy()[idx]
    sym = np.rint(rng.standard_normal(n) * sigma * (4.0 if heavy else 1.0)).astype(np.int32)
    return sym, idx


@pytest.mark.parametrize("symbols,hexstr", KATS)
def test_known_answers_python_api(symbols, hexstr):
    """Same call shape as the reference's python call sites (entropy_models.py:227-235, 276-286)."""
    from progressivecodec_b200 import ans

    enc = ans.RansEncoder().encode_with_indexes(symbols, [0] * len(symbols), CDF4, [5], [-1])
    assert enc.hex() == hexstr
    assert ans.RansDecoder().decode_with_indexes(enc, [0] * len(symbols), CDF4, [5], [-1]) == symbols


def test_buffered_encoder_and_decode_stream():
    from progressivecodec_b200 import ans

    be = ans.BufferedRansEncoder()
    be.encode_with_indexes([0, 1, 1, 0], [0] * 4, CDF4, [5], [-1])
    be.encode_with_indexes([-1, 3, 0, 0], [0] * 4, CDF4, [5], [-1])
    s = be.flush()
    assert s.hex() == "21401a0314008000"
    d = ans.RansDecoder()
    d.set_stream(s)
    assert d.decode_stream([0] * 4, CDF4, [5], [-1]) == [0, 1, 1, 0]
    assert d.decode_stream([0] * 4, CDF4, [5], [-1]) == [-1, 3, 0, 0]


def test_short_and_empty_streams():
    from progressivecodec_b200 import ans

    port = EP.CPortCoder()
    for syms in ([], [0], [1, -1], [3, 0, -9]):
        enc = ans.RansEncoder().encode_with_indexes(syms, [0] * len(syms), CDF4, [5], [-1])
        assert enc == port.encode_with_indexes(syms, [0] * len(syms), CDF4, [5], [-1])
        assert ans.RansDecoder().decode_with_indexes(enc, [0] * len(syms), CDF4, [5], [-1]) == syms


@pytest.mark.parametrize("n,heavy", [(1, False), (31, True), (32, False), (33, True), (1000, False), (49152, False),
                                      (49152, True), (18432, False)])
def test_batch_bit_exact_and_cross_decode(tables, n, heavy):
    from progressivecodec_b200 import ans

    t, dev_tables = tables
    cd, cs, of = t.cdf.numpy(), t.cdf_length.numpy(), t.offset.numpy()
    rng = np.random.default_rng(n + heavy)
    S = 5
    streams = [_stream(rng, t, n, heavy) for _ in range(S)]
    sym = torch.tensor(np.stack([s for s, _ in streams]), device="cuda")
    idx = torch.tensor(np.stack([i for _, i in streams]), device="cuda")
    data, offs = ans.encode_batch(sym, idx, dev_tables)
    got = ans.split_streams(data, offs)
    coders = [EP.CPortCoder()] + ([EP.RefCoder()] if build_ref.ref_built() and n >= 8 else [])
    for (s, i), g in zip(streams, got):
        for c in coders:
            assert g == c.encode_with_indexes(s, i, cd, cs, of), c.name      # GPU bytes == CPU bytes
            assert (c.decode_with_indexes(g, i, cd, cs, of) == s).all()       # CPU decodes GPU stream
    # GPU decodes CPU streams (independent encoder)
    cpu = [coders[-1].encode_with_indexes(s, i, cd, cs, of) for s, i in streams]
    blob, o = ans.pack_streams(cpu, sym.device)
    out = ans.decode_batch(blob, o, idx, dev_tables)
    assert torch.equal(out.cpu(), sym.cpu())


def test_masked_progressive_like_streams(tables):
    """Mostly (symbol 0, table 0) with a few live elements — the low-quality progressive slices."""
    from progressivecodec_b200 import ans

    t, dev_tables = tables
    cd, cs, of = t.cdf.numpy(), t.cdf_length.numpy(), t.offset.numpy()
    rng = np.random.default_rng(7)
    n = 49152
    sym, idx = _stream(rng, t, n, False)
    keep = rng.random(n) < 0.005
    sym, idx = np.where(keep, sym, 0).astype(np.int32), np.where(keep, idx, 0).astype(np.int32)
    data, offs = ans.encode_batch(torch.tensor(sym[None], device="cuda"), torch.tensor(idx[None], device="cuda"), dev_tables)
    g = ans.split_streams(data, offs)[0]
    assert g == EP.CPortCoder().encode_with_indexes(sym, idx, cd, cs, of)
    out = ans.decode_batch(data, offs, torch.tensor(idx[None], device="cuda"), dev_tables)
    assert (out.cpu().numpy()[0] == sym).all()


def test_full_size_round_trip_checksum(tables):
    """BASELINE config sizes: 21 streams x 8 images of 49152 symbols: encode -> decode is the identity and the
    per-stream lengths equal the oracle's on a sample of streams (size-independent property + spot check)."""
    from progressivecodec_b200 import ans

    t, dev_tables = tables
    g = torch.Generator(device="cuda").manual_seed(0)
    S, n = 168, 49152
    idx = torch.randint(0, 64, (S, n), generator=g, device="cuda", dtype=torch.int32)
    sigma = t.scale_table.cuda()[idx.long()]
    sym = torch.round(torch.randn((S, n), generator=g, device="cuda") * sigma).int()
    data, offs = ans.encode_batch(sym, idx, dev_tables)
    out = ans.decode_batch(data, offs, idx, dev_tables)
    assert torch.equal(out, sym)
    cd, cs, of = t.cdf.numpy(), t.cdf_length.numpy(), t.offset.numpy()
    port = EP.CPortCoder()
    streams = ans.split_streams(data, offs)
    for s in (0, 77, 167):
        assert streams[s] == port.encode_with_indexes(sym[s].cpu().numpy(), idx[s].cpu().numpy(), cd, cs, of)


def test_entropy_model_api_matches_oracle(tables):
    """GaussianConditional.compress/decompress/build_indexes (reference API, NCHW tensors)."""
    from progressivecodec_b200 import GaussianConditional
    from progressivecodec_b200.models import get_scale_table

    t, _ = tables
    gc = GaussianConditional(None)
    gc.update_scale_table(get_scale_table())
    assert torch.equal(gc._quantized_cdf, t.cdf) and torch.equal(gc._cdf_length, t.cdf_length)
    gc = gc.cuda()
    g = torch.Generator().manual_seed(5)
    scales = torch.rand(2, 32, 8, 12, generator=g) * 3 - 0.2
    means = torch.randn(2, 32, 8, 12, generator=g)
    y = means + torch.randn(2, 32, 8, 12, generator=g) * scales.clamp_min(0.11)
    idx = gc.build_indexes(scales.cuda())
    assert torch.equal(idx.cpu(), t.build_indexes(scales))
    strings = gc.compress(y.cuda(), idx, means.cuda())
    sym = torch.round(y - means).int()
    assert strings == t.encode(sym, idx.cpu(), EP.CPortCoder())
    y_hat = gc.decompress(strings, idx, means.cuda())
    assert torch.equal(y_hat.cpu(), sym.float() + means)
    out, lik = gc(y.cuda(), scales.cuda(), means.cuda(), training=False)
    ref_lik = t.likelihood(y, scales, means)
    assert torch.allclose(lik.cpu(), ref_lik, rtol=2e-4, atol=1e-7)
    assert torch.equal(out.cpu(), torch.round(y - means) + means)
