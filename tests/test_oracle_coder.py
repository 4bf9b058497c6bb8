"""Pin the oracle coder (C port) and the C-ABI host pieces against the reference's known answers.

Known-answer vectors: SURVEY.md §4 (produced there by the reference's own rans_interface.cpp / ops.cpp compiled
unmodified).  When oracle/_ref is built (always in the build container, and it travels to the GPU box) the C port
is additionally cross-checked against the compiled reference on random streams in both directions.
"""
import ctypes

import numpy as np
import pytest

from oracle import build_ref, entropy_port as EP

CDF4 = [[0, 16384, 32768, 49152, 65536]]
KATS = [
    ([-1, 0, 1, 0, 0, 1, 1, -1], "0000590a00800000"),
    ([-1, 0, 1, 0, 0, 5, -7, 0], "03040080000000006171d602"),
    ([0] * 8, "0040551500800000"),
    ([0, 0, 0, 40000, 0, 0, 0, 0], "45553500800000007c380100"),
    ([0, 1, 1, 0, -1, 3, 0, 0], "21401a0314008000"),
]


def _coders():
    out = [EP.CPortCoder()]
    if build_ref.ref_built() or build_ref.reference_available():
        out.append(EP.RefCoder())
    return out


@pytest.mark.parametrize("symbols,hexstr", KATS)
def test_known_answers(symbols, hexstr):
    for coder in _coders():
        enc = coder.encode_with_indexes(symbols, [0] * len(symbols), CDF4, [5], [-1])
        assert enc.hex() == hexstr, coder.name
        dec = coder.decode_with_indexes(enc, [0] * len(symbols), CDF4, [5], [-1])
        assert list(dec) == symbols, coder.name


def test_pmf_to_quantized_cdf_known_answers():
    assert EP.pmf_to_quantized_cdf([0.1, 0.2, 0.7]).tolist() == [0, 6554, 19661, 65536]
    assert EP.pmf_to_quantized_cdf([0.5, 0.5, 1e-9, 1e-9]).tolist() == [0, 32766, 65534, 65535, 65536]
    if build_ref.ref_built():
        _ans, cxx = build_ref.import_ref_coder()
        rng = np.random.default_rng(0)
        for n in (2, 5, 33, 300):
            p = rng.random(n).astype(np.float32) ** 4
            p /= p.sum()
            assert EP.pmf_to_quantized_cdf(p).tolist() == list(cxx.pmf_to_quantized_cdf(p.tolist(), 16))


def test_gaussian_table_build_matches_known_values():
    t = EP.GaussianTables.build()
    assert tuple(t.cdf.shape) == (64, 3133)
    assert t.cdf_length[:10].tolist() == [5, 5, 5, 5, 7, 7, 7, 7, 7, 9]
    assert t.offset[:5].tolist() == [-1, -1, -1, -1, -2]
    assert int(t.cdf_length.sum()) == 27256


def _random_stream(rng, tables, n, heavy=False):
    idx = rng.integers(0, tables.cdf.shape[0], size=n).astype(np.int32)
    sigma = tables.scale_table.numpy()[idx]
    sym = np.rint(rng.standard_normal(n) * sigma * (4.0 if heavy else 1.0)).astype(np.int32)
    return sym, idx


def test_c_port_matches_reference_on_random_streams():
    if not build_ref.ref_built():
        pytest.skip("oracle/_ref not built")
    t = EP.GaussianTables.build()
    cd, cs, of = t.cdf.numpy(), t.cdf_length.numpy(), t.offset.numpy()
    rng = np.random.default_rng(1)
    ref, port = EP.RefCoder(), EP.CPortCoder()
    for n, heavy in ((8, False), (100, True), (4097, False), (3000, True)):
        sym, idx = _random_stream(rng, t, n, heavy)
        a = ref.encode_with_indexes(sym, idx, cd, cs, of)
        b = port.encode_with_indexes(sym, idx, cd, cs, of)
        assert a == b
        assert (port.decode_with_indexes(a, idx, cd, cs, of) == sym).all()
        assert (ref.decode_with_indexes(b, idx, cd, cs, of) == sym).all()


def test_c_port_short_and_empty_streams_are_safe():
    """The reference corrupts the heap for < 4 symbols (rans_interface.cpp:170); the port must not."""
    port = EP.CPortCoder()
    for syms in ([], [0], [1, -1], [3, 0, -9]):
        enc = port.encode_with_indexes(syms, [0] * len(syms), CDF4, [5], [-1])
        assert len(enc) % 4 == 0 and len(enc) >= 8
        assert list(port.decode_with_indexes(enc, [0] * len(syms), CDF4, [5], [-1])) == syms
    assert port.encode_with_indexes([], [], CDF4, [5], [-1]).hex() == "0000008000000000"


def test_cabi_core_arithmetic_matches_port():
    """csrc/rans_core.h (shared by the CUDA kernels) walked on the host == oracle bytes."""
    from progressivecodec_b200 import _lib

    lib = _lib.lib()
    t = EP.GaussianTables.build()
    cd = np.ascontiguousarray(t.cdf.numpy(), dtype=np.int32)
    cs = np.ascontiguousarray(t.cdf_length.numpy(), dtype=np.int32)
    of = np.ascontiguousarray(t.offset.numpy(), dtype=np.int32)
    rng = np.random.default_rng(2)
    port = EP.CPortCoder()
    for n, heavy in ((1, False), (9, True), (5000, False), (5000, True)):
        sym, idx = _random_stream(rng, t, n, heavy)
        cap = 6 * n + 16
        words = np.zeros(cap, dtype=np.uint32)
        used = lib.pcodec_selftest_rans_core_encode(sym.ctypes.data, idx.ctypes.data, n, cd.ctypes.data, cd.shape[1],
                                                    cs.ctypes.data, of.ctypes.data, words.ctypes.data, cap)
        assert used > 0
        assert words[cap - used:].tobytes() == port.encode_with_indexes(sym, idx, cd, cs, of)


def test_cabi_pmf_to_quantized_cdf():
    from progressivecodec_b200 import pmf_to_quantized_cdf
    import torch

    assert pmf_to_quantized_cdf(torch.tensor([0.1, 0.2, 0.7])).tolist() == [0, 6554, 19661, 65536]
    assert pmf_to_quantized_cdf(torch.tensor([0.5, 0.5, 1e-9, 1e-9])).tolist() == [0, 32766, 65534, 65535, 65536]
    rng = np.random.default_rng(3)
    for n in (3, 17, 257, 3000):
        p = rng.random(n).astype(np.float32) ** 6
        p /= p.sum()
        assert pmf_to_quantized_cdf(torch.from_numpy(p)).tolist() == EP.pmf_to_quantized_cdf(p).tolist()
