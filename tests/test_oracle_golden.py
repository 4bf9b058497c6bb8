"""Pin the oracle port (oracle/codec_port.py, oracle/entropy_port.py) against golden vectors produced by the REAL
reference (oracle/gen_golden.py, run in the build container where /root/reference is mounted)."""
import numpy as np
import pytest
import torch

from conftest import CASE_KWARGS, build_pair, load_golden
from oracle import entropy_port as EP
from oracle.codec_port import bpp_from_strings, psnr
from oracle.gen_golden import unpack_strings

torch.set_num_threads(8)


def test_entropy_tables_match_reference_update():
    G = load_golden("authors")
    t = EP.GaussianTables.build()
    assert np.array_equal(t.cdf.numpy(), G["gc_cdf"])
    assert np.array_equal(t.cdf_length.numpy(), G["gc_cdf_length"])
    assert np.array_equal(t.offset.numpy(), G["gc_offset"])
    assert np.array_equal(t.scale_table.numpy(), G["gc_scale_table"])
    _net, orc = build_pair("authors")
    eb = orc.eb
    assert np.array_equal(eb.cdf.numpy(), G["eb_cdf"])
    eb.rebuild()  # the port's own table builder on the same parameters
    assert np.array_equal(eb.cdf.numpy(), G["eb_cdf"])
    assert np.array_equal(eb.cdf_length.numpy(), G["eb_cdf_length"])
    assert np.array_equal(eb.offset.numpy(), G["eb_offset"])


def test_product_update_builds_the_reference_tables():
    """ChannelProgresssiveWACNN.update() (host set-up through the C-ABI pmf_to_quantized_cdf) == reference update()."""
    net, _ = build_pair("authors")
    G = load_golden("authors")
    gc, eb = net.gaussian_conditional, net.entropy_bottleneck
    assert np.array_equal(gc._quantized_cdf.cpu().numpy(), G["gc_cdf"])
    assert np.array_equal(gc._cdf_length.cpu().numpy(), G["gc_cdf_length"])
    assert np.array_equal(gc._offset.cpu().numpy(), G["gc_offset"])
    assert np.array_equal(eb._quantized_cdf.cpu().numpy(), G["eb_cdf"])
    assert np.array_equal(eb._cdf_length.cpu().numpy(), G["eb_cdf_length"])
    assert np.array_equal(eb._offset.cpu().numpy(), G["eb_offset"])


def _same_or_close(a, b, what):
    """Bit-identical on the CPU the goldens were generated on; a different host CPU may pick other oneDNN kernels,
    so fall back to a tight numeric bound."""
    if torch.equal(a, b):
        return True
    assert torch.allclose(a, b, rtol=1e-3, atol=2e-3), what
    return False


@pytest.mark.parametrize("case", ["authors", "plain"])
def test_compress_decompress_match_reference(case):
    _net, orc = build_pair(case)
    G = load_golden(case)
    x = torch.from_numpy(G["x"])
    pol = CASE_KWARGS[case]["mask_policy"]
    npx = x.shape[0] * x.shape[2] * x.shape[3]
    for q in ([0, 0.5, 10] if pol != "two-levels" else [0, 10]):
        ref = unpack_strings(G, f"q{q}_")
        out = orc.compress(x, quality=q, mask_pol=pol)
        if out["strings"][0] != ref[0] or out["strings"][1] != ref[1]:
            assert abs(bpp_from_strings(out["strings"], npx) - bpp_from_strings(ref, npx)) <= 0.005 * bpp_from_strings(ref, npx)
        rec = orc.decompress(ref, tuple(G[f"q{q}_shape"]), quality=q, mask_pol=pol)["x_hat"]
        ref_x = torch.from_numpy(G[f"q{q}_x_hat"])
        if not torch.equal(rec, ref_x):
            assert abs(psnr(rec, x) - psnr(ref_x, x)) <= 0.02


@pytest.mark.parametrize("case", ["multienc", "allscalable"])
def test_forward_paths_match_reference(case):
    _net, orc = build_pair(case)
    G = load_golden(case)
    x = torch.from_numpy(G["x"])
    pol = CASE_KWARGS[case]["mask_policy"]
    o = orc.forward_single_quality(x, 5, mask_pol=pol)
    _same_or_close(o["x_hat"], torch.from_numpy(G["fsq5_x_hat"]), "fsq x_hat")
    _same_or_close(o["likelihoods"]["y"], torch.from_numpy(G["fsq5_lik_y"]), "fsq lik_y")
    _same_or_close(o["likelihoods"]["z"], torch.from_numpy(G["fsq5_lik_z"]), "fsq lik_z")
    ql = [int(v) if v == int(v) else float(v) for v in G["fwd_qualities"]]
    o = orc.forward(x, quality=ql, mask_pol=pol)
    _same_or_close(o["x_hat"], torch.from_numpy(G["fwd_x_hat"]), "forward x_hat")
    _same_or_close(o["likelihoods"]["y_prog"], torch.from_numpy(G["fwd_lik_y_prog"]), "forward lik_y_prog")


def test_quantile_restatement_matches_torch():
    """sort / fp32 rank / lerp restatement reproduces torch.quantile bit for bit."""
    g = torch.Generator().manual_seed(0)
    for n in (1024, 8192, 49152):
        v = torch.randn(n, generator=g) * 0.5 + 0.2
        for pr in (0.05, 0.25, 0.6, 1, 1.25, 3, 5, 9.5):
            q = 1.0 - pr * 0.1
            assert EP.quantile_threshold_np(v.numpy(), q) == torch.quantile(v, q).item()
    s = torch.randn(2, 32, 8, 12, generator=g)
    m = EP.point_based_std_mask_np(s.numpy(), 2.5)
    ref = torch.stack([(s[b] >= torch.quantile(s[b].reshape(-1), 0.75)).float() for b in range(2)])
    assert np.array_equal(m, ref.numpy())


def test_cust_map_masks_match_reference():
    """cust_map path (masking.py:171-194 via CHProg_cnn.py:721,823,850,964): oracle vs the real reference's streams and
    reconstruction (tests/golden/authors_custmap.npz, oracle/gen_golden.py --cases custmap)."""
    _net, orc = build_pair("authors")
    G = load_golden("authors_custmap")
    x, cm = torch.from_numpy(G["x"]), torch.from_numpy(G["cust_map"])
    npx = x.shape[0] * x.shape[2] * x.shape[3]
    for q in (0.5, 5):
        ref = unpack_strings(G, f"q{q}_")
        out = orc.compress(x, quality=q, mask_pol="point-based-std", cust_map=cm)
        assert np.abs(np.array([float(m.sum()) for m in out["masks"]]) - G[f"q{q}_mask_sum"]).max() == 0
        if out["strings"][0] != ref[0] or out["strings"][1] != ref[1]:
            assert abs(bpp_from_strings(out["strings"], npx) - bpp_from_strings(ref, npx)) <= 0.005 * bpp_from_strings(ref, npx)
        rec = orc.decompress(ref, tuple(G[f"q{q}_shape"]), quality=q, mask_pol="point-based-std", cust_map=cm)["x_hat"]
        ref_x = torch.from_numpy(G[f"q{q}_x_hat"])
        if not torch.equal(rec, ref_x):
            assert abs(psnr(rec, x) - psnr(ref_x, x)) <= 0.02


def test_headline_shape_oracle_matches_the_real_reference():
    """The shape BASELINE.json's metric is quoted on: one 768x512 image, authors' flags, q = 0 and 5.  The fixture
    (tests/golden/headline_768x512.npz, `python -m oracle.gen_golden --cases headline`) keeps length + crc32 of every
    stream the REAL reference wrote, the crc32 / PSNR / a pooled copy of its reconstruction; the oracle must reproduce
    them (bit for bit on the generating CPU; within the north star's bars where another CPU's kernels round differently)."""
    import zlib

    from oracle.gen_golden import HEADLINE_QUALITIES, HEADLINE_SEED, HEADLINE_SHAPE, headline_digest, synthetic_image

    _net, orc = build_pair("authors")
    G = load_golden("headline_768x512")
    x = synthetic_image(HEADLINE_SHAPE, seed=HEADLINE_SEED)
    for q in HEADLINE_QUALITIES:
        c = orc.compress(x, quality=q, mask_pol="point-based-std")
        rec = orc.decompress(c["strings"], tuple(c["shape"]), quality=q, mask_pol="point-based-std")["x_hat"]
        d = headline_digest(c["strings"], rec, x)
        n_streams = (10 if q == 0 else 20) + 1
        assert len(G[f"q{q}_lens"]) == n_streams == len(d["lens"])
        if np.array_equal(d["crcs"], G[f"q{q}_crcs"]):          # same bytes in => the decoder is deterministic
            assert np.array_equal(d["lens"], G[f"q{q}_lens"])
            if int(d["x_hat_crc"]) != int(G[f"q{q}_x_hat_crc"]):
                assert np.abs(d["x_hat_pooled"] - G[f"q{q}_x_hat_pooled"]).max() <= 1e-4
        else:
            ref_bytes = int(G[f"q{q}_lens"].sum())
            assert abs(int(d["lens"].sum()) - ref_bytes) <= 0.005 * ref_bytes
        assert abs(float(d["psnr"]) - float(G[f"q{q}_psnr"])) <= 0.02
        assert zlib.crc32(b"") == 0  # (crc convention of the fixture: plain zlib.crc32 of each stream)


def test_config3_batched_forward_oracle_matches_the_real_reference():
    """BASELINE.json configs[2]: forward() on 16x3x256x256 at all 13 levels with variance-aware masking.  The fixture
    (tests/golden/config3_forward_16x256x256.npz, `python -m oracle.gen_golden --cases config3`) holds, per level, the
    crc32 and PSNR of the REAL reference's reconstruction and the rate its likelihoods imply."""
    from oracle.gen_golden import CONFIG3_SEED, CONFIG3_SHAPE, SWEEP_LEVELS, forward_digest, synthetic_image

    _net, orc = build_pair("authors")
    G = load_golden("config3_forward_16x256x256")
    x = synthetic_image(CONFIG3_SHAPE, seed=CONFIG3_SEED)
    d = forward_digest(orc.forward(x, quality=SWEEP_LEVELS, mask_pol="point-based-std"), x)
    assert len(d["psnr"]) == len(SWEEP_LEVELS) == len(G["psnr"]) and len(d["bpp_y_prog"]) == len(G["bpp_y_prog"])
    if not (np.array_equal(d["x_hat_crc"], G["x_hat_crc"]) and np.array_equal(d["lik_crc"], G["lik_crc"])):
        assert np.abs(d["psnr"] - G["psnr"]).max() <= 0.02          # another CPU's kernels: the north star's bars
        assert np.abs(d["bpp_y_prog"] - G["bpp_y_prog"]).max() <= 0.005 * G["bpp_y_prog"].max()
    assert abs(float(d["bpp_y"]) - float(G["bpp_y"])) <= 0.005 * float(G["bpp_y"])
    assert abs(float(d["bpp_z"]) - float(G["bpp_z"])) <= 0.005 * float(G["bpp_z"])
    assert np.all(np.diff(G["bpp_y_prog"]) >= 0)                    # the rate grows with the quality level


def test_config4_clic_shape_oracle_matches_the_real_reference():
    """BASELINE.json configs[3] shape (2048x1365 padded to 2048x1408, slices of 360 448 elements): the oracle against the
    REAL reference's stream lengths / crc32s and reconstruction at q = 5 (tests/golden/config4_2048x1408.npz)."""
    from oracle.gen_golden import config4_image, headline_digest

    _net, orc = build_pair("authors")
    G = load_golden("config4_2048x1408")
    x = config4_image()
    q = 5
    c = orc.compress(x, quality=q, mask_pol="point-based-std")
    rec = orc.decompress(c["strings"], tuple(c["shape"]), quality=q, mask_pol="point-based-std")["x_hat"]
    d = headline_digest(c["strings"], rec, x)
    assert len(d["lens"]) == len(G[f"q{q}_lens"]) == 21
    ref_bytes = int(G[f"q{q}_lens"].sum())
    if np.array_equal(d["crcs"], G[f"q{q}_crcs"]):
        assert np.array_equal(d["lens"], G[f"q{q}_lens"])
    else:
        assert abs(int(d["lens"].sum()) - ref_bytes) <= 0.005 * ref_bytes
    assert abs(float(d["psnr"]) - float(G[f"q{q}_psnr"])) <= 0.02


def build_table800_pair(device=None):
    """Authors' flags with the reference's 800-level scale table (CHProg_cnn.py:16-26) passed to update()."""
    from conftest import CASE_KWARGS
    from oracle.codec_port import CodecConfig, OracleCodec
    from progressivecodec_b200 import ChannelProgresssiveWACNN, apply_synthetic_weights, get_scale_table

    kw = CASE_KWARGS["authors"]
    net = ChannelProgresssiveWACNN(**kw).eval()
    apply_synthetic_weights(net, seed=0)
    net.update(scale_table=get_scale_table(0.04, 256, 800), force=True)
    orc = OracleCodec(net.state_dict(), CodecConfig(**kw))
    return (net.to(device) if device is not None else net), orc


def test_800_level_scale_table_matches_reference():
    """update(scale_table=<800 levels>) (the table CHProg_cnn.py:16-26 defines): our table construction + the oracle
    reproduce the real reference's streams (tests/golden/authors_table800.npz, oracle/gen_golden.py --cases table800)."""
    net, orc = build_table800_pair()
    G = load_golden("authors_table800")
    assert np.array_equal(net.gaussian_conditional.scale_table.numpy(), G["scale_table"])
    assert tuple(net.gaussian_conditional._quantized_cdf.shape)[0] == 800
    x = torch.from_numpy(G["x"])
    for q in (0, 5):
        ref = unpack_strings(G, f"q{q}_")
        out = orc.compress(x, quality=q, mask_pol="point-based-std")
        assert out["strings"][0] == ref[0] and out["strings"][1] == ref[1], q
        rec = orc.decompress(ref, tuple(G[f"q{q}_shape"]), quality=q, mask_pol="point-based-std")["x_hat"]
        assert torch.equal(rec, torch.from_numpy(G[f"q{q}_x_hat"])), q


def test_oracle_side_synthetic_state_dict_equals_the_products():
    """bench.py's reference arm / cpu_baseline build their model from oracle/synthetic_state.py (no product package, no
    libpcodec_b200.so in the process).  It must be the SAME model the CUDA arm runs: every entry bit-identical to
    the product's state_dict() after apply_synthetic_weights() + update(), in the same order."""
    from conftest import build_pair
    from oracle.synthetic_state import synthetic_image, synthetic_state_dict
    from progressivecodec_b200.synthetic import synthetic_image as product_image

    net, _ = build_pair("authors")
    ref = net.state_dict()
    sd = synthetic_state_dict("authors", seed=0)
    assert list(sd.keys()) == list(ref.keys())
    for k, v in ref.items():
        assert sd[k].dtype == v.dtype and sd[k].shape == v.shape and torch.equal(sd[k], v.cpu()), k
    assert torch.equal(synthetic_image((1, 3, 64, 128), 3), product_image((1, 3, 64, 128), 3))
