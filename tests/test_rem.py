"""REM wrapper (PostRateProcessedNetwork, CHProgREM.py:205-1126): oracle port vs the real reference's golden vectors
(CPU), and the B200 implementation vs oracle + goldens (GPU)."""
import os

import numpy as np
import pytest
import torch

from conftest import CASE_KWARGS, GOLDEN, load_golden
from oracle.codec_port import CodecConfig, bpp_from_strings, psnr
from oracle.gen_golden import unpack_strings

REM_KW = dict(check_levels=[0.01, 0.25, 1.75], mu_std=False, dimension="big")
QS = [0, 0.1, 1, 5, 10]
VARIANTS = {"rem": (REM_KW, QS), "rem_mustd": (dict(check_levels=[0.05, 1.0], mu_std=True, dimension="middle"), [0.5, 5])}
_CACHE = {}


def build_rem(device=None, name="rem"):
    """(B200 REM model, oracle REM) sharing name-keyed synthetic weights with the golden generator."""
    from oracle.rem_port import OracleREM
    from progressivecodec_b200 import ChannelProgresssiveWACNN, PostRateProcessedNetwork, apply_synthetic_weights

    rem_kw = VARIANTS[name][0]
    if name not in _CACHE:
        kw = CASE_KWARGS["authors"]
        base = ChannelProgresssiveWACNN(**kw).eval()
        apply_synthetic_weights(base, seed=0)
        base.update(force=True)
        rem = PostRateProcessedNetwork(base, **rem_kw).eval()
        apply_synthetic_weights(rem.post_latent, seed=1)
        orc = OracleREM(base.state_dict(), rem.post_latent.state_dict(), CodecConfig(**kw), **rem_kw)
        _CACHE[name] = (rem, orc)
    rem, orc = _CACHE[name]
    if device is not None:
        rem = rem.to(device)
    return rem, orc


def test_rem_module_tree_matches_reference_keys():
    for name in VARIANTS:
        rem, _ = build_rem(name=name)
        want = [l.split(" ", 1)[0] for l in open(os.path.join(GOLDEN, f"{name}_post_latent_keys.txt")).read().splitlines()]
        assert list(rem.post_latent.state_dict().keys()) == want
    rem, _ = build_rem()
    assert all(k.startswith("base_net.") or k.startswith("post_latent.") for k in rem.state_dict().keys())
    assert rem.find_check_quality(0.005) == (0, 0) and rem.find_check_quality(0.1) == (0.01, 0.25)
    assert rem.find_check_quality(1) == (0.25, 1.75) and rem.find_check_quality(5) == (1.75, 10)


@pytest.mark.parametrize("name", ["rem", "rem_mustd"])
def test_rem_oracle_matches_reference_golden(name):
    _rem, orc = build_rem(name=name)
    G = load_golden(name)
    QS = {"rem": [0.1, 5], "rem_mustd": [5]}[name]  # a subset keeps the CPU suite short; the GPU test walks all levels
    x = torch.from_numpy(G["x"])
    npx = x.shape[0] * x.shape[2] * x.shape[3]
    from oracle.codec_port import OracleCodec

    plain = OracleCodec(_rem.base_net.state_dict(), CodecConfig(**CASE_KWARGS["authors"]))
    changed = False
    for q in QS:
        ref = unpack_strings(G, f"q{q}_")
        out = orc.compress(x, quality=q, mask_pol="point-based-std")
        if out["strings"][0] != ref[0] or out["strings"][1] != ref[1]:
            assert abs(bpp_from_strings(out["strings"], npx) - bpp_from_strings(ref, npx)) <= 0.005 * bpp_from_strings(ref, npx)
        rec = orc.decompress(ref, tuple(G[f"q{q}_shape"]), quality=q, mask_pol="point-based-std")["x_hat"]
        ref_x = torch.from_numpy(G[f"q{q}_x_hat"])
        if not torch.equal(rec, ref_x):
            assert abs(psnr(rec, x) - psnr(ref_x, x)) <= 0.02
        if q == 5:  # the refinement must actually change the streams w.r.t. the plain base codec
            changed |= plain.compress(x, quality=q, mask_pol="point-based-std")["strings"][0] != ref[0]
    assert changed


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["rem", "rem_mustd"])
def test_rem_gpu_vs_oracle_and_golden(name):
    rem, orc = build_rem("cuda", name)
    G = load_golden(name)
    QS = VARIANTS[name][1]
    x = torch.from_numpy(G["x"])
    for q in QS:
        dbg, odbg = {}, {}
        out = rem.compress(x.cuda(), quality=q, mask_pol="point-based-std", debug=dbg)
        o = orc.compress(x, quality=q, mask_pol="point-based-std", debug=odbg)
        sym, idx = dbg["symbols"].cpu(), dbg["indexes"].cpu()
        first = None
        for s in range(sym.shape[0]):
            bad = (sym[s] != odbg["symbols"][s].reshape(1, -1)) | (idx[s] != odbg["indexes"][s].reshape(1, -1))
            if bad.any():
                first = s
                assert float(bad.float().mean()) * sym.shape[2] <= 1.5, (q, s)
                break
        nb = sum(len(s) for sl in out["strings"][0] for s in sl)
        nb_ref = sum(len(s) for sl in o["strings"][0] for s in sl)
        assert abs(nb - nb_ref) <= 0.005 * nb_ref + 8, (q, nb, nb_ref)
        if first is None:
            assert out["strings"][0] == o["strings"][0] and out["strings"][1] == o["strings"][1]
        dec = rem.decompress(out["strings"], out["shape"], quality=q, mask_pol="point-based-std")
        assert abs(psnr(dec["x_hat"].cpu(), x) - psnr(torch.from_numpy(G[f"q{q}_x_hat"]), x)) <= 0.02, q
        # the decoder reproduces the encoder-side latent exactly (same kernels, same order)
        y_dec = torch.cat(dec["y_hat"], 1) if q == 0 else dec["y_hat"]
        assert torch.equal(y_dec, out["y_hat"]), q
        if q in (0, 1):  # real_compress=False (CHProgREM.py:857-860): same latent, no entropy coder
            nc = rem.compress(x.cuda(), quality=q, mask_pol="point-based-std", real_compress=False)
            assert nc["strings"] is None and torch.equal(nc["y_hat"], out["y_hat"]), q
        if f"q{q}_mask_sum" in G.files and first is None:
            got = np.array([float(m.sum()) for m in out["masks"]])
            assert np.abs(got - G[f"q{q}_mask_sum"]).max() <= 2
    # batch > 8 exercises the decode groups (threads / streams) with the refinement nets
    xb = x.repeat(9, 1, 1, 1).cuda()
    ob = rem.compress(xb, quality=1, mask_pol="point-based-std")
    rb = rem.decompress(ob["strings"], ob["shape"], quality=1, mask_pol="point-based-std")["x_hat"]
    r1 = rem.decompress([[s[:1] for s in ob["strings"][0]], ob["strings"][1][:1]], ob["shape"], quality=1,
                        mask_pol="point-based-std")["x_hat"]
    assert torch.equal(rb[0], r1[0]) and torch.equal(rb[8], r1[0])
    with pytest.raises(ValueError):
        rem.compress(x.cuda(), quality=1, checkpoint_rep=out["y_hat"][:, :32])


def test_rem_escalation_oracle_matches_reference_golden():
    """checkpoint_rep path (CHProgREM.py:336-372, :773, :989): golden = the real reference with escalation=True."""
    _rem, orc = build_rem()
    G = load_golden("rem_escalation")
    x = torch.from_numpy(G["x"])
    cl = REM_KW["check_levels"]
    r = orc.compress(x, quality=cl[0], mask_pol="point-based-std")["y_hat"]
    assert torch.equal(r, torch.from_numpy(G["rep0"]))
    r = orc.compress(x, quality=cl[1], mask_pol="point-based-std", checkpoint_rep=r)["y_hat"]
    assert torch.equal(r, torch.from_numpy(G["rep1"]))
    rep2 = torch.from_numpy(G["rep2"])
    ref = unpack_strings(G, "q5_")
    d = orc.decompress(ref, tuple(G["q5_shape"]), quality=5, mask_pol="point-based-std", checkpoint_rep=rep2)
    assert torch.equal(d["x_hat"], torch.from_numpy(G["q5_x_hat"]))
    # the representation matters: without it the same strings decode to something else
    d0 = orc.decompress(ref, tuple(G["q5_shape"]), quality=5, mask_pol="point-based-std")
    assert not torch.equal(d0["x_hat"], d["x_hat"])


@pytest.mark.gpu
def test_rem_escalation_gpu_vs_golden():
    from progressivecodec_b200 import PostRateProcessedNetwork

    rem0, orc = build_rem("cuda")
    rem = PostRateProcessedNetwork(rem0.base_net, **dict(REM_KW, escalation=True)).eval()
    rem.post_latent.load_state_dict(rem0.post_latent.state_dict())
    rem = rem.cuda()
    G = load_golden("rem_escalation")
    x = torch.from_numpy(G["x"])
    cl = REM_KW["check_levels"]
    npx = x.shape[0] * x.shape[2] * x.shape[3]
    for k, q in enumerate(cl):  # the chained representations: latents agree with the reference's up to rounding-tie flips
        r = rem.extract_chekpoint_representation_from_images(x.cuda(), q).cpu()
        ref = torch.from_numpy(G[f"rep{k}"])
        assert r.shape == ref.shape
        assert float(((r - ref).abs() > 0.5).float().mean()) <= 1e-3, (q, float((r - ref).abs().max()))
    rep2 = torch.from_numpy(G["rep2"]).cuda()
    out = rem.compress(x.cuda(), quality=5, mask_pol="point-based-std", checkpoint_rep=rep2)
    ref = unpack_strings(G, "q5_")
    assert abs(bpp_from_strings(out["strings"], npx) - bpp_from_strings(ref, npx)) <= 0.005 * bpp_from_strings(ref, npx)
    dec = rem.decompress(out["strings"], out["shape"], quality=5, mask_pol="point-based-std", checkpoint_rep=rep2)
    assert abs(psnr(dec["x_hat"].cpu(), x) - psnr(torch.from_numpy(G["q5_x_hat"]), x)) <= 0.02
    assert torch.equal(dec["y_hat"], out["y_hat"])
    # and the representation is really used
    d0 = rem.decompress(out["strings"], out["shape"], quality=5, mask_pol="point-based-std")
    assert not torch.equal(d0["x_hat"], dec["x_hat"])
