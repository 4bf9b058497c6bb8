"""Masking / quantisation / index kernels vs the oracle on identical inputs (exact for integer outputs)."""
import numpy as np
import pytest
import torch

from oracle import entropy_port as EP

pytestmark = pytest.mark.gpu


def _engine():
    from progressivecodec_b200.engine import Engine

    return Engine(torch.device("cuda", 0))


@pytest.mark.parametrize("hw", [(4, 8), (16, 16), (32, 48), (88, 128)])
def test_quantile_threshold_bit_exact_vs_torch(hw):
    """Same cases as SURVEY.md §4: n in {1024, 8192, 49152, 360448} x the quality sweep, incl. ties."""
    from progressivecodec_b200.engine import Act

    E = _engine()
    g = torch.Generator().manual_seed(hw[0])
    B = 3
    scale = torch.randn(B, hw[0], hw[1], 32, generator=g) * 0.4 + 0.3
    scale[1] = torch.round(scale[1] * 8) / 8          # heavy ties
    scale[2, :, :, :4] = 0.25                          # a constant block
    a = Act(scale.cuda().contiguous())
    for pr in [0.05, 0.1, 0.25, 0.5, 0.6, 0.75, 1, 1.25, 2, 3, 5, 7.5, 9.99]:
        q = 1.0 - pr * 0.1
        thr = E.quantile_threshold(a, q).cpu()
        for b in range(B):
            flat = scale[b].permute(2, 0, 1).reshape(-1)
            ref = torch.quantile(flat, q)
            assert thr[b].item() == ref.item(), (hw, pr, b, thr[b].item(), ref.item())
            assert EP.quantile_threshold_np(flat.numpy(), q) == ref.item()


def test_quantile_on_strided_channel_view():
    from progressivecodec_b200.engine import Act

    E = _engine()
    g = torch.Generator().manual_seed(1)
    big = torch.randn(2, 8, 12, 96, generator=g)
    a = Act(big.cuda().contiguous()).slice(32, 32)
    thr = E.quantile_threshold(a, 0.5).cpu()
    for b in range(2):
        assert thr[b].item() == torch.quantile(big[b, :, :, 32:64].reshape(-1), 0.5).item()


@pytest.mark.parametrize("mode", ["ones", "zeros", "threshold"])
@pytest.mark.parametrize("delta", [False, True])
def test_slice_quantize_matches_oracle(mode, delta):
    from progressivecodec_b200 import _lib as L
    from progressivecodec_b200.engine import Act, new_act

    E = _engine()
    t = EP.GaussianTables.build()
    g = torch.Generator().manual_seed(11)
    B, H, W = 2, 16, 24
    y_all = torch.randn(B, 64, H, W, generator=g) * 2
    y_all[0, 32, 0, :8] = torch.tensor([0.5, 1.5, 2.5, -0.5, -1.5, 3.5, -2.5, 0.0])  # rounding ties (half-even)
    mu = torch.randn(B, 32, H, W, generator=g) * 0.3
    mu[0, 0, 0, :8] = 0.0
    scale = torch.rand(B, 32, H, W, generator=g) * 4 - 0.5
    scale[0, 1, 0, :4] = torch.tensor([0.11, 0.110001, 256.0, 300.0])
    nhwc = lambda x: x.permute(0, 2, 3, 1).contiguous().cuda()
    ya, mua, sa = Act(nhwc(y_all)), Act(nhwc(mu)), Act(nhwc(scale))
    q = 0.7
    if mode == "threshold":
        thr = E.quantile_threshold(sa, q)
        m = torch.stack([(scale[b] >= torch.quantile(scale[b].reshape(-1), q)).float() for b in range(B)])
        mm = L.MASK_THRESHOLD
    else:
        thr = None
        m = torch.ones_like(scale) if mode == "ones" else torch.zeros_like(scale)
        mm = L.MASK_ONES if mode == "ones" else L.MASK_ZEROS
    sym = torch.empty((B, 32 * H * W), dtype=torch.int32, device="cuda")
    idx = torch.empty_like(sym)
    msk = torch.empty((B, 32, H, W), device="cuda")
    lik = torch.empty((B, 32, H, W), device="cuda")
    y_pre = new_act(B, H, W, 32, E.device)
    E.slice_quantize(ya.slice(32, 32), ya.slice(0, 32) if delta else None, mua, sa, mm, thr,
                     t.scale_table.cuda(), t.scale_bound, sym, idx, msk, lik, y_pre)
    ys = y_all[:, 32:] - (y_all[:, :32] if delta else 0)
    ref_sym = torch.round((ys - mu) * m).int()
    ref_idx = t.build_indexes(scale * m)
    assert torch.equal(msk.cpu(), m)
    assert torch.equal(sym.cpu().reshape(B, 32, H, W), ref_sym)
    assert torch.equal(idx.cpu().reshape(B, 32, H, W), ref_idx)
    assert torch.equal(y_pre.t.cpu().permute(0, 3, 1, 2), ref_sym.float() + mu)
    ref_lik = t.likelihood((ys - mu) * m, scale * m, None)
    assert torch.allclose(lik.cpu(), ref_lik, rtol=2e-4, atol=1e-7)
    # decoder side: dequantize from NCHW symbols
    y2 = new_act(B, H, W, 32, E.device)
    E.slice_dequantize(sym, mua, y2)
    assert torch.equal(y2.t, y_pre.t)


def test_bottleneck_ops_match_oracle():
    from conftest import build_pair
    from progressivecodec_b200.engine import Act, new_act

    net, orc = build_pair("authors")
    E = _engine()
    g = torch.Generator().manual_seed(2)
    B, Cz, H, W = 2, 192, 3, 5
    z = torch.randn(B, Cz, H, W, generator=g) * 3
    z[0, 0, 0, :4] = orc.eb.medians[0] + torch.tensor([0.5, -0.5, 1.5, 20.0])
    za = Act(z.permute(0, 2, 3, 1).contiguous().cuda())
    med = orc.eb.medians.cuda().contiguous()
    sym = torch.empty((B, Cz * H * W), dtype=torch.int32, device="cuda")
    idx = torch.empty_like(sym)
    zh = new_act(B, H, W, Cz, E.device)
    E.bottleneck_quantize(za, med, sym, idx, zh)
    ref_sym = orc.eb.symbols(z)
    assert torch.equal(sym.cpu().reshape(B, Cz, H, W), ref_sym)
    assert torch.equal(idx.cpu().reshape(B, Cz, H, W), orc.eb._indexes((B, Cz, H, W)))
    assert torch.equal(zh.t.cpu().permute(0, 3, 1, 2), orc.eb.dequantize(ref_sym))
    assert torch.equal(E.bottleneck_indexes(B, H * W, Cz), idx)
    zh2 = new_act(B, H, W, Cz, E.device)
    E.bottleneck_dequantize(sym, med, zh2)
    assert torch.equal(zh2.t, zh.t)
    lik = E.bottleneck_likelihood(zh, net.entropy_bottleneck.likelihood_params(E.device))
    ref = orc.eb.likelihood(orc.eb.dequantize(ref_sym))
    assert torch.allclose(lik.cpu(), ref, rtol=1e-3, atol=1e-8)
    # module-level API (reference call shape)
    eb = net.entropy_bottleneck.cuda()
    strings = eb.compress(z.cuda())
    assert strings == orc.eb.encode(ref_sym, EP.CPortCoder())
    assert torch.equal(eb.decompress(strings, (H, W)).cpu(), orc.eb.dequantize(ref_sym))
    out, lik2 = eb(z.cuda(), training=False)
    assert torch.equal(out.cpu(), orc.eb.dequantize(ref_sym))
    assert torch.allclose(lik2.cpu(), ref, rtol=1e-3, atol=1e-8)


def test_layout_round_trip():
    E = _engine()
    x = torch.randn(2, 37, 5, 7)
    a = E.from_nchw(x.cuda(), c_pad=48)
    assert torch.equal(a.t[..., :37].cpu(), x.permute(0, 2, 3, 1))
    assert (a.t[..., 37:] == 0).all()
    assert torch.equal(E.to_nchw(a.slice(0, 37)).cpu(), x)
