"""fp16-split tcgen05 convolution (csrc/conv_tc16.cu, impl 3) vs plain PyTorch fp32 references: same contract and the
same tolerance (fp32 accumulation-order noise) as the SIMT and 3xTF32 kernels."""
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

RTOL = 2e-5


def _engine(impl=3):
    from progressivecodec_b200.engine import Engine

    return Engine(torch.device("cuda", 0), impl)


def _nhwc(x):
    from progressivecodec_b200.engine import Act

    return Act(x.permute(0, 2, 3, 1).contiguous().cuda())


def _nchw(a):
    return a.t[..., a.c0:a.c0 + a.C].permute(0, 3, 1, 2).cpu()


def _close(got, ref, rtol=RTOL):
    scale = ref.pow(2).mean().sqrt().item() + 1e-12
    err = (got - ref).abs().max().item()
    assert err <= rtol * scale * 8, (err, scale)


def _planes_value(a):
    """fp32 reconstruction hi + lo * 2^-11 of an activation's planes (window a)."""
    r = a._root
    n = a.B * a.H * a.W * a.ps
    import ctypes

    buf = torch.empty(2 * n, dtype=torch.float16, device="cuda")
    ctypes.pythonapi  # noqa: B018 (keep ctypes imported)
    torch.cuda.synchronize()
    # planes live at raw device pointers: copy them out with cudaMemcpy through torch's from-pointer route
    from progressivecodec_b200.engine import Act  # noqa: F401

    hi = _from_ptr(r.hi, n)
    lo = _from_ptr(r.lo, n)
    v = hi.float() + lo.float() / 2048.0
    return v.view(a.B, a.H, a.W, a.ps)[..., a.c0:a.c0 + a.C].permute(0, 3, 1, 2).cpu()


def _from_ptr(ptr, n):
    import ctypes

    out = torch.empty(n, dtype=torch.float16, device="cuda")
    cudart = ctypes.CDLL("libcudart.so")
    rc = cudart.cudaMemcpy(ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(ptr), ctypes.c_size_t(2 * n), ctypes.c_int(3))
    assert rc == 0
    return out


def test_split_planes_hold_22_bits():
    E = _engine()
    torch.manual_seed(0)
    x = torch.randn(2, 64, 5, 7) * torch.logspace(-6, 3, 64).view(1, 64, 1, 1)
    a = _nhwc(x)
    assert E.planes(a) is not None
    rec = _planes_value(a)
    big = x.abs() >= 1e-4  # hi is a NORMAL fp16 number: 11 + 11 significant bits
    rel = ((rec - x).abs() / x.abs().clamp_min(1e-30))[big].max().item()
    assert rel <= 2.0 ** -21, rel
    assert (rec - x).abs()[~big].max().item() <= 4e-11
    # values below the fp16 normal range keep an ABSOLUTE error far below anything a convolution output can see
    tiny = torch.randn(1, 8, 4, 4) * 1e-7
    b = _nhwc(tiny)
    E.planes(b)
    assert (_planes_value(b) - tiny).abs().max().item() <= 4e-11


@pytest.mark.parametrize("cin,cout,k,stride,hw", [(192, 192, 5, 2, (32, 48)), (192, 320, 5, 2, (16, 24)),
                                                   (96, 96, 3, 1, (16, 24)), (192, 96, 1, 1, (9, 7)),
                                                   (288, 256, 3, 2, (8, 12)), (64, 32, 3, 1, (32, 48)),
                                                   (224, 176, 3, 1, (5, 11)), (176, 128, 3, 1, (7, 9)),
                                                   (320, 640, 5, 2, (8, 8)), (192, 576, 1, 1, (16, 16)),
                                                   (512, 224, 3, 1, (32, 48)), (80, 192, 1, 1, (20, 24))])
def test_tc16_conv2d_matches_torch(cin, cout, k, stride, hw):
    from progressivecodec_b200 import _lib as L
    from progressivecodec_b200.engine import pack_conv2d

    E = _engine()
    torch.manual_seed(cin + cout + k)
    m = nn.Conv2d(cin, cout, k, stride, k // 2)
    x = torch.randn(3, cin, *hw)
    ref = m(x).detach()
    pc = pack_conv2d(m, E.device, "t").attach_tc(3)
    assert pc.tc is not None
    out = E.conv_new(pc, [_nhwc(x)])
    _close(_nchw(out), ref)
    # the planes the epilogue wrote for the next convolution hold the same values to 22 bits
    rec = _planes_value(out)
    assert (rec - _nchw(out)).abs().max().item() <= 2.0 ** -21 * _nchw(out).abs().max().item()
    _close(_nchw(E.conv_new(pc, [_nhwc(x)], L.EPI_GELU)), F.gelu(ref))


def test_tc16_chain_through_planes_only():
    """conv -> GELU -> conv where the intermediate exists ONLY as split planes (fmt = 2), as inside the slice stacks."""
    from progressivecodec_b200 import _lib as L
    from progressivecodec_b200.engine import pack_conv2d

    E = _engine()
    torch.manual_seed(1)
    m1, m2 = nn.Conv2d(96, 224, 3, 1, 1), nn.Conv2d(224, 64, 3, 1, 1)
    x = torch.randn(2, 96, 16, 24)
    ref = m2(F.gelu(m1(x))).detach()
    p1, p2 = pack_conv2d(m1, E.device, "a").attach_tc(3), pack_conv2d(m2, E.device, "b").attach_tc(3)
    h = E.act(2, 16, 24, 224)
    E.conv(p1, [_nhwc(x)], h, L.EPI_GELU, fmt=2)
    _close(_nchw(E.conv_new(p2, [h])), ref)


def test_tc16_first_conv_gdn_deconv_shuffle_concat():
    from progressivecodec_b200 import _lib as L
    from progressivecodec_b200.engine import (new_act, pack_conv2d, pack_deconv_phases, pack_first_conv_im2col, pack_gdn)
    from progressivecodec_b200.layers import GDN
    from progressivecodec_b200.synthetic import synthetic_tensor

    E = _engine()
    torch.manual_seed(0)
    m = nn.Conv2d(3, 192, 5, 2, 2)
    x = torch.rand(2, 3, 64, 128)
    pc = pack_first_conv_im2col(m, E.device, 80, "c0").attach_tc()
    _close(_nchw(E.conv_new(pc, [E.im2col_first(x.cuda(), 5, 2, 2, 80)])), m(x).detach())
    for inverse in (False, True):
        g = GDN(192, inverse=inverse)
        with torch.no_grad():
            g.beta.copy_(synthetic_tensor("g.beta", g.beta, 0))
            g.gamma.copy_(synthetic_tensor("g.gamma", g.gamma, 0))
        xx = torch.randn(2, 192, 12, 20)
        beta, gamma = g.effective()
        norm = F.conv2d(xx * xx, gamma.reshape(192, 192, 1, 1), beta)
        ref = xx * (torch.sqrt(norm) if inverse else torch.rsqrt(norm))
        _close(_nchw(E.gdn_new(pack_gdn(g, E.device, "g").attach_tc(), _nhwc(xx), inverse)), ref)
    md = nn.ConvTranspose2d(320, 192, 5, 2, 2, 1)
    xd = torch.randn(2, 320, 4, 6)
    ph = [p.attach_tc() for p in pack_deconv_phases(md, E.device, "d")]
    _close(_nchw(E.deconv_new(ph, _nhwc(xd))), md(xd).detach())
    ms = nn.Conv2d(192, 224 * 4, 3, 1, 1)
    xs = torch.randn(2, 192, 4, 6)
    sh = E.conv_shuffle_new(pack_conv2d(ms, E.device, "s").attach_tc(), _nhwc(xs), L.EPI_GELU)
    ref_sh = F.gelu(F.pixel_shuffle(ms(xs), 2)).detach()
    _close(_nchw(sh), ref_sh)
    assert (_planes_value(sh) - _nchw(sh)).abs().max().item() <= 2.0 ** -21 * ref_sh.abs().max().item()
    lm, yb, yp = torch.randn(2, 640, 8, 12), torch.randn(2, 320, 8, 12), torch.randn(2, 32, 8, 12)
    mc = nn.Conv2d(320 + 96 + 32, 224, 3, 1, 1)
    ref = mc(torch.cat([lm[:, 320:], yb[:, 64:160], yp], 1)).detach()
    segs = [_nhwc(lm).slice(320, 320), _nhwc(yb).slice(64, 96), _nhwc(yp)]
    _close(_nchw(E.conv_new(pack_conv2d(mc, E.device, "cat").attach_tc(), segs)), ref)
    m2 = nn.Conv2d(64, 32, 3, 1, 1)
    hh, r1, r2 = torch.randn(2, 64, 8, 12), torch.randn(2, 32, 8, 12), torch.randn(2, 32, 8, 12)
    buf = new_act(2, 8, 12, 320, E.device)
    buf.t.fill_(7.0)
    out = buf.slice(96, 32)
    E.conv(pack_conv2d(m2, E.device, "e").attach_tc(), [_nhwc(hh)], out, L.EPI_LRP, r1=_nhwc(r1), r2=_nhwc(r2))
    _close(_nchw(out), r1 + 0.5 * torch.tanh(m2(hh).detach()) + r2)
    assert (buf.t[..., :96] == 7.0).all() and (buf.t[..., 128:] == 7.0).all()


def test_tc16_conv_is_batch_invariant_and_deterministic():
    from progressivecodec_b200.engine import pack_conv2d

    E = _engine()
    torch.manual_seed(5)
    m = nn.Conv2d(352, 224, 3, 1, 1)
    x = torch.randn(5, 352, 32, 48)
    pc = pack_conv2d(m, E.device, "b").attach_tc()
    full = _nchw(E.conv_new(pc, [_nhwc(x)]))
    assert torch.equal(full, _nchw(E.conv_new(pc, [_nhwc(x)])))
    for b in (0, 3, 4):
        assert torch.equal(_nchw(E.conv_new(pc, [_nhwc(x[b:b + 1])]))[0], full[b])


@pytest.mark.parametrize("hw,batch", [((32, 48), 2), ((7, 9), 3), ((64, 96), 1)])
def test_tc16_merged_image_layer(hw, batch):
    from progressivecodec_b200 import _lib as L
    from progressivecodec_b200.engine import pack_deconv_merged_image

    E = _engine()
    torch.manual_seed(hw[0])
    m = nn.ConvTranspose2d(192, 3, 5, 2, 2, 1)
    x = torch.randn(batch, 192, *hw)
    ref = m(x).detach()
    pc = pack_deconv_merged_image(m, E.device, "img").attach_tc(3)
    got = E.deconv_image(pc, _nhwc(x), L.EPI_LINEAR)
    _close(got.cpu(), ref)
    _close(E.deconv_image(pc, _nhwc(x), L.EPI_CLAMP01).cpu(), ref.clamp(0, 1))


def test_tc16_large_m_many_tiles():
    """3x3 32->32 over 5 x 1300 x 1301 pixels (66 k tiles per image row block; ragged right / bottom tiles)."""
    from progressivecodec_b200 import _lib as L
    from progressivecodec_b200.engine import Act, pack_conv2d

    E = _engine()
    torch.manual_seed(3)
    m = nn.Conv2d(32, 32, 3, 1, 1)
    x = torch.randn(5, 32, 1300, 1301, device="cuda")
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        ref = F.gelu(m.cuda()(x)).detach()
    finally:
        torch.backends.cudnn.allow_tf32 = old
    pc = pack_conv2d(m, E.device, "big").attach_tc(3)
    out = E.conv_new(pc, [Act(x.permute(0, 2, 3, 1).contiguous())], L.EPI_GELU)
    got = out.t.permute(0, 3, 1, 2)
    scale = ref.pow(2).mean().sqrt().item()
    assert (got - ref).abs().max().item() <= 8 * RTOL * scale


def test_tc16_residual_from_planes_and_squared_planes_for_gdn():
    """(a) a ResidualUnit-like chain whose tensors exist only as split planes: the 1x1 epilogue reads the residual operand
    from the planes; (b) conv -> GDN where the conv's epilogue writes the planes of x*x, so the GDN launch needs no
    separate squaring pass."""
    from progressivecodec_b200 import _lib as L
    from progressivecodec_b200.engine import pack_conv2d, pack_gdn
    from progressivecodec_b200.layers import GDN
    from progressivecodec_b200.synthetic import synthetic_tensor

    E = _engine()
    torch.manual_seed(2)
    m0, m1 = nn.Conv2d(64, 192, 1), nn.Conv2d(192, 192, 1)
    x = torch.randn(2, 64, 16, 24)
    h0 = m0(x)
    ref = F.gelu(m1(h0) + h0).detach()
    p0, p1 = pack_conv2d(m0, E.device, "a").attach_tc(3), pack_conv2d(m1, E.device, "b").attach_tc(3)
    a0 = E.conv_new(p0, [_nhwc(x)], fmt=2)          # planes only
    assert a0.base == 0
    out = E.conv_new(p1, [a0], L.EPI_ADD_GELU, r1=a0)
    _close(_nchw(out), ref)
    g = GDN(192)
    with torch.no_grad():
        g.beta.copy_(synthetic_tensor("g.beta", g.beta, 0))
        g.gamma.copy_(synthetic_tensor("g.gamma", g.gamma, 0))
    beta, gamma = g.effective()
    hx = m0(x).detach()
    refg = hx * torch.rsqrt(F.conv2d(hx * hx, gamma.reshape(192, 192, 1, 1), beta))
    c = E.conv_new(p0, [_nhwc(x)], square_planes=True)
    assert c._root.sq
    _close(_nchw(E.gdn_new(pack_gdn(g, E.device, "g").attach_tc(), c, False)), refg)
