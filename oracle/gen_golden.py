"""Generate tests/golden/*.npz by running the REAL reference (container only; needs /root/reference).

    python -m oracle.gen_golden            # writes tests/golden/
    python -m oracle.gen_golden --check    # additionally runs the oracle port and prints max deviations

The reference model (``compress.models.ChannelProgresssiveWACNN``, imported unmodified from
/root/reference/src with the native coder compiled into oracle/_ref) is filled with
``progressivecodec_b200.synthetic.apply_synthetic_weights`` (name-keyed, so every implementation
gets the same weights) and run on seeded inputs.  Saved per case: bit streams, reconstructions,
likelihood tensors and intermediate latents.  Fixtures are small (64x128 inputs) so they can live
in git; full-size parity is checked on the GPU box against the oracle port, which these fixtures pin.
"""
from __future__ import annotations

import argparse
import io
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")

CASES = {
    # name: (ctor kwargs, input shape)
    "authors": (dict(multiple_decoder=True, multiple_encoder=False, multiple_hyperprior=True, delta_encode=True,
                     support_progressive_slices=5, mask_policy="point-based-std"), (1, 3, 64, 128)),
    "multienc": (dict(multiple_decoder=True, multiple_encoder=True, multiple_hyperprior=True, delta_encode=True,
                      support_progressive_slices=5, mask_policy="point-based-std"), (2, 3, 64, 64)),
    "allscalable": (dict(multiple_decoder=True, multiple_encoder=False, multiple_hyperprior=True, delta_encode=True,
                         support_progressive_slices=5, mask_policy="point-based-std", all_scalable=True),
                    (1, 3, 64, 64)),
    "plain": (dict(multiple_decoder=True, multiple_encoder=True, multiple_hyperprior=False, delta_encode=False,
                   support_progressive_slices=0, mask_policy="two-levels"), (1, 3, 64, 64)),
}
QUALITIES = [0, 0.05, 0.5, 1.25, 5, 10]
FWD_QUALITIES = [0, 0.5, 5, 10]


def synthetic_image(shape, seed: int) -> torch.Tensor:
    """Seeded low-pass-filtered noise in [0,1] (SURVEY.md §8d 'Synthetic inputs')."""
    g = torch.Generator().manual_seed(1000 + seed)
    x = torch.rand(*shape, generator=g)
    x = torch.nn.functional.avg_pool2d(torch.nn.functional.pad(x, (2, 2, 2, 2), mode="reflect"), 5, 1)
    return x.contiguous()


def pack_strings(strings) -> dict:
    """strings = [list_over_slices[list_over_batch[bytes]], list_over_batch[bytes]] -> flat uint8 + lengths."""
    ys, zs = strings
    flat = [s for sl in ys for s in sl] + list(zs)
    lens = np.array([len(s) for s in flat], dtype=np.int64)
    blob = np.frombuffer(b"".join(flat), dtype=np.uint8) if lens.sum() else np.zeros(0, np.uint8)
    return {"blob": blob, "lens": lens, "n_slices": np.int64(len(ys)), "batch": np.int64(len(zs))}


def unpack_strings(d, prefix: str):
    blob, lens = d[prefix + "blob"], d[prefix + "lens"]
    ns, b = int(d[prefix + "n_slices"]), int(d[prefix + "batch"])
    offs = np.concatenate([[0], np.cumsum(lens)])
    flat = [blob[offs[i]:offs[i + 1]].tobytes() for i in range(len(lens))]
    ys = [flat[s * b:(s + 1) * b] for s in range(ns)]
    zs = flat[ns * b:]
    return [ys, zs]


def build_reference(kwargs):
    from . import build_ref

    build_ref.import_reference()
    warnings.filterwarnings("ignore")
    from compress.models import ChannelProgresssiveWACNN  # type: ignore

    sys.path.insert(0, ROOT)
    from progressivecodec_b200.synthetic import apply_synthetic_weights

    out = io.StringIO()
    stdout, sys.stdout = sys.stdout, out  # the reference constructor print()s
    try:
        torch.manual_seed(0)
        net = ChannelProgresssiveWACNN(lmbda_list=[0.005, 0.05], **kwargs)
    finally:
        sys.stdout = stdout
    net.eval()
    apply_synthetic_weights(net, seed=0)
    net.update(force=True)
    return net


def run_case(name: str, check: bool):
    kwargs, shape = CASES[name]
    net = build_reference(kwargs)
    x = synthetic_image(shape, seed=len(name))
    rec = {"x": x.numpy()}
    pol = kwargs["mask_policy"]
    with torch.no_grad():
        for q in QUALITIES:
            if pol == "two-levels" and q not in (0, 10):
                continue
            c = net.compress(x, quality=q, mask_pol=pol)
            d = net.decompress(c["strings"], c["shape"], quality=q, mask_pol=pol)
            tag = f"q{q}_"
            for k, v in pack_strings(c["strings"]).items():
                rec[tag + k] = v
            rec[tag + "shape"] = np.array(list(c["shape"]), dtype=np.int64)
            rec[tag + "x_hat"] = d["x_hat"].numpy()
            if c["masks"]:
                rec[tag + "mask_sum"] = np.array([float(m.sum()) for m in c["masks"]])
        for q in FWD_QUALITIES:
            if pol == "two-levels" and q not in (0, 10):
                continue
            o = net.forward_single_quality(x, q, mask_pol=pol, training=False)
            tag = f"fsq{q}_"
            rec[tag + "x_hat"] = o["x_hat"].numpy()
            rec[tag + "lik_y"] = o["likelihoods"]["y"].numpy()
            rec[tag + "lik_z"] = o["likelihoods"]["z"].numpy()
            rec[tag + "y_hat"] = o["y_hat"].numpy()
        ql = [0, 0.5, 5, 10] if pol != "two-levels" else [0, 10]
        o = net.forward(x, quality=ql, mask_pol=pol, training=False)
        rec["fwd_qualities"] = np.array(ql, dtype=np.float64)
        rec["fwd_x_hat"] = o["x_hat"].numpy()
        rec["fwd_lik_y"] = o["likelihoods"]["y"].numpy()
        rec["fwd_lik_y_prog"] = o["likelihoods"]["y_prog"].numpy()
        rec["fwd_lik_z"] = o["likelihoods"]["z"].numpy()
    # entropy tables (update() output) — pins GaussianTables.build / BottleneckTables.rebuild
    gc, eb = net.gaussian_conditional, net.entropy_bottleneck
    rec["gc_cdf"] = gc._quantized_cdf.numpy().astype(np.int32)
    rec["gc_cdf_length"] = gc._cdf_length.numpy().astype(np.int32)
    rec["gc_offset"] = gc._offset.numpy().astype(np.int32)
    rec["gc_scale_table"] = gc.scale_table.numpy()
    rec["eb_cdf"] = eb._quantized_cdf.numpy().astype(np.int32)
    rec["eb_cdf_length"] = eb._cdf_length.numpy().astype(np.int32)
    rec["eb_offset"] = eb._offset.numpy().astype(np.int32)
    os.makedirs(GOLD, exist_ok=True)
    with open(os.path.join(GOLD, f"{name}_state_dict_keys.txt"), "w") as f:
        for k, v in net.state_dict().items():
            f.write(f"{k} {tuple(v.shape)} {v.dtype}\n")
    path = os.path.join(GOLD, f"{name}.npz")
    np.savez_compressed(path, **rec)
    print(f"[golden] {name}: {os.path.getsize(path) / 1024:.0f} KiB, {len(rec)} arrays")
    if check:
        from .codec_port import CodecConfig, OracleCodec

        orc = OracleCodec(net.state_dict(), CodecConfig(**kwargs))
        for q in QUALITIES:
            if pol == "two-levels" and q not in (0, 10):
                continue
            c = orc.compress(x, quality=q, mask_pol=pol)
            ref = unpack_strings(rec, f"q{q}_")
            same = c["strings"][0] == ref[0] and c["strings"][1] == ref[1]
            d = orc.decompress(ref, tuple(rec[f"q{q}_shape"]), quality=q, mask_pol=pol)
            err = float(np.abs(d["x_hat"].numpy() - rec[f"q{q}_x_hat"]).max())
            print(f"   q={q}: strings identical={same}, x_hat max|d|={err:.3g}")
        for q in FWD_QUALITIES:
            if pol == "two-levels" and q not in (0, 10):
                continue
            o = orc.forward_single_quality(x, q, mask_pol=pol)
            e1 = float(np.abs(o["x_hat"].numpy() - rec[f"fsq{q}_x_hat"]).max())
            e2 = float(np.abs(o["likelihoods"]["y"].numpy() - rec[f"fsq{q}_lik_y"]).max())
            e3 = float(np.abs(o["likelihoods"]["z"].numpy() - rec[f"fsq{q}_lik_z"]).max())
            print(f"   fsq q={q}: x_hat {e1:.3g} lik_y {e2:.3g} lik_z {e3:.3g}")
        o = orc.forward(x, quality=list(rec["fwd_qualities"]), mask_pol=pol)
        print("   forward: x_hat %.3g lik_y %.3g lik_y_prog %.3g lik_z %.3g" % (
            float(np.abs(o["x_hat"].numpy() - rec["fwd_x_hat"]).max()),
            float(np.abs(o["likelihoods"]["y"].numpy() - rec["fwd_lik_y"]).max()),
            float(np.abs(o["likelihoods"]["y_prog"].numpy() - rec["fwd_lik_y_prog"]).max()),
            float(np.abs(o["likelihoods"]["z"].numpy() - rec["fwd_lik_z"]).max())))


def run_custmap(check: bool):
    """cust_map path (masking.py:171-194 via CHProg_cnn.py:721,823,850,964): the importance map replaces sigma in the
    mask; authors' flags, two levels.  -> tests/golden/authors_custmap.npz"""
    kwargs, shape = CASES["authors"]
    net = build_reference(kwargs)
    x = synthetic_image(shape, seed=77)
    g = torch.Generator().manual_seed(4242)
    cmap = torch.rand(shape[0], 320, shape[2] // 16, shape[3] // 16, generator=g)
    rec = {"x": x.numpy(), "cust_map": cmap.numpy()}
    qs = [0.5, 5]
    with torch.no_grad():
        for q in qs:
            c = net.compress(x, quality=q, mask_pol="point-based-std", cust_map=cmap)
            d = net.decompress(c["strings"], c["shape"], quality=q, mask_pol="point-based-std", cust_map=cmap)
            tag = f"q{q}_"
            for k, v in pack_strings(c["strings"]).items():
                rec[tag + k] = v
            rec[tag + "shape"] = np.array(list(c["shape"]), dtype=np.int64)
            rec[tag + "x_hat"] = d["x_hat"].numpy()
            rec[tag + "mask_sum"] = np.array([float(m.sum()) for m in c["masks"]])
    path = os.path.join(GOLD, "authors_custmap.npz")
    np.savez_compressed(path, **rec)
    print(f"[golden] authors_custmap: {os.path.getsize(path) / 1024:.0f} KiB")
    if check:
        from .codec_port import CodecConfig, OracleCodec

        orc = OracleCodec(net.state_dict(), CodecConfig(**kwargs))
        for q in qs:
            c = orc.compress(x, quality=q, mask_pol="point-based-std", cust_map=cmap)
            ref = unpack_strings(rec, f"q{q}_")
            same = c["strings"][0] == ref[0] and c["strings"][1] == ref[1]
            d = orc.decompress(ref, tuple(rec[f"q{q}_shape"]), quality=q, mask_pol="point-based-std", cust_map=cmap)
            err = float(np.abs(d["x_hat"].numpy() - rec[f"q{q}_x_hat"]).max())
            print(f"   custmap q={q}: strings identical={same}, x_hat max|d|={err:.3g}")


def run_table800(check: bool):
    """The 800-level scale table the reference defines but does not use (CHProg_cnn.py:16-26: 0.04 .. 256, 800 levels)
    passed to ``update(scale_table=...)``; authors' flags, two levels.  -> tests/golden/authors_table800.npz"""
    kwargs, shape = CASES["authors"]
    net = build_reference(kwargs)
    from compress.models.CHProg_cnn import get_scale_table as table800  # type: ignore

    net.update(scale_table=table800(), force=True)
    x = synthetic_image(shape, seed=78)
    rec = {"x": x.numpy(), "scale_table": table800().numpy()}
    qs = [0, 5]
    with torch.no_grad():
        for q in qs:
            c = net.compress(x, quality=q, mask_pol="point-based-std")
            d = net.decompress(c["strings"], c["shape"], quality=q, mask_pol="point-based-std")
            tag = f"q{q}_"
            for k, v in pack_strings(c["strings"]).items():
                rec[tag + k] = v
            rec[tag + "shape"] = np.array(list(c["shape"]), dtype=np.int64)
            rec[tag + "x_hat"] = d["x_hat"].numpy()
    path = os.path.join(GOLD, "authors_table800.npz")
    np.savez_compressed(path, **rec)
    print(f"[golden] authors_table800: {os.path.getsize(path) / 1024:.0f} KiB")
    if check:
        from .codec_port import CodecConfig, OracleCodec

        orc = OracleCodec(net.state_dict(), CodecConfig(**kwargs))
        for q in qs:
            c = orc.compress(x, quality=q, mask_pol="point-based-std")
            ref = unpack_strings(rec, f"q{q}_")
            same = c["strings"][0] == ref[0] and c["strings"][1] == ref[1]
            d = orc.decompress(ref, tuple(rec[f"q{q}_shape"]), quality=q, mask_pol="point-based-std")
            err = float(np.abs(d["x_hat"].numpy() - rec[f"q{q}_x_hat"]).max())
            print(f"   table800 q={q}: strings identical={same}, x_hat max|d|={err:.3g}")


HEADLINE_SHAPE = (1, 3, 512, 768)   # BASELINE.json configs[1]: one 768x512 image
HEADLINE_QUALITIES = [0, 5]
HEADLINE_SEED = 0


def headline_digest(strings, x_hat: torch.Tensor, x: torch.Tensor) -> dict:
    """What the headline-shape fixture keeps of one (compress, decompress) result: length and crc32 of every stream, the
    crc32 of the reconstruction's bytes, its PSNR and a 16x16 average-pooled copy (for hosts whose CPU kernels round
    differently from the generating one's)."""
    import zlib

    ys, zs = strings
    flat = [s for sl in ys for s in sl] + list(zs)
    mse = float(((x_hat.double() - x.double()) ** 2).mean())
    return {"lens": np.array([len(s) for s in flat], dtype=np.int64),
            "crcs": np.array([zlib.crc32(s) for s in flat], dtype=np.uint32),
            "x_hat_crc": np.uint32(zlib.crc32(x_hat.contiguous().numpy().tobytes())),
            "psnr": np.float64(10.0 * np.log10(1.0 / mse)),
            "x_hat_pooled": torch.nn.functional.avg_pool2d(x_hat, 16).numpy()}


def run_headline(check: bool):
    """The shape the bench and BASELINE.json's metric are quoted on (768x512, authors' flags): the REAL reference's
    compress() / decompress() at q = 0 and 5, kept as digests.  -> tests/golden/headline_768x512.npz"""
    kwargs, _ = CASES["authors"]
    net = build_reference(kwargs)
    x = synthetic_image(HEADLINE_SHAPE, seed=HEADLINE_SEED)
    rec = {}
    keep = {}
    with torch.no_grad():
        for q in HEADLINE_QUALITIES:
            c = net.compress(x, quality=q, mask_pol="point-based-std")
            d = net.decompress(c["strings"], c["shape"], quality=q, mask_pol="point-based-std")
            keep[q] = (c["strings"], d["x_hat"])
            for k, v in headline_digest(c["strings"], d["x_hat"], x).items():
                rec[f"q{q}_{k}"] = v
    path = os.path.join(GOLD, "headline_768x512.npz")
    np.savez_compressed(path, **rec)
    print(f"[golden] headline_768x512: {os.path.getsize(path) / 1024:.0f} KiB")
    if check:
        from .codec_port import CodecConfig, OracleCodec

        orc = OracleCodec(net.state_dict(), CodecConfig(**kwargs))
        for q in HEADLINE_QUALITIES:
            c = orc.compress(x, quality=q, mask_pol="point-based-std")
            same = c["strings"][0] == keep[q][0][0] and c["strings"][1] == keep[q][0][1]
            d = orc.decompress(keep[q][0], c["shape"], quality=q, mask_pol="point-based-std")
            err = float((d["x_hat"] - keep[q][1]).abs().max())
            print(f"   headline q={q}: strings identical={same} ({int(rec[f'q{q}_lens'].sum())} bytes), "
                  f"x_hat max|d|={err:.3g}, psnr {float(rec[f'q{q}_psnr']):.3f} dB")


SWEEP_LEVELS = [0, 0.05, 0.1, 0.25, 0.5, 0.6, 0.75, 1, 1.25, 2, 3, 5, 10]   # reference train.py:293
CONFIG3_SHAPE, CONFIG3_SEED = (16, 3, 256, 256), 11                       # BASELINE.json configs[2]
CONFIG4_SHAPE, CONFIG4_SEED, CONFIG4_PAD = (1, 3, 1365, 2048), 6, (0, 0, 21, 22)   # configs[3]: CLIC size -> 1408x2048
CONFIG4_QUALITIES = [0, 5]


def config4_image() -> torch.Tensor:
    return torch.nn.functional.pad(synthetic_image(CONFIG4_SHAPE, seed=CONFIG4_SEED), CONFIG4_PAD)


def forward_digest(o: dict, x: torch.Tensor) -> dict:
    """What the config-3 fixture keeps of forward(x, quality=[levels]): per level the reconstruction's crc32 and PSNR,
    and the rate the likelihoods imply (bits per pixel, float64 sums); crc32 of the three likelihood tensors."""
    import zlib

    L = o["x_hat"].shape[0]
    npx = x.shape[0] * x.shape[2] * x.shape[3]
    crc = lambda t: np.uint32(zlib.crc32(t.contiguous().numpy().tobytes()))
    bits = lambda t: float(-torch.log2(t.double()).sum()) / npx
    yp = o["likelihoods"]["y_prog"]
    return {"x_hat_crc": np.array([crc(o["x_hat"][l]) for l in range(L)], dtype=np.uint32),
            "psnr": np.array([10.0 * np.log10(1.0 / float(((o["x_hat"][l].clamp(0, 1).double() - x.double()) ** 2).mean()))
                              for l in range(L)]),
            "bpp_y_prog": np.array([bits(yp[l]) for l in range(yp.shape[0])]),
            "bpp_y": np.float64(bits(o["likelihoods"]["y"])), "bpp_z": np.float64(bits(o["likelihoods"]["z"])),
            "lik_crc": np.array([crc(o["likelihoods"][k]) for k in ("y", "y_prog", "z")], dtype=np.uint32)}


def run_config3(check: bool):
    """BASELINE.json configs[2]: batched forward() on the 16x3x256x256 training-crop shape at all 13 levels, real
    reference, kept as digests.  -> tests/golden/config3_forward_16x256x256.npz"""
    kwargs, _ = CASES["authors"]
    net = build_reference(kwargs)
    x = synthetic_image(CONFIG3_SHAPE, seed=CONFIG3_SEED)
    with torch.no_grad():
        o = net.forward(x, quality=SWEEP_LEVELS, mask_pol="point-based-std", training=False)
    rec = forward_digest(o, x)
    path = os.path.join(GOLD, "config3_forward_16x256x256.npz")
    np.savez_compressed(path, **rec)
    print(f"[golden] config3_forward_16x256x256: {os.path.getsize(path) / 1024:.1f} KiB")
    if check:
        from .codec_port import CodecConfig, OracleCodec

        orc = OracleCodec(net.state_dict(), CodecConfig(**kwargs))
        oo = orc.forward(x, quality=SWEEP_LEVELS, mask_pol="point-based-std")
        d = forward_digest(oo, x)
        print(f"   config3: x_hat identical at {int((d['x_hat_crc'] == rec['x_hat_crc']).sum())} of {len(SWEEP_LEVELS)} levels, "
              f"likelihoods identical={bool((d['lik_crc'] == rec['lik_crc']).all())}, "
              f"x_hat max|d|={float((oo['x_hat'] - o['x_hat']).abs().max()):.3g}")


def run_config4(check: bool):
    """BASELINE.json configs[3] shape: one 2048x1365 image padded to 2048x1408, real reference compress() /
    decompress() at q = 0 and 5, kept as digests.  -> tests/golden/config4_2048x1408.npz"""
    kwargs, _ = CASES["authors"]
    net = build_reference(kwargs)
    x = config4_image()
    rec, keep = {}, {}
    with torch.no_grad():
        for q in CONFIG4_QUALITIES:
            c = net.compress(x, quality=q, mask_pol="point-based-std")
            d = net.decompress(c["strings"], c["shape"], quality=q, mask_pol="point-based-std")
            keep[q] = (c["strings"], d["x_hat"])
            for k, v in headline_digest(c["strings"], d["x_hat"], x).items():
                if k != "x_hat_pooled":
                    rec[f"q{q}_{k}"] = v
    path = os.path.join(GOLD, "config4_2048x1408.npz")
    np.savez_compressed(path, **rec)
    print(f"[golden] config4_2048x1408: {os.path.getsize(path) / 1024:.1f} KiB")
    if check:
        from .codec_port import CodecConfig, OracleCodec

        orc = OracleCodec(net.state_dict(), CodecConfig(**kwargs))
        for q in CONFIG4_QUALITIES:
            c = orc.compress(x, quality=q, mask_pol="point-based-std")
            same = c["strings"][0] == keep[q][0][0] and c["strings"][1] == keep[q][0][1]
            d = orc.decompress(keep[q][0], c["shape"], quality=q, mask_pol="point-based-std")
            print(f"   config4 q={q}: strings identical={same} ({int(rec[f'q{q}_lens'].sum())} bytes), "
                  f"x_hat max|d|={float((d['x_hat'] - keep[q][1]).abs().max()):.3g}, psnr {float(rec[f'q{q}_psnr']):.3f} dB")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--cases", nargs="*", default=list(CASES))
    a = ap.parse_args()
    torch.set_num_threads(os.cpu_count() or 1)
    for name in a.cases:
        if name == "custmap":
            run_custmap(a.check)
        elif name == "table800":
            run_table800(a.check)
        elif name == "headline":
            run_headline(a.check)
        elif name == "config3":
            run_config3(a.check)
        elif name == "config4":
            run_config4(a.check)
        else:
            run_case(name, a.check)


if __name__ == "__main__":
    main()
