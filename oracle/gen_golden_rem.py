"""Golden vectors for the REM wrapper from the REAL reference (compress.models.CHProgREM.PostRateProcessedNetwork).

    python -m oracle.gen_golden_rem [--check]      -> tests/golden/rem.npz

Base net = authors' flags with synthetic weights (as the other goldens); the 3 x 10 LatentRateReduction nets get
name-keyed synthetic weights too (`apply_synthetic_weights(rem.post_latent)`, names relative to post_latent).
Levels: 0, one per refinement interval (0.1 in (0.01, 0.25], 1 in (0.25, 1.75], 5 above), and the top level 10.
`rem_escalation.npz`: the ``checkpoint_rep`` path (CHProgREM.py:338-372, :773, :989) — the reference's own
``extract_chekpoint_representation_from_images`` with ``escalation=True`` chained over the three check levels, then one
compress + decompress at quality 5 on top of the last representation.
"""
from __future__ import annotations

import argparse
import io
import os
import sys
import warnings

import numpy as np
import torch

from .gen_golden import CASES, GOLD, ROOT, build_reference, pack_strings, synthetic_image, unpack_strings

REM_QUALITIES = [0, 0.1, 1, 5, 10]
REM_KW = dict(check_levels=[0.01, 0.25, 1.75], mu_std=False, dimension="big")
# second variant: mu and sigma both refined, the smaller nets, two check levels
REM_KW2 = dict(check_levels=[0.05, 1.0], mu_std=True, dimension="middle")
REM_QUALITIES2 = [0.5, 5]


def build_reference_rem(kwargs, rem_kw=None):
    rem_kw = REM_KW if rem_kw is None else rem_kw
    base = build_reference(kwargs)
    warnings.filterwarnings("ignore")
    from compress.models.CHProgREM import PostRateProcessedNetwork  # type: ignore

    sys.path.insert(0, ROOT)
    from progressivecodec_b200.synthetic import apply_synthetic_weights

    out = io.StringIO()
    stdout, sys.stdout = sys.stdout, out
    try:
        rem = PostRateProcessedNetwork(base, **rem_kw)
    finally:
        sys.stdout = stdout
    rem.eval()
    apply_synthetic_weights(rem.post_latent, seed=1)
    return rem


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    a = ap.parse_args()
    torch.set_num_threads(os.cpu_count() or 1)
    run("rem", REM_KW, REM_QUALITIES, 55, a.check)
    run("rem_mustd", REM_KW2, REM_QUALITIES2, 56, a.check)
    run_escalation(a.check)


def run_escalation(check):
    kwargs, shape = CASES["authors"]
    rem_kw = dict(REM_KW, escalation=True)
    rem = build_reference_rem(kwargs, rem_kw)
    x = synthetic_image(shape, seed=57)
    rec = {"x": x.numpy()}
    with torch.no_grad():
        out = io.StringIO()
        stdout, sys.stdout = sys.stdout, out
        try:
            reps = [rem.extract_chekpoint_representation_from_images(x, q) for q in rem_kw["check_levels"]]
        finally:
            sys.stdout = stdout
        for k, r in enumerate(reps):
            rec[f"rep{k}"] = r.numpy()
        c = rem.compress(x, quality=5, mask_pol="point-based-std", checkpoint_rep=reps[-1])
        d = rem.decompress(c["strings"], c["shape"], quality=5, mask_pol="point-based-std", checkpoint_rep=reps[-1])
        for k, v in pack_strings(c["strings"]).items():
            rec["q5_" + k] = v
        rec["q5_shape"] = np.array(list(c["shape"]), dtype=np.int64)
        rec["q5_x_hat"] = d["x_hat"].numpy()
        rec["q5_y_hat"] = c["y_hat"].numpy()
    path = os.path.join(GOLD, "rem_escalation.npz")
    np.savez_compressed(path, **rec)
    print(f"[golden] rem_escalation: {os.path.getsize(path) / 1024:.0f} KiB")
    if check:
        from .codec_port import CodecConfig
        from .rem_port import OracleREM

        orc = OracleREM(rem.base_net.state_dict(), rem.post_latent.state_dict(), CodecConfig(**kwargs), **REM_KW)
        cl = REM_KW["check_levels"]
        r = orc.compress(x, quality=cl[0], mask_pol="point-based-std")["y_hat"]
        errs = [float(np.abs(r.numpy() - rec["rep0"]).max())]
        for k in (1, 2):
            r = orc.compress(x, quality=cl[k], mask_pol="point-based-std", checkpoint_rep=r)["y_hat"]
            errs.append(float(np.abs(r.numpy() - rec[f"rep{k}"]).max()))
        c = orc.compress(x, quality=5, mask_pol="point-based-std", checkpoint_rep=torch.from_numpy(rec["rep2"]))
        ref = unpack_strings(rec, "q5_")
        same = c["strings"][0] == ref[0] and c["strings"][1] == ref[1]
        d = orc.decompress(ref, tuple(rec["q5_shape"]), quality=5, mask_pol="point-based-std",
                           checkpoint_rep=torch.from_numpy(rec["rep2"]))
        print(f"   escalation: rep max|d|={errs}, strings identical={same}, "
              f"x_hat max|d|={float(np.abs(d['x_hat'].numpy() - rec['q5_x_hat']).max()):.3g}")


def run(name, rem_kw, qualities, seed, check):
    kwargs, shape = CASES["authors"]
    rem = build_reference_rem(kwargs, rem_kw)
    x = synthetic_image(shape, seed=seed)
    rec = {"x": x.numpy()}
    with torch.no_grad():
        for q in qualities:
            c = rem.compress(x, quality=q, mask_pol="point-based-std")
            d = rem.decompress(c["strings"], c["shape"], quality=q, mask_pol="point-based-std")
            tag = f"q{q}_"
            for k, v in pack_strings(c["strings"]).items():
                rec[tag + k] = v
            rec[tag + "shape"] = np.array(list(c["shape"]), dtype=np.int64)
            rec[tag + "x_hat"] = d["x_hat"].numpy()
            rec[tag + "y_hat"] = c["y_hat"].numpy()
            if c["masks"]:
                rec[tag + "mask_sum"] = np.array([float(m.sum()) for m in c["masks"]])
    with open(os.path.join(GOLD, f"{name}_post_latent_keys.txt"), "w") as f:
        for k, v in rem.post_latent.state_dict().items():
            f.write(f"{k} {tuple(v.shape)} {v.dtype}\n")
    path = os.path.join(GOLD, f"{name}.npz")
    np.savez_compressed(path, **rec)
    print(f"[golden] {name}: {os.path.getsize(path) / 1024:.0f} KiB")
    if check:
        from .codec_port import CodecConfig
        from .rem_port import OracleREM

        orc = OracleREM(rem.base_net.state_dict(), rem.post_latent.state_dict(), CodecConfig(**kwargs), **rem_kw)
        for q in qualities:
            c = orc.compress(x, quality=q, mask_pol="point-based-std")
            ref = unpack_strings(rec, f"q{q}_")
            same = c["strings"][0] == ref[0] and c["strings"][1] == ref[1]
            d = orc.decompress(ref, tuple(rec[f"q{q}_shape"]), quality=q, mask_pol="point-based-std")
            err = float(np.abs(d["x_hat"].numpy() - rec[f"q{q}_x_hat"]).max())
            print(f"   rem q={q}: strings identical={same}, x_hat max|d|={err:.3g}, "
                  f"y_hat max|d|={float(np.abs(c['y_hat'].numpy() - rec[f'q{q}_y_hat']).max()):.3g}")


if __name__ == "__main__":
    main()
