"""Evaluation harness (training/step.py:277-404 protocol) on the B200 path vs the same protocol on the CPU oracle."""
import os

import pytest
import torch
import torch.nn.functional as F

from conftest import build_pair
from oracle.codec_port import psnr
from oracle.gen_golden import synthetic_image

pytestmark = pytest.mark.gpu


def test_compress_with_ac_matches_oracle_protocol(tmp_path):
    from progressivecodec_b200.evaluation import compress_with_ac, compute_padding

    net, orc = build_pair("authors", "cuda")
    imgs = [synthetic_image((1, 3, 200, 300), seed=31)[0], synthetic_image((1, 3, 170, 190), seed=32)[0]]
    levels = [0, 1, 10]
    bpp, ps, dt = compress_with_ac(net, imgs, torch.device("cuda", 0), pr_list=levels, mask_pol="point-based-std",
                                   writing=str(tmp_path))
    assert len(bpp) == len(ps) == len(dt) == 3 and all(t > 0 for t in dt)
    # the same protocol on the oracle
    ref_bpp, ref_psnr = [0.0] * 3, [0.0] * 3
    for x in imgs:
        x = x.unsqueeze(0)
        pad, unpad = compute_padding(x.shape[2], x.shape[3], min_div=64)
        xp = F.pad(x, pad)
        for j, q in enumerate(levels):
            c = orc.compress(xp, quality=q, mask_pol="point-based-std")
            r = F.pad(orc.decompress(c["strings"], c["shape"], quality=q, mask_pol="point-based-std")["x_hat"], unpad).clamp(0, 1)
            npx = r.shape[0] * r.shape[2] * r.shape[3]
            ref_bpp[j] += (sum(len(s[0]) for s in c["strings"][0]) + sum(len(s) for s in c["strings"][1])) * 8.0 / npx / len(imgs)
            ref_psnr[j] += psnr(r, x) / len(imgs)
    for j in range(3):
        assert abs(bpp[j] - ref_bpp[j]) <= 0.005 * ref_bpp[j] + 1e-3, (j, bpp[j], ref_bpp[j])
        assert abs(ps[j] - ref_psnr[j]) <= 0.02, (j, ps[j], ref_psnr[j])
    assert bpp[0] <= bpp[1] <= bpp[2]
    for j in range(3):
        lines = open(os.path.join(tmp_path, f"level_{j}_.txt")).read().strip().splitlines()
        assert len(lines) == 3 and lines[0].startswith("SEQUENCE image0 BITS ") and " PSNR " in lines[0]
        assert lines[-1].startswith("SEQUENCE AVG BITS ") and " YPSNR " in lines[-1]
