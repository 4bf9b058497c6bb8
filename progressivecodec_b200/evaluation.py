"""Evaluation harness mirroring the reference's ``compress_with_ac`` protocol (training/step.py:277-404): pad to a
multiple of 64, ``compress`` then ``decompress`` at every quality level, crop, bpp from the real stream lengths, PSNR
and MS-SSIM(dB), decode wall time; optional per-level text files in the reference's line format.

The metrics are host-side bookkeeping (plain torch); the codec calls are the B200 path."""
from __future__ import annotations

import math
import os
import time
from typing import List, Optional, Sequence, Tuple, Union

import torch
import torch.nn.functional as F
from torch import Tensor


def compute_padding(in_h: int, in_w: int, *, out_h: Optional[int] = None, out_w: Optional[int] = None, min_div: int = 1):
    """compressai.ops.compute_padding (called at training/step.py:318): centred padding to a multiple of `min_div`;
    returns (pad, unpad) tuples for F.pad, ordered (left, right, top, bottom)."""
    if out_h is None:
        out_h = (in_h + min_div - 1) // min_div * min_div
    if out_w is None:
        out_w = (in_w + min_div - 1) // min_div * min_div
    if out_h % min_div != 0 or out_w % min_div != 0:
        raise ValueError(f"Padded output height and width are not divisible by min_div={min_div}.")
    left = (out_w - in_w) // 2
    right = out_w - in_w - left
    top = (out_h - in_h) // 2
    bottom = out_h - in_h - top
    return (left, right, top, bottom), (-left, -right, -top, -bottom)


def compute_psnr(a: Tensor, b: Tensor) -> float:
    """training/step.py:13-15."""
    mse = torch.mean((a - b) ** 2).item()
    return -10 * math.log10(mse)


def _gauss_window(size: int = 11, sigma: float = 1.5, device=None) -> Tensor:
    c = torch.arange(size, dtype=torch.float32, device=device) - size // 2
    g = torch.exp(-(c ** 2) / (2 * sigma ** 2))
    return g / g.sum()


def _ssim_cs(x: Tensor, y: Tensor, win: Tensor, data_range: float) -> Tuple[Tensor, Tensor]:
    C1, C2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    ch = x.shape[1]

    def blur(t):
        t = F.conv2d(t, win.view(1, 1, -1, 1).repeat(ch, 1, 1, 1), groups=ch)
        return F.conv2d(t, win.view(1, 1, 1, -1).repeat(ch, 1, 1, 1), groups=ch)

    mu1, mu2 = blur(x), blur(y)
    s11, s22, s12 = blur(x * x) - mu1 * mu1, blur(y * y) - mu2 * mu2, blur(x * y) - mu1 * mu2
    cs = (2 * s12 + C2) / (s11 + s22 + C2)
    ssim = ((2 * mu1 * mu2 + C1) / (mu1 * mu1 + mu2 * mu2 + C1)) * cs
    return ssim.flatten(2).mean(-1), cs.flatten(2).mean(-1)


def compute_msssim(a: Tensor, b: Tensor, data_range: float = 1.0) -> float:
    """pytorch_msssim.ms_ssim(a, b, data_range=1.) as used at training/step.py:17-18 (5 scales, 11-tap Gaussian,
    sigma 1.5, the standard weights); needs min(H, W) > 160."""
    weights = torch.tensor([0.0448, 0.2856, 0.3001, 0.2363, 0.1333], dtype=torch.float32, device=a.device)
    win = _gauss_window(device=a.device)
    x, y = a.float(), b.float()
    mcs = []
    for i in range(5):
        ssim, cs = _ssim_cs(x, y, win, data_range)
        if i < 4:
            mcs.append(torch.relu(cs))
            pad = [s % 2 for s in x.shape[2:]]
            x = F.avg_pool2d(x, 2, padding=pad)
            y = F.avg_pool2d(y, 2, padding=pad)
    vals = torch.stack(mcs + [torch.relu(ssim)], 0)  # [5, B, C]
    return float(torch.prod(vals ** weights.view(-1, 1, 1), 0).mean())


def read_image(path: str) -> Tensor:
    """training/step.py read_image: RGB image file -> float tensor [3,H,W] in [0,1]."""
    from PIL import Image
    import numpy as np

    img = np.asarray(Image.open(path).convert("RGB"), dtype="float32") / 255.0
    return torch.from_numpy(img).permute(2, 0, 1).contiguous()


class TestKodakDataset(torch.utils.data.Dataset):
    """datasets/utils.py:58-74: every file of `data_dir` as (transform(RGB image), path) — Kodak / CLIC test folders."""

    def __init__(self, data_dir: str, transform=None):
        if not os.path.exists(data_dir):
            raise Exception(f"[!] {data_dir} not exitd")  # (the reference's message)
        self.data_dir = data_dir
        self.transform = transform
        self.image_path = [os.path.join(data_dir, f) for f in sorted(os.listdir(data_dir))]

    def __getitem__(self, item):
        from PIL import Image

        path = self.image_path[item]
        image = Image.open(path).convert("RGB")
        if self.transform is None:
            return read_image(path), path
        return self.transform(image), path

    def __len__(self):
        return len(self.image_path)


class AverageMeter:
    """utils/functions.py AverageMeter."""

    def __init__(self):
        self.val = self.avg = self.sum = 0.0
        self.count = 0

    def update(self, val, n: int = 1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count


@torch.no_grad()
def compress_with_ac(model, filelist: Sequence[Union[str, Tensor]], device, epoch: int = -1,
                     pr_list: Sequence[float] = (0.05, 0.01), mask_pol: Optional[str] = None, writing: Optional[str] = None,
                     cheating: bool = False, with_msssim: bool = True,
                     custom_maps: Optional[Sequence[Optional[Tensor]]] = None, save_images: Optional[str] = None):
    """training/step.py:277-404 (without the wandb logging).  `filelist` holds image paths or [3,H,W] tensors.
    `custom_maps[i]` ([1, 32*n_prog, h/16, w/16] of the PADDED image, or None) replaces sigma in the masks of image i at
    every level p > 0, as `customs_maps=True` does in the reference (:300-309, :326, :334) — the reference derives the
    map from decoder gradients (`extract_dec_importance_map`), here the caller supplies it.  `save_images`: directory
    for the reconstructions `<name><level index>.png` (:345-347).
    Returns ([bpp avg per level], [psnr avg], [decode seconds avg])."""
    l = len(pr_list)
    bpp_loss = [AverageMeter() for _ in range(l)]
    psnr = [AverageMeter() for _ in range(l)]
    mssim = [AverageMeter() for _ in range(l)]
    dec_time = [AverageMeter() for _ in range(l)]
    for i, d in enumerate(filelist):
        name = os.path.splitext(os.path.basename(d))[0] if isinstance(d, str) else f"image{i}"
        x = (read_image(d) if isinstance(d, str) else d).to(device).unsqueeze(0)
        h, w = x.size(2), x.size(3)
        pad, unpad = compute_padding(h, w, min_div=2 ** 6)  # pad to allow 6 strides of 2
        x_padded = F.pad(x, pad, mode="constant", value=0)
        cmap = custom_maps[i] if custom_maps is not None else None
        if cmap is not None:
            cmap = cmap.to(device)
        for j, p in enumerate(pr_list):
            extra = {"cust_map": cmap} if (cmap is not None and p > 0) else {}
            data = model.compress(x_padded, quality=p, mask_pol=mask_pol, **extra)
            if torch.device(device).type == "cuda":
                torch.cuda.synchronize(device)
            start = time.time()
            out_dec = model.decompress(data["strings"], data["shape"], quality=p, mask_pol=mask_pol, **extra)
            if torch.device(device).type == "cuda":
                torch.cuda.synchronize(device)
            decoded_time = time.time() - start
            x_hat = F.pad(out_dec["x_hat"], unpad).clamp_(0.0, 1.0)
            if save_images is not None:
                from PIL import Image

                os.makedirs(save_images, exist_ok=True)
                img = (x_hat[0].permute(1, 2, 0).cpu() * 255.0).round().clamp_(0, 255).to(torch.uint8).numpy()
                Image.fromarray(img).save(os.path.join(save_images, f"{name}{j}.png"))
            psnr_im = compute_psnr(x, x_hat)
            ms = compute_msssim(x, x_hat) if with_msssim and min(h, w) > 160 else float("nan")
            ms_db = -10 * math.log10(1 - ms) if ms == ms and ms < 1 else float("nan")
            psnr[j].update(psnr_im)
            mssim[j].update(ms_db)
            dec_time[j].update(decoded_time)
            num_pixels = x_hat.size(0) * x_hat.size(2) * x_hat.size(3)
            bpp_scale = sum(len(s[0]) for s in data["strings"][0]) * 8.0 / num_pixels
            bpp_hype = sum(len(s) for s in data["strings"][1]) * 8.0 / num_pixels
            bpp = bpp_hype + bpp_scale if cheating is False or j == 0 else bpp_scale
            bpp_loss[j].update(bpp)
            if writing is not None:
                with open(os.path.join(writing, f"level_{j}_.txt"), "a+") as f:
                    f.write("SEQUENCE " + name + " BITS " + str(bpp) + " PSNR " + str(psnr_im) + " MSSIM " + str(ms_db) + "\n")
    if writing is not None:
        for j in range(l):
            with open(os.path.join(writing, f"level_{j}_.txt"), "a+") as f:
                f.write("SEQUENCE " + "AVG " + "BITS " + str(bpp_loss[j].avg) + " YPSNR " + str(psnr[j].avg) + " YMSSIM " +
                        str(mssim[j].avg) + "\n")
    return [m.avg for m in bpp_loss], [m.avg for m in psnr], [m.avg for m in dec_time]
