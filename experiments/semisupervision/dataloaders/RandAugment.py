"""RandAugment (Cubuk et al. 2019) as tensor ops (ref. vendors a PIL implementation from Fast-AutoAugment,
``experiments/semisupervision/dataloaders/RandAugment.py``).  ``RandAugment(n, m)`` applies ``n`` random ops at
magnitude ``m`` ∈ [0, 30] to CHW float images in [0, 1]; works on CPU or GPU tensors, whole batches at once."""
import random

import torch
import torch.nn.functional as F


def _affine(img, theta):
    grid = F.affine_grid(theta.unsqueeze(0), img.unsqueeze(0).shape, align_corners=False)
    return F.grid_sample(img.unsqueeze(0), grid, padding_mode="zeros", align_corners=False)[0]


def shear_x(img, v):
    return _affine(img, torch.tensor([[1.0, v, 0.0], [0.0, 1.0, 0.0]], device=img.device))


def shear_y(img, v):
    return _affine(img, torch.tensor([[1.0, 0.0, 0.0], [v, 1.0, 0.0]], device=img.device))


def translate_x(img, v):
    return _affine(img, torch.tensor([[1.0, 0.0, v], [0.0, 1.0, 0.0]], device=img.device))


def translate_y(img, v):
    return _affine(img, torch.tensor([[1.0, 0.0, 0.0], [0.0, 1.0, v]], device=img.device))


def rotate(img, deg):
    r = torch.deg2rad(torch.tensor(float(deg)))
    c, s = torch.cos(r).item(), torch.sin(r).item()
    return _affine(img, torch.tensor([[c, -s, 0.0], [s, c, 0.0]], device=img.device))


def auto_contrast(img, _):
    lo, hi = img.amin(dim=(1, 2), keepdim=True), img.amax(dim=(1, 2), keepdim=True)
    return torch.where(hi > lo, (img - lo) / (hi - lo).clamp(min=1e-6), img)


def invert(img, _):
    return 1.0 - img


def equalize(img, _):
    out = []
    for ch in img:
        q = (ch.clamp(0, 1) * 255).long().reshape(-1)
        hist = torch.bincount(q, minlength=256).float()
        cdf = hist.cumsum(0)
        cdf = (cdf - cdf.min()) / (cdf.max() - cdf.min()).clamp(min=1)
        out.append(cdf[q].view_as(ch))
    return torch.stack(out)


def solarize(img, v):
    return torch.where(img < v, img, 1.0 - img)


def posterize(img, bits):
    bits = max(int(bits), 1)
    q = 2 ** (8 - bits)
    return torch.floor(img * 255 / q) * q / 255


def contrast(img, v):
    m = img.mean()
    return ((img - m) * v + m).clamp(0, 1)


def color(img, v):
    g = img.mean(dim=0, keepdim=True)
    return ((img - g) * v + g).clamp(0, 1)


def brightness(img, v):
    return (img * v).clamp(0, 1)


def sharpness(img, v):
    k = torch.tensor([[1, 1, 1], [1, 5, 1], [1, 1, 1]], dtype=img.dtype, device=img.device) / 13.0
    blur = F.conv2d(img.unsqueeze(1), k.view(1, 1, 3, 3), padding=1).squeeze(1)
    return (blur + (img - blur) * v).clamp(0, 1)


def cutout(img, v):
    if v <= 0:
        return img
    _, h, w = img.shape
    size = int(v * w)
    cy, cx = random.randint(0, h - 1), random.randint(0, w - 1)
    y0, x0 = max(cy - size // 2, 0), max(cx - size // 2, 0)
    img = img.clone()
    img[:, y0:y0 + size, x0:x0 + size] = 0.5
    return img


def augment_list():
    return [(shear_x, -0.3, 0.3), (shear_y, -0.3, 0.3), (translate_x, -0.45, 0.45), (translate_y, -0.45, 0.45),
            (rotate, -30, 30), (auto_contrast, 0, 1), (invert, 0, 1), (equalize, 0, 1), (solarize, 0, 1),
            (posterize, 4, 8), (contrast, 0.1, 1.9), (color, 0.1, 1.9), (brightness, 0.1, 1.9), (sharpness, 0.1, 1.9),
            (cutout, 0, 0.2)]


class RandAugment:
    def __init__(self, n, m):
        self.n, self.m, self.ops = n, m, augment_list()

    def __call__(self, img):
        for op, lo, hi in random.choices(self.ops, k=self.n):
            img = op(img, (float(self.m) / 30) * (hi - lo) + lo)
        return img
