"""Image / path metrics shared by the parity tests and the fixture generators.

`pixel_mre` is the per-pixel form of BASELINE.json's "mean relative error per channel": the mean over pixels of
|a - b| / max(b, floor) after a small box blur (relative error of single 8-bit pixels is dominated by
quantisation and Monte-Carlo noise; a 5x5 box leaves spatial structure intact).  `channel_mre` is the weaker
relative error of the channel means that round 1 used; both are reported, the per-pixel one gates.
"""
import numpy as np


def psnr(a, b):
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return 99.0 if mse == 0 else float(10 * np.log10(255.0 ** 2 / mse))


def channel_mre(a, b):
    a, b = a.reshape(-1, 3).astype(np.float64), b.reshape(-1, 3).astype(np.float64)
    return np.abs(a.mean(0) - b.mean(0)) / b.mean(0)


def box_blur(img, k=5):
    """k x k box mean with edge replication (pure numpy, separable cumulative sums)."""
    if k <= 1:
        return img.astype(np.float64)
    r = k // 2
    x = np.pad(img.astype(np.float64), ((r, r), (r, r), (0, 0)), mode="edge")
    c = np.cumsum(np.pad(x, ((1, 0), (0, 0), (0, 0))), axis=0)
    x = (c[k:] - c[:-k]) / k
    c = np.cumsum(np.pad(x, ((0, 0), (1, 0), (0, 0))), axis=1)
    return (c[:, k:] - c[:, :-k]) / k


def pixel_mre(a, b, blur=5, floor=4.0):
    """mean over pixels of |a - b| / max(b, floor), per channel, on box-blurred RGB8 images"""
    fa, fb = box_blur(a, blur), box_blur(b, blur)
    return (np.abs(fa - fb) / np.maximum(fb, floor)).reshape(-1, 3).mean(0)


def tiles(img, ty=9, tx=12):
    h, w, _ = img.shape
    return img.astype(np.float64).reshape(ty, h // ty, tx, w // tx, 3).mean((1, 3))


def blocks(img, size=10):
    h, w, _ = img.shape
    return tiles(img, h // size, w // size)


def median5(lum):
    """5x5 median of a 2-D array (edge replication), numpy only"""
    p = np.pad(lum, 2, mode="edge")
    h, w = lum.shape
    stack = np.stack([p[dy:dy + h, dx:dx + w] for dy in range(5) for dx in range(5)], axis=0)
    return np.median(stack, axis=0)


def image_stats(img):
    """[mean r, g, b, noise std (luma minus its 5x5 median), fireflies (luma > median + 80), pixels == 255]"""
    img = img.astype(np.float64)
    lum = img.mean(2)
    med = median5(lum)
    return np.array([*img.reshape(-1, 3).mean(0), (lum - med).std(), float((lum > med + 80).sum()), float((img.min(2) == 255).sum())])


def compare_regions(a, b):
    """[tile mean |rel|, tile max |rel|, block rms rel] of image a against image b (50x50 tiles, 10x10 blocks)"""
    ta, tb = tiles(a), tiles(b)
    rel = np.abs(ta - tb) / tb
    ba, bb = blocks(a), blocks(b)
    brel = (ba - bb) / np.maximum(bb, 8.0)
    return np.array([rel.mean(), rel.max(), np.sqrt((brel ** 2).mean())])


def tone_curve_residual(t_img, t_ref):
    """Fit ONE quadratic tone curve ref = f(img) over the R and G tile means (the recorder's palette treats them
    alike) and return (rms, max) of the relative residual — a structural comparison that is blind to a global
    monotone tone change."""
    x, y = t_img[..., :2].ravel(), t_ref[..., :2].ravel()
    p = np.polyfit(x, y, 2)
    r = (np.polyval(p, x) - y) / y
    return float(np.sqrt((r ** 2).mean())), float(np.abs(r).max())


def path_error(Lg, Lo):
    """relative error per path between two [n, 3] radiance arrays (max over channels, floor 1e-3)"""
    return np.abs(Lg - Lo).max(axis=1) / (np.abs(Lo).max(axis=1) + 1e-3)
