"""Seeded synthetic inputs shared by the tests, bench.py and tests/golden/make_golden.py."""
import math

import numpy as np


def make_rois(rng, R, B, im_w, im_h, adversarial=True):
    cx, cy = rng.uniform(0, im_w, R), rng.uniform(0, im_h, R)
    w, h = rng.uniform(8, 0.6 * im_w, R), rng.uniform(8, 0.6 * im_h, R)
    x1, x2 = np.clip(cx - w / 2, 0, im_w - 1), np.clip(cx + w / 2, 0, im_w - 1)
    y1, y2 = np.clip(cy - h / 2, 0, im_h - 1), np.clip(cy + h / 2, 0, im_h - 1)
    r = np.stack([rng.integers(0, B, R), x1, y1, x2, y2], 1).astype(np.float32)
    if adversarial and R >= 10:
        r[0, 1:] = [10, 10, 10, 10]  # zero size -> forced 1x1 in feature pixels
        r[1, 1:] = [50, 50, 40, 30]  # reversed corners
        r[2, 1:] = [-100, -100, -50, -50]  # fully outside (more than 1 px)
        r[3, 1:] = [-16, -16, 4 * im_w, 4 * im_h]  # huge: large adaptive grid
        r[4, 1:] = [0, 0, im_w - 1, im_h - 1]  # whole image
        r[5, 1:] = [-16.0, 0, 100, im_h]  # exactly on the -1.0 / H boundaries at scale 1/16
        r[6, 1:] = [im_w - 1, im_h - 1, im_w + 40, im_h + 40]  # hanging off the bottom right
        r[7, 1:] = [3.3, 4.4, 3.9, 60.0]  # width < 1 px
    return r


def make_boxes(rng, n, im_w=1000, im_h=600, unique_scores=True):
    cx, cy = rng.uniform(0, im_w, n), rng.uniform(0, im_h, n)
    w, h = rng.uniform(10, 300, n), rng.uniform(10, 300, n)
    b = np.stack([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2], 1).astype(np.float32)
    half = n // 2
    if half:  # clusters of near-duplicates so that a good share is suppressed
        b[half:] = b[: n - half] + rng.normal(0, 4, (n - half, 4)).astype(np.float32)
    if unique_scores:
        s = (rng.permutation(n).astype(np.float32) + 1) / np.float32(n + 1)
    else:
        s = rng.uniform(0, 1, n).astype(np.float32)
    return b, s


def make_anchors(H, W, stride, sizes=(32, 64, 128, 256, 512), ratios=(0.5, 1.0, 2.0)):
    """A plain anchor grid in (h, w, a) order for synthetic inputs (test data, not a restatement of anchor_generator.py)."""
    base = []
    for s in sizes:
        for r in ratios:
            w, h = s / math.sqrt(r), s * math.sqrt(r)
            base.append([-(w - 1) / 2, -(h - 1) / 2, (w - 1) / 2, (h - 1) / 2])
    base = np.asarray(base, np.float32)
    ys, xs = np.meshgrid(np.arange(H, dtype=np.float32) * stride, np.arange(W, dtype=np.float32) * stride, indexing="ij")
    shifts = np.stack([xs, ys, xs, ys], -1).reshape(-1, 1, 4)
    return (shifts + base[None]).reshape(-1, 4).astype(np.float32)
