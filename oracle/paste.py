"""numpy restatement of the ABR mixup / mosaic paste -- TEST INFRASTRUCTURE ONLY.

Follows data/datasets/voc_abr.py:512-858 of the reference (``PascalVOCDataset_ABR``): same random
draws in the same order (Python ``random`` + ``torch.distributions.Beta``), same integer coordinate
arithmetic, same pixel arithmetic (float64 blend truncated to uint8; float32 canvas truncated to
uint8).  Prototypes come from an in-memory list of ``(file_name, PIL.Image)`` instead of JPEG files;
everything else is observable behaviour of the reference.  Pinned by tests/golden/paste_*.npz, which
were produced by running the reference's own ``_start_mixup`` / ``_start_boxes_mosaic``
(tests/golden/make_golden.py).
"""
from __future__ import annotations

import os
import random

import numpy as np
import torch
from PIL import Image

from . import paste_copy, paste_mixup


class BoxRehearsalState:
    """The mutable fields of ``PascalVOCDataset_ABR`` the paste reads and writes
    (voc_abr.py:332-335,395-399): the shuffled prototype list, the shrinking ``boxes_index``,
    the batch size that triggers a refill and ``bg_size``."""

    def __init__(self, prototypes, batch_size, bg_size=0):
        self.names = [n for n, _ in prototypes]
        self.images = {n: im for n, im in prototypes}
        self.boxes_index = list(range(len(self.names)))
        self.batch_size = batch_size
        self.bg_size = bg_size


def sample_prototype(st: BoxRehearsalState, i: int, im_shape):
    """voc_abr.py:512-553.  Returns (resized PIL image, np.array([[0,0,w,h,cls]]), prototype id)."""
    name = st.names[st.boxes_index[i]]
    box_im = st.images[name].convert("RGB")
    cls_name, _ = os.path.splitext(name)[0].split("_")
    box_o_w, box_o_h = box_im.size
    im_mean_size = np.mean(im_shape)
    box_mean_size = np.mean(np.array([int(box_o_w), int(box_o_h)]))
    if float(im_mean_size * 0.2) <= float(box_mean_size) <= float(im_mean_size * 0.7):
        box_scale = 1.0
    else:
        box_scale = random.uniform(float(im_mean_size * 0.4), float(im_mean_size * 0.6)) / float(box_mean_size)
    box_im = box_im.resize((int(box_scale * box_o_w), int(box_scale * box_o_h)))
    gt = np.array([[0, 0, box_im.size[0], box_im.size[1], int(cls_name)]])
    return box_im, gt, st.boxes_index[i]


def compute_overlap(a, b):
    """voc_abr.py:932-954: True when the intersection covers > 0.3 of either box (+1 convention)."""
    area_b = (b[2] - b[0] + 1) * (b[3] - b[1] + 1)
    iw = np.maximum(np.minimum(a[2], b[2]) - np.maximum(a[0], b[0]) + 1, 0)
    ih = np.maximum(np.minimum(a[3], b[3]) - np.maximum(a[1], b[1]) + 1, 0)
    area_a = (a[2] - a[0] + 1) * (a[3] - a[1] + 1)
    inter = iw * ih
    return bool(inter / area_a > 0.3 or inter / area_b > 0.3)


def mixup(st: BoxRehearsalState, image, gts, alpha=2.0, beta=5.0):
    """voc_abr.py:555-698.  ``image``: HWC uint8 array (copied); ``gts``: float array [G,5]
    (x1,y1,x2,y2,label).  Returns (uint8 image, float64 gts [G',5])."""
    image = np.array(image)
    H, W = image.shape[0], image.shape[1]
    gts = np.array(gts, dtype=np.float64).reshape(-1, 5)
    do_mix = True
    if gts.shape[0] == 1:
        gw, gh = gts[0][2] - gts[0][0], gts[0][3] - gts[0][1]
        if (W - gw) < (W * 0.25) and (H - gh) < (H * 0.25):
            do_mix = False
    if do_mix:
        lam = torch.distributions.beta.Beta(alpha, beta).sample().item()
        if len(st.boxes_index) < st.batch_size:
            st.boxes_index = list(range(len(st.names)))
        done = 0
        for i in range(3):
            c_img, c_gt, b_id = sample_prototype(st, i, image.shape)
            c_img = np.ascontiguousarray(np.asarray(c_img))
            bw, bh = int(c_gt[0][2]), int(c_gt[0][3])
            pos_x = random.randint(0, int(W * 0.6))
            pos_y = random.randint(0, int(H * 0.4))
            new_gt = [pos_x, pos_y, bw + pos_x, bh + pos_y]
            tries = 0
            restart = True
            if gts.shape[0] == 0:
                raise RuntimeError("mixup with no ground truth never terminates in the reference (voc_abr.py:613)")
            while restart:
                for g in gts:
                    overlap = compute_overlap(g, new_gt)
                    if tries >= 20:
                        restart = False
                    elif tries < 10 and overlap:
                        pos_x = random.randint(0, int(W * 0.6))
                        pos_y = random.randint(0, int(H * 0.4))
                        new_gt = [pos_x, pos_y, bw + pos_x, bh + pos_y]
                        tries += 1
                        restart = True
                        break
                    elif 10 <= tries < 20 and overlap:
                        pos_x = random.randint(int(W * 0.4), W)
                        pos_y = random.randint(int(H * 0.6), H)
                        new_gt = [pos_x - bw, pos_y - bh, pos_x, pos_y]
                        tries += 1
                        restart = True
                        break
                    else:
                        restart = False
            if tries < 20:
                a = b = c = d = 0
                if new_gt[3] >= H:
                    a, new_gt[3] = new_gt[3] - H, H
                if new_gt[2] >= W:
                    b, new_gt[2] = new_gt[2] - W, W
                if new_gt[0] < 0:
                    c, new_gt[0] = -new_gt[0], 0
                if new_gt[1] < 0:
                    d, new_gt[1] = -new_gt[1], 0
                # source window selected by the seven branches of voc_abr.py:663-678
                if a == 0 and b == 0:
                    sy0, sx0, sy1, sx1 = d, c, bh, bw
                elif a == 0:
                    sy0, sx0, sy1, sx1 = 0, 0, bh, bw - b
                elif b == 0:
                    sy0, sx0, sy1, sx1 = 0, 0, bh - a, bw
                else:
                    sy0, sx0, sy1, sx1 = 0, 0, bh - a, bw - b
                x0, y0, x1, y1 = new_gt
                if (y1 - y0, x1 - x0) != (sy1 - sy0, sx1 - sx0):
                    raise ValueError("could not broadcast input array (shape mismatch, as numpy raises in the reference)")
                paste_mixup(image, c_img, y0, x0, y1, x1, sy0, sx0, lam)
                row = np.array([[new_gt[0], new_gt[1], new_gt[2], new_gt[3], c_gt[0][4]]], dtype=np.float64)
                gts = row if gts.shape[0] == 0 else np.insert(gts, 0, values=row, axis=0)
                if b_id in st.boxes_index:
                    st.boxes_index.remove(b_id)
            done += 1
            if done >= 2:
                break
    return image, gts


def mosaic(st: BoxRehearsalState, image_size, num_boxes=4):
    """voc_abr.py:700-816 with ``targets=[]`` as called at :841.  ``image_size`` is PIL's (W, H) of the
    current image (its pixels are discarded).  Returns (uint8 canvas [s,s,3], gts [G,5])."""
    s = int(np.mean(image_size))
    yc = int(random.uniform(s * 0.4, s * 0.6))
    xc = int(random.uniform(s * 0.4, s * 0.6))
    if len(st.boxes_index) < st.batch_size:
        st.boxes_index = list(range(len(st.names)))
    picks = [sample_prototype(st, i, image_size) for i in range(num_boxes)]
    canvas = None
    gt4 = []
    for i, (img, target, b_id) in enumerate(picks):
        w, h = img.size
        if i % 4 == 0:  # top right
            xc_, yc_ = xc + st.bg_size, yc - st.bg_size
            canvas = np.full((s, s, 3), 114, dtype=np.uint8)
            x1a, y1a, x2a, y2a = xc_, max(yc_ - h, 0), min(xc_ + w, s), yc_
            x1b, y1b, x2b, y2b = 0, h - (y2a - y1a), min(w, x2a - x1a), h
        elif i % 4 == 1:  # bottom left
            xc_, yc_ = xc - st.bg_size, yc + st.bg_size
            x1a, y1a, x2a, y2a = max(xc_ - w, 0), yc_, xc_, min(s, yc_ + h)
            x1b, y1b, x2b, y2b = w - (x2a - x1a), 0, max(xc_, w), min(y2a - y1a, h)
        elif i % 4 == 2:  # bottom right
            xc_, yc_ = xc + st.bg_size, yc + st.bg_size
            x1a, y1a, x2a, y2a = xc_, yc_, min(xc_ + w, s), min(s, yc_ + h)
            x1b, y1b, x2b, y2b = 0, 0, min(w, x2a - x1a), min(y2a - y1a, h)
        else:  # top left
            xc_, yc_ = xc - st.bg_size, yc - st.bg_size
            x1a, y1a, x2a, y2a = max(xc_ - w, 0), max(yc_ - h, 0), xc_, yc_
            x1b, y1b, x2b, y2b = w - (x2a - x1a), h - (y2a - y1a), w, h
        src = np.ascontiguousarray(np.asarray(img))
        x2b, y2b = min(x2b, w), min(y2b, h)  # numpy slicing clamps the stop
        if (y2a - y1a, x2a - x1a) != (y2b - y1b, x2b - x1b) or min(x1b, y1b) < 0:
            raise ValueError("could not broadcast input array (shape mismatch, as numpy raises in the reference)")
        paste_copy(canvas, src, y1a, x1a, y2a, x2a, y1b, x1b)
        padw, padh = x1a - x1b, y1a - y1b
        g = np.array(target)
        g[:, 0] += padw
        g[:, 1] += padh
        g[:, 2] += padw
        g[:, 3] += padh
        gt4.append(g)
        if b_id in st.boxes_index:
            st.boxes_index.remove(b_id)
    gt4 = np.concatenate(gt4, 0)
    for col, hi in ((0, s), (2, s), (1, s), (3, s)):
        np.clip(gt4[:, col], 0, hi, out=gt4[:, col])
    keep = [r for r in range(gt4.shape[0]) if not ((gt4[r][2] - gt4[r][0]) <= 2.0 or (gt4[r][3] - gt4[r][1]) <= 2.0)]
    return canvas, gt4[keep]


def transform_current_data_with_abr(st: BoxRehearsalState, image, gts):
    """voc_abr.py:821-858: 25 % mixup, 25 % mosaic, 50 % untouched.  ``image``: PIL image.
    Returns (kind, uint8 array, gts) with kind in {"none","mixup","mosaic"}."""
    kind = "none"
    if random.randint(0, 1) == 0:
        kind = "mixup" if random.randint(0, 1) == 0 else "mosaic"
    if kind == "mosaic":
        out, g = mosaic(st, image.size)
    elif kind == "mixup":
        out, g = mixup(st, image, gts)
    else:
        out, g = np.array(image), np.array(gts, dtype=np.float64).reshape(-1, 5)
    return kind, out, g


def as_pil(arr: np.ndarray) -> Image.Image:
    return Image.fromarray(np.uint8(arr))
