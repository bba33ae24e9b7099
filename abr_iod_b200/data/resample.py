"""Tap tables of Pillow's bicubic resampling (third-party arithmetic the reference calls through ``Image.resize``,
data/datasets/voc_abr.py:548) for ``abr_resize_bicubic_batch``.

Restates ``precompute_coeffs`` and ``normalize_coeffs_8bpc`` of Pillow's src/libImaging/Resample.c in float64 with the
same operation order (the window sum is accumulated tap by tap, as the C loop does), so the fixed-point taps -- and with
them every output byte -- equal Pillow's.  Host-side table building only; the pixels are resampled on the device."""
import functools

import numpy as np

PRECISION_BITS = 32 - 8 - 2  # Resample.c


def _bicubic(x):
    """bicubic_filter of Resample.c (a = -0.5), elementwise."""
    x = np.abs(x)
    near = ((1.5 * x - 2.5) * x) * x + 1.0
    far = (((x - 5.0) * x + 8.0) * x - 4.0) * -0.5
    return np.where(x < 1.0, near, np.where(x < 2.0, far, 0.0))


@functools.lru_cache(maxsize=4096)
def bicubic_taps(in_size, out_size):
    """(table int32 [out_size, 2 + ksize], ksize): per output coordinate {first input index, tap count, taps...}."""
    scale = in_size / out_size
    filterscale = scale if scale > 1.0 else 1.0
    support = 2.0 * filterscale
    ksize = int(np.ceil(support)) * 2 + 1
    center = (np.arange(out_size, dtype=np.float64) + 0.5) * scale
    ss = 1.0 / filterscale
    first = np.maximum((center - support + 0.5).astype(np.int64), 0)       # (int) truncation of a non-negative value ...
    first = np.where(center - support + 0.5 < 0, 0, first)                  # ... and the clamp of a negative one
    last = np.minimum((center + support + 0.5).astype(np.int64), in_size)
    count = last - first
    k = np.zeros((out_size, ksize), np.float64)
    ww = np.zeros(out_size, np.float64)
    for x in range(ksize):                                                   # tap by tap: the C loop's summation order
        on = x < count
        w = np.where(on, _bicubic((x + first - center + 0.5) * ss), 0.0)
        k[:, x] = w
        ww = np.where(on, ww + w, ww)
    nz = ww != 0.0
    k[nz] = k[nz] / ww[nz, None]
    fixed = np.trunc(np.where(k < 0, -0.5 + k * (1 << PRECISION_BITS), 0.5 + k * (1 << PRECISION_BITS))).astype(np.int32)
    fixed[np.arange(ksize)[None, :] >= count[:, None]] = 0
    table = np.concatenate([first[:, None].astype(np.int32), count[:, None].astype(np.int32), fixed], 1)
    return np.ascontiguousarray(table), ksize
