#!/usr/bin/env python
"""Kernel-route timings on the GPU box (CUDA events, warm, inputs > L2 where it matters): the separate ROIAlign / ARD ops
on the TMA-staged and on the gather-form (v2) kernels, and the fused ARD step, at configs[0] (P = 14 and P = 7) and at
the round-1 bench shapes.  Prints one JSON object per line; `python tools/v2_timing.py > gpurun_out/v2_timing.jsonl`."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from abr_iod_b200 import _lib  # noqa: E402
from abr_iod_b200.distillation.distillation import _ard_launch  # noqa: E402
from abr_iod_b200.layers.roi_align import roi_align_backward, roi_align_forward  # noqa: E402


def rois_like_bench(rng, R, B, im_w, im_h):
    cx, cy = rng.uniform(0, im_w, R), rng.uniform(0, im_h, R)
    bw, bh = rng.uniform(16, 400, R), rng.uniform(16, 400, R)
    deg = rng.uniform(0, 1, R) < 0.05
    bw[deg] = rng.uniform(0, 1, deg.sum())
    x1, x2 = np.clip(cx - bw / 2, 0, im_w - 1), np.clip(cx + bw / 2, 0, im_w - 1)
    y1, y2 = np.clip(cy - bh / 2, 0, im_h - 1), np.clip(cy + bh / 2, 0, im_h - 1)
    img = np.repeat(np.arange(B), -(-R // B))[:R]
    return np.stack([img, x1, y1, x2, y2], 1).astype(np.float32)


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / reps


def run(tag, B, C, H, W, R, P, im_w, im_h):
    rng = np.random.default_rng(0)
    dev = torch.device("cuda")
    t = torch.randn(B, C, H, W, device=dev).contiguous(memory_format=torch.channels_last)
    s = (t + 0.1 * torch.randn_like(t)).contiguous(memory_format=torch.channels_last)
    rois = torch.from_numpy(rois_like_bench(rng, R, B, im_w, im_h)).to(dev)
    scale = 1.0 / 16
    L = _lib.lib()
    res = {"config": tag, "B": B, "C": C, "H": H, "W": W, "R": R, "P": P}
    for v2 in ((0, 1) if P <= 7 else (1,)):
        _lib.set_option("roi_v2", v2)
        key = "v2" if v2 else "staged"
        f_old, plan = roi_align_forward(t, rois, scale, P, P, 0, return_plan=True)
        f_new = roi_align_forward(s, rois, scale, P, P, 0, plan=plan)
        g = torch.randn_like(f_new)
        res[key + "_fwd_with_plan_ms"] = timeit(lambda: roi_align_forward(t, rois, scale, P, P, 0))
        res[key + "_fwd_ms"] = timeit(lambda: roi_align_forward(s, rois, scale, P, P, 0, plan=plan))
        res[key + "_bwd_ms"] = timeit(
            lambda: roi_align_backward(g, rois, scale, P, P, B, C, H, W, 0, layout=_lib.ABR_NHWC, plan=plan))
        res["ard_ms"] = timeit(lambda: _ard_launch(f_old, f_new, 1.0, True))
        del f_old, f_new, g
    _lib.set_option("roi_v2", -1)
    # fused call, the three kernels together and its parts
    f_old = torch.empty((R, C, P, P), device=dev).contiguous(memory_format=torch.channels_last)
    f_new = torch.empty_like(f_old)
    gmap = torch.empty_like(s)
    loss3 = torch.empty(3, device=dev)
    nb = int(L.abr_roi_ard_fused_workspace_bytes(R, C, P, P))
    ws = torch.empty(nb, dtype=torch.uint8, device=dev)
    st = _lib.stream_ptr(dev)

    def fused(with_grad=True, has_plan=0):
        _lib.check(L.abr_roi_ard_fused(t.data_ptr(), s.data_ptr(), rois.data_ptr(), f_old.data_ptr(), f_new.data_ptr(),
                                       gmap.data_ptr() if with_grad else None, loss3.data_ptr(), B, C, H, W, R, P, P, scale, 0,
                                       1.0, 1.0, _lib.ABR_F32, _lib.ABR_NHWC, 1, ws.data_ptr(), nb, has_plan, st))

    res["fused_total_ms"] = timeit(fused)
    res["fused_plan_reused_ms"] = timeit(lambda: fused(True, 1))
    res["fused_forward_only_ms"] = timeit(lambda: fused(False, 1))
    pooled = R * C * P * P * 4
    fmap = B * C * H * W * 4
    res["algorithmic_bytes_composite"] = 3 * fmap + 60 * R + 6 * pooled
    res["fused_RoIs_per_s"] = R / (res["fused_total_ms"] * 1e-3)
    res["fused_algorithmic_GBs"] = res["algorithmic_bytes_composite"] / (res["fused_total_ms"] * 1e-3) / 1e9
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    run("configs[0] P=14", 2, 1024, 38, 63, 1024, 14, 1000, 600)
    run("configs[0] P=7", 2, 1024, 38, 63, 1024, 7, 1000, 600)
    run("configs[1] shapes P=7 (round-1 bench)", 4, 1024, 50, 76, 2048, 7, 1216, 800)
