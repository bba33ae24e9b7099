"""SURVEY 8d config 1 (ROIAlign 14x14 + ARD): teacher/student maps [2,1024,38,63] fp32, 512 RoIs per image, P = 14, the
same four-stage unit as bench.py (teacher fwd, student fwd, ARD fwd+bwd, student bwd), device-timed per stage.
A supplementary measurement (bench.py's line of record is config 2's P = 7 workload)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import json

import numpy as np
import torch

from abr_iod_b200 import _lib
from abr_iod_b200.distillation.distillation import _ard_launch
from abr_iod_b200.layers.roi_align import roi_align_backward, roi_align_forward


def main():
    P = int(sys.argv[1]) if len(sys.argv) > 1 else 14
    ratio = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rng = np.random.default_rng(0)
    B, C, H, W, per = 2, 1024, 38, 63, 512
    t_np = rng.standard_normal((B, C, H, W)).astype(np.float32)
    s_np = (t_np + 0.1 * rng.standard_normal(t_np.shape)).astype(np.float32)
    R = B * per
    cx, cy = rng.uniform(0, 1000, R), rng.uniform(0, 600, R)
    w, h = rng.uniform(16, 400, R), rng.uniform(16, 400, R)
    deg = rng.random(R) < 0.05
    w[deg] = rng.uniform(0.1, 1.0, deg.sum())
    rois = np.stack([np.repeat(np.arange(B), per), np.clip(cx - w / 2, 0, 999), np.clip(cy - h / 2, 0, 599),
                     np.clip(cx + w / 2, 0, 999), np.clip(cy + h / 2, 0, 599)], 1).astype(np.float32)
    dev = torch.device("cuda")
    fmt = torch.channels_last
    teacher = torch.from_numpy(t_np).to(dev).contiguous(memory_format=fmt)
    student = torch.from_numpy(s_np).to(dev).contiguous(memory_format=fmt)
    r = torch.from_numpy(rois).to(dev)
    names = ("roi_align_fwd_teacher", "roi_align_fwd_student", "ard", "roi_align_bwd")

    def step(marks=None):
        def mark():
            if marks is not None:
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                marks.append(e)
        mark()
        f_old, plan = roi_align_forward(teacher, r, 1 / 16, P, P, ratio, return_plan=True)
        mark()
        f_new = roi_align_forward(student, r, 1 / 16, P, P, ratio, plan=plan)
        mark()
        loss3, g = _ard_launch(f_old, f_new, 1.0, True)
        mark()
        gin = roi_align_backward(g, r, 1 / 16, P, P, B, C, H, W, ratio, layout=_lib.ABR_NHWC, plan=plan)
        mark()
        return loss3, gin

    for _ in range(5):
        step()
    torch.cuda.synchronize()
    all_marks = []
    steps = 50
    for _ in range(steps):
        m = []
        step(m)
        all_marks.append(m)
    torch.cuda.synchronize()
    per_k = {n: sum(m[i].elapsed_time(m[i + 1]) for m in all_marks) / steps for i, n in enumerate(names)}
    total = sum(per_k.values())
    s4 = 4
    alg = 3 * B * C * H * W * s4 + 60 * R + 6 * R * C * P * P * s4
    print(json.dumps({"config": "config 1: [2,1024,38,63] fp32, 1024 RoIs, P=%d, sampling_ratio=%d" % (P, ratio),
                      "RoIs/s": round(R / (total * 1e-3)), "ms_per_step": round(total, 4),
                      "algorithmic_GB/s": round(alg / (total * 1e-3) / 1e9, 1),
                      "kernels_ms": {k: round(v, 4) for k, v in per_k.items()}}))


if __name__ == "__main__":
    main()
