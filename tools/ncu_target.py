#!/usr/bin/env python
"""One short pass of a kernel route for `ncu` (run under gpurun):

    ncu --set full --clock-control none --import-source on -k regex:v2_ -c 6 -o gpurun_out/prof \
        python tools/ncu_target.py fused 14

Routes: `fused` (abr_roi_ard_fused: plan, teacher+student pooling, coefficients, fused backward), `separate` (teacher
forward, student forward, ARD, backward on the default kernel route), at BASELINE.json configs[0] shapes with the given
output size, or `bench1` shapes ([4,1024,50,76], 2048 RoIs) when a third argument is given."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

from v2_timing import rois_like_bench  # noqa: E402

from abr_iod_b200 import _lib  # noqa: E402
from abr_iod_b200.distillation.distillation import _ard_launch  # noqa: E402
from abr_iod_b200.layers.roi_align import roi_align_backward, roi_align_forward  # noqa: E402


def main():
    route = sys.argv[1] if len(sys.argv) > 1 else "fused"
    P = int(sys.argv[2]) if len(sys.argv) > 2 else 14
    if len(sys.argv) > 3:
        B, C, H, W, R, im_w, im_h = 4, 1024, 50, 76, 2048, 1216, 800
    else:
        B, C, H, W, R, im_w, im_h = 2, 1024, 38, 63, 1024, 1000, 600
    dev = torch.device("cuda")
    rng = np.random.default_rng(0)
    t = torch.randn(B, C, H, W, device=dev).contiguous(memory_format=torch.channels_last)
    s = (t + 0.1 * torch.randn_like(t)).contiguous(memory_format=torch.channels_last)
    rois = torch.from_numpy(rois_like_bench(rng, R, B, im_w, im_h)).to(dev)
    scale = 1.0 / 16
    L = _lib.lib()
    for _ in range(2):
        if route == "fused":
            f_old = torch.empty((R, C, P, P), device=dev).contiguous(memory_format=torch.channels_last)
            f_new = torch.empty_like(f_old)
            gmap = torch.empty_like(s)
            loss3 = torch.empty(3, device=dev)
            nb = int(L.abr_roi_ard_fused_workspace_bytes(R, C, P, P))
            ws = torch.empty(nb, dtype=torch.uint8, device=dev)
            _lib.check(L.abr_roi_ard_fused(t.data_ptr(), s.data_ptr(), rois.data_ptr(), f_old.data_ptr(), f_new.data_ptr(),
                                           gmap.data_ptr(), loss3.data_ptr(), B, C, H, W, R, P, P, scale, 0, 1.0, 1.0,
                                           _lib.ABR_F32, _lib.ABR_NHWC, 1, ws.data_ptr(), nb, 0, _lib.stream_ptr(dev)))
        else:
            f_old, plan = roi_align_forward(t, rois, scale, P, P, 0, return_plan=True)
            f_new = roi_align_forward(s, rois, scale, P, P, 0, plan=plan)
            _, g = _ard_launch(f_old, f_new, 1.0, True)
            roi_align_backward(g, rois, scale, P, P, B, C, H, W, 0, layout=_lib.ABR_NHWC, plan=plan)
        torch.cuda.synchronize()


if __name__ == "__main__":
    main()
