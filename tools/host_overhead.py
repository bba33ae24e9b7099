import sys, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
from abr_iod_b200.layers.roi_align import roi_align_forward, roi_align_backward
from abr_iod_b200.distillation.distillation import _ard_launch
from abr_iod_b200 import _lib
x = torch.randn(4, 1024, 50, 76, device='cuda').contiguous(memory_format=torch.channels_last)
rois = torch.tensor([[0, 10., 10., 200., 200.]] * 16, device='cuda')
for name, fn in (("roi_align_forward", lambda: roi_align_forward(x, rois, 1/16, 7, 7, 0)),):
    for _ in range(20): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(200): fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(name, "host %.1f us/call issue, %.1f us/call incl. drain" % ((t1 - t0) / 200 * 1e6, (t2 - t0) / 200 * 1e6))
f, plan = roi_align_forward(x, rois, 1/16, 7, 7, 0, return_plan=True)
g = torch.randn_like(f)
for name, fn in (("roi_align_backward", lambda: roi_align_backward(g, rois, 1/16, 7, 7, 4, 1024, 50, 76, 0, layout=_lib.ABR_NHWC, plan=plan)),
                 ("ard", lambda: _ard_launch(f, f, 1.0, True))):
    for _ in range(20): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(200): fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(name, "host %.1f us/call issue, %.1f us/call incl. drain" % ((t1 - t0) / 200 * 1e6, (t2 - t0) / 200 * 1e6))
