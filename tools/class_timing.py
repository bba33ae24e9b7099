"""Times the ROIAlign forward / backward of the bench workload per plan class (each class replicated to the same RoI
count), to see which RoIs cost the most per RoI."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

import bench
from abr_iod_b200 import _lib
from abr_iod_b200.layers.roi_align import roi_align_backward, roi_align_forward

w = bench.WORKLOADS["configs1_p7"]
t, s, rois = bench.make_workload(w, seed=0)
x = torch.from_numpy(s).cuda().contiguous(memory_format=torch.channels_last)
r_all = torch.from_numpy(rois).cuda()
P, ratio, scale = w["P"], w["sampling_ratio"], w["scale"]
out, plan = roi_align_forward(x, r_all, scale, P, P, ratio, return_plan=True)
torch.cuda.synchronize()
stride = 16 + P * 20 + w["H"] * 8 + 128
R = rois.shape[0]
p = plan.cpu().numpy().view(np.int32)[: R * stride].reshape(R, stride)
cls = {}
for r in range(R):
    key = {0: "EMPTY", 1: "ROLLING", 2: "THIN", 3: "GENERIC"}[int(p[r, 0])] + "/" + {0: "cols", 1: "pairs", 2: "pixels"}[int(p[r, 10])]
    cls.setdefault(key, []).append(r)


def time_it(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


print("all: fwd %.3f ms" % time_it(lambda: roi_align_forward(x, r_all, scale, P, P, ratio)))
for key, idx in sorted(cls.items()):
    sel = np.resize(np.asarray(idx), R)
    rr = torch.from_numpy(rois[sel]).cuda()
    f, pl = roi_align_forward(x, rr, scale, P, P, ratio, return_plan=True)
    g = torch.randn_like(f)
    tf = time_it(lambda: roi_align_forward(x, rr, scale, P, P, ratio, plan=pl))
    tb = time_it(lambda: roi_align_backward(g, rr, scale, P, P, w["B"], w["C"], w["H"], w["W"], ratio, layout=_lib.ABR_NHWC, plan=pl))
    area = np.mean([(p[i, 4] - p[i, 3] + 1) * (p[i, 9] - p[i, 8] + 1) for i in idx])
    print("%-16s %4d RoIs (%.0f px mean footprint): fwd %.3f ms, bwd %.3f ms for %d of them" % (key, len(idx), area, tf, tb, R))
