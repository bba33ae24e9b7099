"""Reads the ROIAlign plans of the bench workload back from the workspace and counts vector reductions per plan mode:
how many the backward issues today vs. one per distinct footprint pixel."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

import bench
from abr_iod_b200.layers.roi_align import roi_align_forward

w = bench.WORKLOADS["configs1_p7"]
t, s, rois = bench.make_workload(w, seed=0)
x = torch.from_numpy(s).cuda().contiguous(memory_format=torch.channels_last)
out, plan = roi_align_forward(x, torch.from_numpy(rois).cuda(), w["scale"], w["P"], w["P"], w["sampling_ratio"], return_plan=True)
torch.cuda.synchronize()
PW, Hs = w["P"], w["H"]
stride = 16 + PW * 20 + Hs * 8 + 128
R = rois.shape[0]
p = plan.cpu().numpy().view(np.int32)[: R * stride].reshape(R, stride)
names = {0: "EMPTY", 1: "ROLLING", 2: "THIN", 3: "GENERIC"}
tot = {}
for r in range(R):
    mode, Y0, Y1, X0, X1 = p[r, 0], p[r, 3], p[r, 4], p[r, 8], p[r, 9]
    if mode in (0, 3):
        d = tot.setdefault(names[int(mode)], [0, 0, 0, 0]); d[0] += 1
        continue
    nrows = Y1 - Y0 + 1
    cols = p[r, 16:16 + PW * 20].reshape(PW, 20)
    rows = p[r, 16 + PW * 20:16 + PW * 20 + nrows * 8].reshape(nrows, 8)
    live_rows = int((rows[:, 0] >= 0).sum()) if mode == 1 else nrows
    emitted = int((cols[:, 1] - cols[:, 2]).sum()) if mode == 1 else int(cols[:, 1].sum())
    scheme = int(p[r, 10])
    if scheme == 2:
        emitted = int(X1 - X0 + 1)  # pixel records: one reduction per footprint pixel
    key = names[int(mode)] + {1: "+pair", 2: "+pixel records", 0: "+neither"}[scheme]
    d = tot.setdefault(key, [0, 0, 0, 0])
    d[0] += 1
    d[1] += live_rows * emitted              # reductions issued per 32*V-channel slice
    d[2] += live_rows * (X1 - X0 + 1)        # one per distinct footprint pixel
    d[3] += live_rows * int(cols[:, 1].sum())  # without any sharing
for k, (n, issued, distinct, naive) in sorted(tot.items()):
    print("%-16s RoIs %5d  reductions issued %8d  distinct pixels %8d  per-column %8d" % (k, n, issued, distinct, naive))
print("total issued %d, distinct %d" % (sum(v[1] for v in tot.values()), sum(v[2] for v in tot.values())))
