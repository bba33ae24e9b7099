"""Two calls of the RPN proposal path on config-2 shapes, for an ncu launch list
(ncu --metrics gpu__time_duration.sum --clock-control none python tools/rpn_profile.py)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

from abr_iod_b200.modeling.rpn import rpn_proposals
from inputs import make_anchors

rng = np.random.default_rng(6)
N, A, H, W = 4, 15, 50, 76
anchors = torch.from_numpy(make_anchors(H, W, 16)).cuda()
obj = torch.from_numpy((rng.standard_normal((N, A, H, W)) * 2).astype(np.float32)).cuda()
reg = torch.from_numpy((rng.standard_normal((N, 4 * A, H, W)) * 0.3).astype(np.float32)).cuda()
sizes = [(W * 16, H * 16)] * N
for pre, post in ((12000, 2000), (12000, 2000), (6000, 1000)):
    p, s, n = rpn_proposals(obj, reg, anchors, sizes, pre, post, 0.7, 0)
    torch.cuda.synchronize()
    print(pre, post, n.tolist())
