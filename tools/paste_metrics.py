#!/usr/bin/env python
"""Config 4 on one GPU: the paste metrics of bench.py alone (GPU call with host planning, H2D + kernel, CPU baselines)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

if __name__ == "__main__":
    import torch

    import bench

    print(json.dumps(bench.paste_metrics(torch.device("cuda"))))
