#!/usr/bin/env python
"""Write-only / read-only / copy bandwidth of this GPU's HBM with plain library fills and copies (CUDA events, warm):
context for the ROIAlign forward, whose DRAM traffic is 89 % writes.  Prints one JSON line."""
import json

import torch


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
    return s.elapsed_time(e) / reps * 1e-3


def main():
    n = 1683431424 // 4  # the two pooled tensors of configs[0] at P = 14, in floats
    a = torch.empty(n, device="cuda")
    b = torch.empty(n, device="cuda")
    res = {"bytes": n * 4}
    res["memset_GBs"] = n * 4 / timeit(lambda: a.zero_()) / 1e9
    res["fill_kernel_GBs"] = n * 4 / timeit(lambda: a.fill_(1.5)) / 1e9
    res["read_sum_GBs"] = n * 4 / timeit(lambda: a.sum()) / 1e9
    res["copy_GBs_read_plus_write"] = 2 * n * 4 / timeit(lambda: b.copy_(a)) / 1e9
    print(json.dumps(res))


if __name__ == "__main__":
    main()
