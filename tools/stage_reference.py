#!/usr/bin/env python
"""Stage the reference's Python sources for the drop-in tests on the GPU box (dev container only).

    python tools/stage_reference.py            # /root/reference/maskrcnn_benchmark/**/*.py -> baseline/_ref/

`/root/reference` does not exist on the GPU box, and the drop-in claim of INTEGRATION.md ("the reference's own
layers/roi_align.py, modeling/poolers.py, structures/boxlist_ops.py ... run unchanged on top of compat.install()") can
only be proven by running those very files there.  `baseline/_ref/` is git-ignored (never committed) but travels with the
gpurun snapshot, like the built .so files.  Only `.py` files are copied (about 1 MB); nothing under baseline/_ref is ever
imported by the product."""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def stage(reference=None, dest=None):
    reference = reference or os.environ.get("ABR_REFERENCE", "/root/reference")
    dest = dest or os.path.join(ROOT, "baseline", "_ref")
    src = os.path.join(reference, "maskrcnn_benchmark")
    if not os.path.isdir(src):
        return 0
    n = 0
    for base, dirs, files in os.walk(src):
        dirs[:] = [d for d in dirs if d not in ("csrc", "__pycache__")]
        for f in files:
            if not f.endswith(".py"):
                continue
            rel = os.path.relpath(os.path.join(base, f), reference)
            out = os.path.join(dest, rel)
            os.makedirs(os.path.dirname(out), exist_ok=True)
            shutil.copyfile(os.path.join(base, f), out)
            n += 1
    return n


if __name__ == "__main__":
    print("staged %d files" % stage(*sys.argv[1:3]))
