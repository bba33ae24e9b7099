"""CPU-side checks of the drop-in boundary: the shared library loads and exports exactly what include/abr_b200.h
declares, the ctypes table mirrors it, and the host wrappers refuse CPU tensors (no fallback)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "abr_b200.h")).read()
    return re.findall(r"^ABR_API [\w\s\*]+?\b(abr_\w+)\(", text, flags=re.M)


def test_header_symbols_are_exported_and_bound():
    from abr_iod_b200 import _lib

    names = _declared()
    assert len(names) >= 16 and len(set(names)) == len(names)
    L = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(L, n), "libabr_b200.so does not export %s" % n
    assert set(names) == set(_lib.SIGNATURES), "ctypes table and header disagree"
    assert _lib.lib().abr_version() == 100


def test_struct_layouts_match_header():
    from abr_iod_b200 import _lib

    assert ctypes.sizeof(_lib.PasteImage) == 24
    assert ctypes.sizeof(_lib.PasteOp) == 56
    assert _lib.PasteOp.src_offset.offset == 32 and _lib.PasteOp.lam.offset == 40


def test_no_cpu_fallback():
    from abr_iod_b200.distillation.distillation import calculate_attentive_roi_feature_distillation as ard
    from abr_iod_b200.layers import ROIAlign, ROIPool, nms

    x, rois = torch.zeros(1, 4, 8, 8), torch.zeros(1, 5)
    for fn in (lambda: ROIAlign((7, 7), 1 / 16, 0)(x, rois), lambda: ROIPool((7, 7), 1 / 16)(x, rois),
               lambda: nms(torch.zeros(2, 4), torch.zeros(2), 0.5), lambda: ard(x, x, 1.0)):
        with pytest.raises(RuntimeError, match="no CPU path"):
            fn()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "abr_iod_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "/root/reference" not in src, f
