"""Make the reference pick up the B200 path without editing it.

    import abr_iod_b200.compat as compat; compat.install()      # before `import maskrcnn_benchmark...`

* registers :mod:`abr_iod_b200._C` as ``maskrcnn_benchmark._C`` (the native extension the reference's
  ``layers/*.py`` import), so ``layers.ROIAlign / ROIPool / nms`` and everything above them run on libabr_b200;
* after the reference modules are imported, ``patch_loaded()`` swaps the Python-level hot-path functions
  (``Pooler``, ``boxlist_nms``, ``calculate_attentive_roi_feature_distillation``) for the fused versions.
"""
import sys


def install():
    from . import _C

    sys.modules["maskrcnn_benchmark._C"] = _C
    pkg = sys.modules.get("maskrcnn_benchmark")
    if pkg is not None:
        pkg._C = _C
    return _C


def patch_loaded():
    """Replace the Python hot-path entry points of already-imported reference modules."""
    from .distillation import distillation as ard
    from .modeling import poolers
    from .structures import boxlist_ops

    done = []
    m = sys.modules.get("maskrcnn_benchmark.distillation.distillation")
    if m is not None:
        m.calculate_attentive_roi_feature_distillation = ard.calculate_attentive_roi_feature_distillation
        done.append("distillation.calculate_attentive_roi_feature_distillation")
    m = sys.modules.get("maskrcnn_benchmark.structures.boxlist_ops")
    if m is not None:
        m.boxlist_nms = boxlist_ops.boxlist_nms
        m.boxlist_nms_batched = boxlist_ops.boxlist_nms_batched
        done.append("structures.boxlist_ops.boxlist_nms")
    m = sys.modules.get("maskrcnn_benchmark.modeling.poolers")
    if m is not None:
        m.Pooler, m.LevelMapper, m.make_pooler = poolers.Pooler, poolers.LevelMapper, poolers.make_pooler
        done.append("modeling.poolers.Pooler")
    return done
