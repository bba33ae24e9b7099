"""Make the reference pick up the B200 path without editing it.

    import abr_iod_b200.compat as compat; compat.install()      # before `import maskrcnn_benchmark...`

* registers :mod:`abr_iod_b200._C` as ``maskrcnn_benchmark._C`` (the native extension the reference's
  ``layers/*.py`` import), so ``layers.ROIAlign / ROIPool / nms`` and everything above them run on libabr_b200;
* after the reference modules are imported, ``patch_loaded()`` swaps the Python-level hot-path entry points (``Pooler``,
  ``boxlist_nms`` / ``boxlist_iou``, ``RPNPostProcessor``, the box head's ``PostProcessor`` and
  ``FastRCNNLossComputation``, ``calculate_attentive_roi_feature_distillation``,
  ``calculate_roi_distillation_losses``) for the fused versions, including every ``from ... import`` alias of them.
"""
import sys


def install():
    from . import _C

    sys.modules["maskrcnn_benchmark._C"] = _C
    pkg = sys.modules.get("maskrcnn_benchmark")
    if pkg is not None:
        pkg._C = _C
    return _C


def _swap(module_name, attr, new, done):
    """Rebind ``module.attr`` and every other loaded module's name that still points at the old object (the reference
    binds hot-path functions with ``from x import f``, e.g. tools/train_incremental.py:36-38)."""
    m = sys.modules.get(module_name)
    if m is None or not hasattr(m, attr):
        return
    old = getattr(m, attr)
    if old is new:
        return
    setattr(m, attr, new)
    for name, other in list(sys.modules.items()):
        if other is None or not (name == "__main__" or name.startswith("maskrcnn_benchmark") or name.startswith("tools")):
            continue
        for k, v in list(vars(other).items()):
            if v is old:
                setattr(other, k, new)
    done.append("%s.%s" % (module_name.replace("maskrcnn_benchmark.", ""), attr))


def _by_device(cuda_fn, reference_fn):
    """Dispatch on the device of the first BoxList argument."""
    def dispatch(boxlist, *args, **kwargs):
        if boxlist.bbox.is_cuda:
            return cuda_fn(boxlist, *args, **kwargs)
        return reference_fn(boxlist, *args, **kwargs)

    dispatch.__name__ = getattr(reference_fn, "__name__", "dispatch")
    dispatch.__doc__ = cuda_fn.__doc__
    dispatch._abr_dispatch = True
    return dispatch


def patch_loaded():
    """Replace the Python hot-path entry points of already-imported reference modules (and every ``from ... import``
    alias of them).  Returns the list of names that were swapped."""
    from .distillation import distillation as dist
    from .modeling import poolers
    from .modeling.roi_heads.box_head import inference as box_inference
    from .modeling.roi_heads.box_head import loss as box_loss
    from .modeling.rpn import inference as rpn_inference
    from .structures import boxlist_ops

    done = []
    ref = "maskrcnn_benchmark."
    _swap(ref + "distillation.distillation", "calculate_attentive_roi_feature_distillation",
          dist.calculate_attentive_roi_feature_distillation, done)
    m = sys.modules.get(ref + "distillation.distillation")
    if m is not None and hasattr(m, "calculate_roi_distillation_losses"):
        original = m.calculate_roi_distillation_losses

        def calculate_roi_distillation_losses(soften_results, target_results, dist="l2", soften_proposal=None, _orig=original):
            # the 'id' setting runs fused; the legacy Faster-ILOD branch keeps the reference's own code
            if dist == "id":
                return dist_mod.calculate_roi_distillation_losses(soften_results, target_results, dist, soften_proposal)
            return _orig(soften_results, target_results, dist, soften_proposal)

        dist_mod = dist
        if getattr(original, "__module__", "").startswith("maskrcnn_benchmark"):
            _swap(ref + "distillation.distillation", "calculate_roi_distillation_losses", calculate_roi_distillation_losses, done)
    # CUDA BoxLists take the kernels; CPU BoxLists keep the reference's own functions (its evaluation calls boxlist_iou on
    # CPU BoxLists built from numpy: data/datasets/evaluation/voc/voc_eval.py:125, coco_eval.py:316 -- this package has no
    # CPU path of its own and does not add one)
    m = sys.modules.get(ref + "structures.boxlist_ops")
    if m is not None:
        for attr in ("boxlist_nms", "boxlist_iou"):
            original = getattr(m, attr, None)
            if original is None or getattr(original, "_abr_dispatch", False):
                continue
            _swap(ref + "structures.boxlist_ops", attr, _by_device(getattr(boxlist_ops, attr), original), done)
    rb = sys.modules.get(ref + "structures.bounding_box")
    if rb is not None and hasattr(rb, "BoxList"):
        from .structures import bounding_box as our_boxes

        our_boxes.OUTPUT_CLASS = rb.BoxList  # the fused ops hand the reference's own BoxList type downstream
    m = sys.modules.get(ref + "structures.boxlist_ops")
    if m is not None:
        m.boxlist_nms_batched = boxlist_ops.boxlist_nms_batched
    for attr in ("Pooler", "LevelMapper", "make_pooler"):
        _swap(ref + "modeling.poolers", attr, getattr(poolers, attr), done)
    for attr in ("RPNPostProcessor", "make_rpn_postprocessor"):
        _swap(ref + "modeling.rpn.inference", attr, getattr(rpn_inference, attr), done)
    for attr in ("PostProcessor", "make_roi_box_post_processor"):
        _swap(ref + "modeling.roi_heads.box_head.inference", attr, getattr(box_inference, attr), done)
    _swap(ref + "modeling.roi_heads.box_head.loss", "FastRCNNLossComputation", box_loss.FastRCNNLossComputation, done)
    return done
