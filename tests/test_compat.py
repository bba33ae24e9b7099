"""The zero-edit route of INTEGRATION.md: compat.install() registers the `_C` stand-in, compat.patch_loaded() swaps the
Python entry points of loaded reference modules and every `from ... import` alias of them.  Uses stand-in modules (the
reference tree is not available on the GPU box); no GPU work is launched."""
import sys
import types


def _fake(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def test_install_and_patch_loaded_swap_entry_points_and_aliases():
    import abr_iod_b200.compat as compat

    saved = {k: v for k, v in sys.modules.items() if k.startswith("maskrcnn_benchmark") or k.startswith("tools")}
    try:
        def ref_ard(a, b, gamma=1.0):
            return "reference"

        def ref_id(soften, target, dist="l2", soften_proposal=None):
            return ("reference", dist)

        ref_id.__module__ = "maskrcnn_benchmark.distillation.distillation"

        class RefRPN(object):
            pass

        _fake("maskrcnn_benchmark")
        d = _fake("maskrcnn_benchmark.distillation.distillation", calculate_attentive_roi_feature_distillation=ref_ard,
                  calculate_roi_distillation_losses=ref_id)
        ops = _fake("maskrcnn_benchmark.structures.boxlist_ops", boxlist_nms=lambda b, *a, **k: "reference nms",
                    boxlist_iou=lambda a, b: "reference iou")
        rpn = _fake("maskrcnn_benchmark.modeling.rpn.inference", RPNPostProcessor=RefRPN, make_rpn_postprocessor=object())
        box = _fake("maskrcnn_benchmark.modeling.roi_heads.box_head.inference", PostProcessor=object(), make_roi_box_post_processor=object())
        loss = _fake("maskrcnn_benchmark.modeling.roi_heads.box_head.loss", FastRCNNLossComputation=object())
        pool = _fake("maskrcnn_benchmark.modeling.poolers", Pooler=object(), LevelMapper=object(), make_pooler=object())
        # tools/train_incremental.py:36-38 binds the functions by name at import time
        tool = _fake("tools.train_incremental", calculate_attentive_roi_feature_distillation=ref_ard,
                     calculate_roi_distillation_losses=ref_id, RPNPostProcessor=RefRPN)

        c = compat.install()
        assert sys.modules["maskrcnn_benchmark._C"] is c and sys.modules["maskrcnn_benchmark"]._C is c
        for fn in ("nms", "roi_align_forward", "roi_align_backward", "roi_pool_forward", "roi_pool_backward"):
            assert callable(getattr(c, fn))

        done = compat.patch_loaded()
        import abr_iod_b200.distillation.distillation as ours_d
        import abr_iod_b200.modeling.poolers as ours_p
        import abr_iod_b200.modeling.roi_heads.box_head.inference as ours_bi
        import abr_iod_b200.modeling.roi_heads.box_head.loss as ours_bl
        import abr_iod_b200.modeling.rpn.inference as ours_r
        import abr_iod_b200.structures.boxlist_ops as ours_o

        assert d.calculate_attentive_roi_feature_distillation is ours_d.calculate_attentive_roi_feature_distillation
        assert tool.calculate_attentive_roi_feature_distillation is ours_d.calculate_attentive_roi_feature_distillation
        # boxlist_nms / boxlist_iou dispatch on the device: CUDA BoxLists -> the kernels, CPU BoxLists -> the reference's own
        assert ops.boxlist_nms._abr_dispatch and ops.boxlist_iou._abr_dispatch
        assert ops.boxlist_nms_batched is ours_o.boxlist_nms_batched
        assert rpn.RPNPostProcessor is ours_r.RPNPostProcessor and tool.RPNPostProcessor is ours_r.RPNPostProcessor
        assert rpn.make_rpn_postprocessor is ours_r.make_rpn_postprocessor
        assert box.PostProcessor is ours_bi.PostProcessor and loss.FastRCNNLossComputation is ours_bl.FastRCNNLossComputation
        assert pool.Pooler is ours_p.Pooler
        # the legacy (non-'id') distillation keeps the reference's code; the alias in the tool module was swapped too
        assert d.calculate_roi_distillation_losses is tool.calculate_roi_distillation_losses is not ref_id
        assert d.calculate_roi_distillation_losses(None, None, dist="l2") == ("reference", "l2")
        import torch

        from abr_iod_b200.structures.bounding_box import BoxList

        cpu_boxes = BoxList(torch.tensor([[0.0, 0.0, 4.0, 4.0]]), (10, 10))
        assert ops.boxlist_iou(cpu_boxes, cpu_boxes) == "reference iou"  # voc_eval.py:125 runs on CPU BoxLists
        assert ops.boxlist_nms(cpu_boxes, 0.5) == "reference nms"
        assert len(done) >= 10
        assert compat.patch_loaded() == [] or all(isinstance(x, str) for x in compat.patch_loaded())  # idempotent
    finally:
        for k in [k for k in sys.modules if k.startswith("maskrcnn_benchmark") or k.startswith("tools")]:
            del sys.modules[k]
        sys.modules.update(saved)
