"""compat.patch_loaded() on the REAL reference modules (not stand-ins), CPU side: after the swap the reference's
evaluation, which runs on CPU BoxLists (data/datasets/evaluation/voc/voc_eval.py:125 calls boxlist_iou on BoxLists built
from numpy), must keep working -- CUDA inputs take the kernels, CPU inputs keep the reference's own functions.  Needs the
staged sources (baseline/_ref, or /root/reference in the dev container)."""
import numpy as np
import pytest
import torch

from refmods import ReferenceModules, reference_root

pytestmark = pytest.mark.skipif(reference_root() is None, reason="reference sources not staged (tools/stage_reference.py)")


def test_patch_loaded_keeps_the_cpu_evaluation_path():
    with ReferenceModules() as ref:
        import maskrcnn_benchmark.modeling.poolers as ref_poolers
        import maskrcnn_benchmark.modeling.roi_heads.box_head.inference as ref_box_inf  # noqa: F401
        import maskrcnn_benchmark.modeling.rpn.inference as ref_rpn_inf  # noqa: F401
        import maskrcnn_benchmark.structures.boxlist_ops as ops
        from maskrcnn_benchmark.structures.bounding_box import BoxList

        original_iou = ops.boxlist_iou
        # voc_eval.py binds boxlist_iou with `from ... import` at import time, like every caller in the reference
        data = __import__("types").ModuleType("maskrcnn_benchmark.data")
        __import__("sys").modules["maskrcnn_benchmark.data"] = data
        voc_eval = ref.load_by_path("maskrcnn_benchmark/data/datasets/evaluation/voc/voc_eval.py", "maskrcnn_benchmark.ref_voc_eval")
        assert voc_eval.boxlist_iou is original_iou

        def evaluate():  # eval_detection_voc (voc_eval.py:45-160) end to end on CPU BoxLists
            rng = np.random.default_rng(0)
            a = np.sort(rng.uniform(0, 300, (6, 2, 2)), 1).reshape(6, 4)[:, [0, 2, 1, 3]].astype(np.float32)
            gt = BoxList(torch.from_numpy(a), (320, 320))
            gt.add_field("labels", torch.tensor([1, 1, 2, 2, 3, 3]))
            gt.add_field("difficult", torch.zeros(6, dtype=torch.uint8))
            pred = BoxList(torch.from_numpy(a + rng.normal(0, 3, a.shape).astype(np.float32)), (320, 320))
            pred.add_field("labels", torch.tensor([1, 1, 2, 2, 3, 1]))
            pred.add_field("scores", torch.linspace(0.9, 0.4, 6))
            return voc_eval.eval_detection_voc([pred], [gt], iou_thresh=0.5, use_07_metric=False)

        before = evaluate()
        done = ref.compat.patch_loaded()
        assert "structures.boxlist_ops.boxlist_iou" in done and "modeling.poolers.Pooler" in done
        import abr_iod_b200.modeling.poolers as ours_poolers
        from abr_iod_b200.structures import bounding_box as ours_boxes

        assert ref_poolers.Pooler is ours_poolers.Pooler
        assert voc_eval.boxlist_iou is ops.boxlist_iou and ops.boxlist_iou is not original_iou  # the alias was swapped too
        assert ours_boxes.OUTPUT_CLASS is BoxList  # fused ops now hand the reference's BoxList downstream

        # the evaluation's own call: CPU BoxLists
        rng = np.random.default_rng(0)
        a = np.sort(rng.uniform(0, 300, (6, 2, 2)), 1).reshape(6, 4)[:, [0, 2, 1, 3]].astype(np.float32)
        b = np.sort(rng.uniform(0, 300, (4, 2, 2)), 1).reshape(4, 4)[:, [0, 2, 1, 3]].astype(np.float32)
        ba, bb = BoxList(torch.from_numpy(a), (320, 320)), BoxList(torch.from_numpy(b), (320, 320))
        iou = ops.boxlist_iou(ba, bb)
        assert torch.equal(iou, original_iou(ba, bb)) and iou.device.type == "cpu"
        after = evaluate()  # the evaluation gives the same answer after the swap (and did not raise on CPU tensors)
        assert np.array_equal(np.nan_to_num(before["ap"], nan=-1), np.nan_to_num(after["ap"], nan=-1)) and before["map"] == after["map"]
        # outputs of the fused ops are the reference's BoxList: the methods its downstream code calls exist
        made = ours_boxes.make_boxlist(torch.from_numpy(a), (320, 320), "xyxy")
        assert isinstance(made, BoxList) and made.resize((640, 640)).size == (640, 640)
    from abr_iod_b200.structures import bounding_box

    assert bounding_box.OUTPUT_CLASS is None


def test_our_boxlist_methods_match_the_reference_boxlist():
    """resize / transpose / crop / clip_to_image / copy_with_fields of abr_iod_b200's BoxList against the reference's class."""
    with ReferenceModules():
        from maskrcnn_benchmark.structures.bounding_box import BoxList as RefBoxList

        from abr_iod_b200.structures.bounding_box import BoxList

        rng = np.random.default_rng(1)
        pts = np.sort(rng.uniform(-20, 340, (9, 2, 2)), 1).reshape(9, 4)[:, [0, 2, 1, 3]].astype(np.float32)
        for mode in ("xyxy", "xywh"):
            ours = BoxList(torch.from_numpy(pts.copy()), (320, 240), "xyxy").convert(mode)
            ref = RefBoxList(torch.from_numpy(pts.copy()), (320, 240), "xyxy").convert(mode)
            for o, r in ((ours, ref),):
                o.add_field("scores", torch.arange(9.0))
                r.add_field("scores", torch.arange(9.0))
            for call in (lambda x: x.resize((640, 480)), lambda x: x.resize((400, 480)), lambda x: x.transpose(0),
                         lambda x: x.transpose(1), lambda x: x.crop((10, 20, 200, 180)), lambda x: x.copy_with_fields("scores")):
                a, b = call(ours), call(ref)
                assert torch.equal(a.bbox, b.bbox) and a.size == b.size and a.mode == b.mode
                assert torch.equal(a.get_field("scores"), b.get_field("scores"))
        a = BoxList(torch.from_numpy(pts.copy()), (320, 240)).clip_to_image(remove_empty=True)
        b = RefBoxList(torch.from_numpy(pts.copy()), (320, 240)).clip_to_image(remove_empty=True)
        assert torch.equal(a.bbox, b.bbox)
