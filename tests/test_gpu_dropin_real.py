"""The zero-edit drop-in, proven on the REFERENCE's own Python: its layers/roi_align.py, layers/roi_pool.py, layers/nms.py,
modeling/poolers.py (the per-level Python loop and all), structures/bounding_box.py and structures/boxlist_ops.py are
imported unmodified from the staged copy (baseline/_ref, made by tools/stage_reference.py) on top of compat.install() --
`maskrcnn_benchmark._C` is this library -- and checked against the CPU oracle on the GPU.  Then compat.patch_loaded()
swaps the Python entry points and the reference's names give the fused versions."""
import numpy as np
import pytest
import torch

import oracle
from inputs import make_boxes, make_rois
from refmods import ReferenceModules, reference_root

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(reference_root() is None, reason="reference sources not staged")]


def close(a, ref, rel=1e-5):
    a, ref = np.asarray(a, np.float64), np.asarray(ref, np.float64)
    scale = np.abs(ref).max() if ref.size else 1.0
    assert (np.abs(a - ref) <= rel * scale + rel * np.abs(ref)).all(), "max err %g (scale %g)" % (np.abs(a - ref).max(), scale)


def test_reference_layers_run_on_this_library():
    with ReferenceModules():
        from maskrcnn_benchmark import _C
        from maskrcnn_benchmark.layers import ROIAlign, ROIPool, nms

        import abr_iod_b200._C as ours_C

        assert _C is ours_C
        rng = np.random.default_rng(0)
        B, C, H, W, P = 2, 48, 25, 38, 7
        x = rng.standard_normal((B, C, H, W)).astype(np.float32)
        rois = make_rois(rng, 40, B, W * 16, H * 16)
        xt = torch.from_numpy(x).cuda().requires_grad_(True)
        out = ROIAlign((P, P), 1 / 16, 2)(xt, torch.from_numpy(rois).cuda())  # layers/roi_align.py:12-70, unmodified
        close(out.detach().cpu().numpy(), oracle.roi_align_forward(x, rois, 1 / 16, P, P, 2))
        g = rng.standard_normal(tuple(out.shape)).astype(np.float32)
        out.backward(torch.from_numpy(g).cuda())
        close(xt.grad.cpu().numpy(), oracle.roi_align_backward(g, rois, 1 / 16, P, P, B, C, H, W, 2))
        xt.grad = None
        out = ROIPool((P, P), 1 / 16)(xt, torch.from_numpy(rois).cuda())  # layers/roi_pool.py:12-65
        ref, arg = oracle.roi_pool_forward(x, rois, 1 / 16, P, P)
        assert np.array_equal(out.detach().cpu().numpy(), ref)
        out.backward(torch.from_numpy(g).cuda())
        close(xt.grad.cpu().numpy(), oracle.roi_pool_backward(g, arg, rois, B, C, H, W))
        b, s = make_boxes(rng, 3000)
        keep = nms(torch.from_numpy(b).cuda(), torch.from_numpy(s).cuda(), 0.7)  # layers/nms.py:8
        assert np.array_equal(keep.cpu().numpy(), oracle.nms(b, s, 0.7, "cuda"))


def test_reference_pooler_and_boxlist_nms_run_on_this_library():
    from oracle import pooler as opooler

    with ReferenceModules():
        from maskrcnn_benchmark.modeling.poolers import Pooler  # modeling/poolers.py:45-105, its own multi-level loop
        from maskrcnn_benchmark.structures.bounding_box import BoxList
        from maskrcnn_benchmark.structures.boxlist_ops import boxlist_nms

        rng = np.random.default_rng(1)
        B, C, im_w, im_h = 2, 32, 640, 512
        scales = (0.25, 0.125, 0.0625, 0.03125)
        feats_np = [rng.standard_normal((B, C, int(im_h * s), int(im_w * s))).astype(np.float32) for s in scales]
        boxes_np = []
        for _ in range(B):
            x1, y1 = rng.uniform(0, im_w - 8, 25), rng.uniform(0, im_h - 8, 25)
            side = np.exp(rng.uniform(np.log(6), np.log(700), 25))
            boxes_np.append(np.stack([x1, y1, np.minimum(x1 + side, im_w - 1), np.minimum(y1 + side, im_h - 1)], 1).astype(np.float32))
        feats = [torch.from_numpy(f).cuda() for f in feats_np]
        boxes = [BoxList(torch.from_numpy(b).cuda(), (im_w, im_h), "xyxy") for b in boxes_np]
        out = Pooler((7, 7), scales, 2)(feats, boxes)
        close(out.cpu().numpy(), opooler.pooler(feats_np, boxes_np, 7, scales, 2))
        b, s = make_boxes(rng, 2000, im_w, im_h)
        bl = BoxList(torch.from_numpy(b).cuda(), (im_w, im_h), "xyxy")
        bl.add_field("scores", torch.from_numpy(s).cuda())
        kept = boxlist_nms(bl, 0.6, max_proposals=300)  # structures/boxlist_ops.py:9-31
        want = oracle.nms(b, s, 0.6, "cuda")[:300]
        assert np.array_equal(kept.bbox.cpu().numpy(), b[want]) and np.array_equal(kept.get_field("scores").cpu().numpy(), s[want])


def test_patch_loaded_on_real_modules_gives_the_fused_ops_and_reference_boxlists():
    from inputs import make_anchors
    from oracle import rpn as orpn

    with ReferenceModules() as ref:
        import maskrcnn_benchmark.distillation.distillation as ref_dist
        import maskrcnn_benchmark.modeling.rpn.inference as ref_rpn
        from maskrcnn_benchmark.modeling.box_coder import BoxCoder
        from maskrcnn_benchmark.structures.bounding_box import BoxList

        done = ref.compat.patch_loaded()
        assert "modeling.rpn.inference.RPNPostProcessor" in done
        # the reference's name now builds the batched, sync-free RPN post-processor; its outputs are reference BoxLists
        rng = np.random.default_rng(2)
        N, A, H, W = 2, 15, 20, 30
        anchors_np = make_anchors(H, W, 16)
        obj = (rng.standard_normal((N, A, H, W)) * 2).astype(np.float32)
        reg = (rng.standard_normal((N, 4 * A, H, W)) * 0.2).astype(np.float32)
        post = ref_rpn.RPNPostProcessor(pre_nms_top_n=1000, post_nms_top_n=200, nms_thresh=0.7, min_size=0,
                                        box_coder=BoxCoder(weights=(1.0, 1.0, 1.0, 1.0)), fpn_post_nms_top_n=200)
        anchors = [[BoxList(torch.from_numpy(anchors_np).cuda(), (W * 16, H * 16), "xyxy")] for _ in range(N)]
        with torch.no_grad():
            boxlists = post(anchors, [torch.from_numpy(obj).cuda()], [torch.from_numpy(reg).cuda()])
        want = orpn.rpn_proposals(obj, reg, [anchors_np], [(W * 16, H * 16)] * N, 1000, 200, 0.7, 0)
        for got, (wb, ws, _) in zip(boxlists, want):
            assert isinstance(got, BoxList) and hasattr(got, "resize")
            assert len(got) == len(wb)
            close(got.bbox.cpu().numpy(), wb, 1e-5)
            close(got.get_field("objectness").cpu().numpy(), ws, 1e-5)
        # the distillation entry point under the reference's name is the fused kernel
        f_old = torch.randn(8, 64, 7, 7, device="cuda")
        f_new = (f_old + 0.1 * torch.randn_like(f_old)).requires_grad_(True)
        loss = ref_dist.calculate_attentive_roi_feature_distillation(f_old, f_new, 1.0)
        want_loss, _, _, want_grad = oracle.ard(f_old.cpu().numpy(), f_new.detach().cpu().numpy(), 1.0)
        loss.backward()
        assert abs(loss.item() - want_loss) <= 1e-5 * abs(want_loss)
        close(f_new.grad.cpu().numpy(), want_grad, 2e-5)
