"""Pin the CPU oracle (oracle/) to the reference: committed golden vectors produced by running the
reference itself (tests/golden/make_golden.py) and, when oracle/_ref is built, the reference's
compiled CPU ops directly."""
import random

import numpy as np
import pytest
import torch

import oracle
from oracle import ard_torch, paste, pooler

from inputs import make_boxes, make_rois


def test_roi_align_forward_bit_exact_vs_reference(golden):
    g = golden("roi_align_fwd.npz")
    for P, ratio in [(7, 0), (7, 2), (14, 0), (3, 1)]:
        out = oracle.roi_align_forward(g["input"], g["rois"], 1 / 16, P, P, ratio)
        assert np.array_equal(out, g["out_p%d_r%d" % (P, ratio)])


def test_nms_cpu_flavour_index_exact_vs_reference(golden):
    g = golden("nms_cpu.npz")
    for n in (1, 63, 64, 65, 300, 1500):
        for thr in (0.5, 0.7):
            keep = oracle.nms(g["boxes_%d" % n], g["scores_%d" % n], thr, "cpu")
            assert np.array_equal(keep, g["keep_%d_t%d" % (n, int(thr * 10))])


def test_nms_flavours_differ_only_on_exact_tie():
    # two boxes whose IoU is exactly 0.5 with the +1 convention: 10x10 vs 10x10 shifted so inter=... use thr = IoU
    a = np.array([[0, 0, 9, 9], [0, 0, 9, 4]], np.float32)  # areas 100 and 50, inter 50 -> IoU 0.5
    s = np.array([0.9, 0.8], np.float32)
    assert oracle.nms(a, s, 0.5, "cuda").tolist() == [0, 1]  # IoU > 0.5 is false: both kept (csrc/cuda/nms.cu:60)
    assert oracle.nms(a, s, 0.5, "cpu").tolist() == [0]  # IoU >= 0.5: suppressed (csrc/cpu/nms_cpu.cpp:60)


def test_nms_empty_and_order():
    assert oracle.nms(np.zeros((0, 4), np.float32), np.zeros((0,), np.float32), 0.5).shape == (0,)
    b = np.array([[0, 0, 10, 10], [100, 100, 110, 110], [1, 1, 11, 11]], np.float32)
    s = np.array([0.1, 0.5, 0.9], np.float32)
    assert oracle.nms(b, s, 0.5).tolist() == [1, 2]  # ascending ORIGINAL index, not score order


@pytest.mark.skipif(not oracle.ref_available(), reason="oracle/_ref not built (needs /root/reference)")
def test_against_compiled_reference_cpu_ops():
    rng = np.random.default_rng(5)
    x = rng.standard_normal((2, 6, 25, 38)).astype(np.float32)
    rois = make_rois(rng, 64, 2, 38 * 16, 25 * 16)
    for P, ratio in [(7, 0), (14, 2), (2, 3)]:
        assert np.array_equal(oracle.roi_align_forward(x, rois, 1 / 16, P, P, ratio),
                              oracle.roi_align_forward(x, rois, 1 / 16, P, P, ratio, use_ref=True))
    b, s = make_boxes(rng, 2000)
    assert np.array_equal(oracle.nms(b, s, 0.7, "cpu"), oracle.nms(b, s, 0.7, "cpu", use_ref=True))


def test_roi_align_backward_and_roi_pool_vs_torchvision_cpu():
    """No CPU reference exists for these (csrc/ROIAlign.h:44, csrc/ROIPool.h:23); torchvision's CPU ops
    (third party, aligned=False) are the second opinion SURVEY.md section 8c names."""
    tv = pytest.importorskip("torchvision")
    rng = np.random.default_rng(6)
    B, C, H, W = 2, 4, 20, 31
    x = rng.standard_normal((B, C, H, W)).astype(np.float32)
    rois = make_rois(rng, 48, B, W * 16, H * 16)
    for P, ratio in [(7, 0), (7, 2), (14, 0)]:
        g = rng.standard_normal((len(rois), C, P, P)).astype(np.float32)
        xt = torch.from_numpy(x).requires_grad_()
        tv.ops.roi_align(xt, torch.from_numpy(rois), (P, P), 1 / 16, ratio, aligned=False).backward(torch.from_numpy(g))
        mine = oracle.roi_align_backward(g, rois, 1 / 16, P, P, B, C, H, W, ratio)
        np.testing.assert_allclose(mine, xt.grad.numpy(), rtol=1e-5, atol=1e-5)
        out, arg = oracle.roi_pool_forward(x, rois, 1 / 16, P, P)
        xt = torch.from_numpy(x).requires_grad_()
        ref = tv.ops.roi_pool(xt, torch.from_numpy(rois), (P, P), 1 / 16)
        assert np.array_equal(out, ref.detach().numpy())
        ref.backward(torch.from_numpy(g))
        np.testing.assert_allclose(oracle.roi_pool_backward(g, arg, rois, B, C, H, W), xt.grad.numpy(),
                                   rtol=1e-5, atol=1e-5)


def test_ard_vs_reference_python(golden):
    g = golden("ard.npz")
    for tag in "abc":
        fo, fn = g["fo_" + tag], g["fn_" + tag]
        for gamma in (1.0, 0.25):
            k = "%s_g%d" % (tag, int(gamma * 100))
            loss, _, _, grad = oracle.ard(fo, fn, gamma)
            # C oracle (double accumulation) against the reference evaluated in float64
            assert abs(loss - g["loss64_" + k]) <= 1e-6 * abs(g["loss64_" + k])
            scale = np.abs(g["grad64_" + k]).max()
            np.testing.assert_allclose(grad, g["grad64_" + k], rtol=1e-5, atol=1e-6 * scale)
            # torch restatement against the reference's own fp32 run: same ops, same rounding
            l32, g32 = ard_torch.ard_fwd_bwd(torch.from_numpy(fo), torch.from_numpy(fn), gamma)
            assert l32.item() == pytest.approx(float(g["loss32_" + k]), rel=1e-6)
            np.testing.assert_allclose(g32.numpy(), g["grad32_" + k], rtol=1e-5, atol=1e-6 * scale)


def test_ard_identical_inputs_give_zero():
    rng = np.random.default_rng(1)
    f = rng.standard_normal((2, 8, 7, 7)).astype(np.float32)
    loss, afd, pad, grad = oracle.ard(f, f.copy(), 1.0)
    assert loss == 0.0 and afd == 0.0 and pad == 0.0 and not grad.any()


def test_pooler_and_level_mapper_vs_reference_python(golden):
    g = golden("pooler.npz")
    feats = [g["feat_%d" % i] for i in range(4)]
    boxes = [g["boxes_0"], g["boxes_1"]]
    scales = tuple(g["scales"].tolist())
    rois = pooler.to_roi_format(boxes)
    assert np.array_equal(pooler.map_levels(rois[:, 1:], 2.0, 5.0), g["levels"])
    assert len(set(g["levels"].tolist())) == 4  # every level is exercised
    for ratio in (2, 0):
        assert np.array_equal(pooler.pooler(feats, boxes, 7, scales, ratio), g["multi_r%d" % ratio])
        assert np.array_equal(pooler.pooler([feats[2]], boxes, 7, (scales[2],), ratio), g["single_r%d" % ratio])


def test_boxlist_nms_vs_reference_python(golden):
    g = golden("boxlist_nms.npz")
    b, s, lab = g["boxes"], g["scores"], g["labels"]
    for thr, maxp in ((0.7, 50), (0.5, -1), (0.0, -1)):
        keep = pooler.boxlist_nms(b, s, thr, maxp, flavour="cpu")
        k = "xyxy_t%d_m%d" % (int(thr * 10), maxp)
        assert np.array_equal(b[keep], g["bbox_" + k])
        assert np.array_equal(s[keep], g["scores_" + k]) and np.array_equal(lab[keep], g["labels_" + k])


def _state(g, batch_size=4):
    names = [str(n) for n in g["proto_names"]]
    protos = [(n, paste.as_pil(g["proto_%02d" % i])) for i, n in enumerate(names)]
    return paste.BoxRehearsalState(protos, batch_size)


def test_paste_bit_exact_vs_reference_python(golden):
    """Replays the exact call stream that produced tests/golden/paste.npz on the reference's
    PascalVOCDataset_ABR (one dataset object, so boxes_index shrinks and refills)."""
    g = golden("paste.npz")
    st = _state(g)
    kinds = set()
    for case in g["cases"]:
        key, kind, seed = str(case).split(":")
        random.seed(int(seed))
        torch.manual_seed(int(seed))
        img, gts = g[key + "_img"], g[key + "_gts"]
        if kind == "mixup":
            out, og = paste.mixup(st, img, gts)
        elif kind == "mosaic":
            out, og = paste.mosaic(st, (img.shape[1], img.shape[0]))
        else:
            kind2, out, og = paste.transform_current_data_with_abr(st, paste.as_pil(img), gts)
            kinds.add(kind2)
        assert out.shape == g[key + "_out_img"].shape, key
        assert np.array_equal(out, g[key + "_out_img"]), key
        assert np.array_equal(og[:, :4].astype(np.float32), g[key + "_out_bbox"]), key
        assert np.array_equal(og[:, 4], g[key + "_out_labels"]), key
        assert st.boxes_index == g[key + "_index_after"].tolist(), key


def test_rpn_proposals_oracle_matches_reference_postprocessor(golden):
    """oracle/rpn.py against RPNPostProcessor.forward_for_single_feature_map run by make_golden.py: same proposals in the
    same order, coordinates and scores bit for bit (the oracle calls the same torch CPU exp / sigmoid)."""
    from oracle import rpn as orpn

    g = golden("rpn.npz")
    sizes = [tuple(int(v) for v in s) for s in g["image_sizes"]]
    for ci, (pre, post, thr, min_size) in enumerate(g["cases"]):
        res = orpn.rpn_proposals(g["objectness"], g["box_regression"], [g["anchors"]], sizes, int(pre), int(post), thr,
                                 min_size, tuple(g["case_weights"][ci]), flavour="cpu")
        for n, (boxes, scores, _) in enumerate(res):
            assert np.array_equal(boxes, g["c%d_i%d_boxes" % (ci, n)])
            assert np.array_equal(scores, g["c%d_i%d_scores" % (ci, n)])
    # the three cases exercise the post-NMS cut, the min_size filter and non-unit coder weights
    assert len(g["c0_i0_boxes"]) == 150 and len(g["c1_i0_boxes"]) < 600


def test_box_postprocess_oracle_matches_reference_postprocessor(golden):
    """oracle/box_post.py against PostProcessor.forward run by make_golden.py (threshold, per-class NMS, the
    detections_per_img cut, class-agnostic regression): boxes, scores and labels bit for bit, background of the last image."""
    from oracle import box_post as obp

    g = golden("box_post.npz")
    counts = g["counts"]
    props = np.split(g["proposals"], np.cumsum(counts)[:-1])
    sizes = [tuple(int(v) for v in s) for s in g["image_sizes"]]
    for ci, (st, nt, det, agn) in enumerate(g["cases"]):
        res, bg = obp.box_postprocess(g["class_logits"], g["box_regression"], props, sizes, st, nt, int(det),
                                      cls_agnostic_bbox_reg=bool(agn), flavour="cpu")
        for n, (b, s, lab, _) in enumerate(res):
            assert np.array_equal(b, g["c%d_i%d_boxes" % (ci, n)])
            assert np.array_equal(s, g["c%d_i%d_scores" % (ci, n)])
            assert np.array_equal(lab, g["c%d_i%d_labels" % (ci, n)])
        assert np.array_equal(bg[-1][0], g["c%d_bg_boxes" % ci]) and np.array_equal(bg[-1][1], g["c%d_bg_scores" % ci])
    assert len(g["c1_i0_scores"]) == 12  # the detections_per_img cut was exercised


def test_logit_loss_oracles_match_reference_python(golden):
    """oracle/logit_losses.py against calculate_roi_distillation_losses(dist='id') and FastRCNNLossComputation.__call__
    run by make_golden.py: losses and autograd gradients bit for bit in fp32 and fp64."""
    import torch

    from oracle import logit_losses as oll

    g = golden("logit_losses.npz")
    for tag in "abc":
        for name, dt in (("32", torch.float32), ("64", torch.float64)):
            ts = torch.from_numpy(g["id_%s_ts" % tag]).to(dt).requires_grad_(True)
            tb = torch.from_numpy(g["id_%s_tb" % tag]).to(dt).requires_grad_(True)
            tot, _, _ = oll.roi_distillation_id(torch.from_numpy(g["id_%s_ss" % tag]).to(dt), torch.from_numpy(g["id_%s_sb" % tag]).to(dt), ts, tb)
            tot.backward()
            assert tot.item() == g["id_%s_loss%s" % (tag, name)]
            assert np.array_equal(ts.grad.numpy(), g["id_%s_gs%s" % (tag, name)])
            assert np.array_equal(tb.grad.numpy(), g["id_%s_gb%s" % (tag, name)])
    for tag in "abcd":
        n_old, agn = (int(v) for v in g["frcnn_%s_cfg" % tag])
        for name, dt in (("32", torch.float32), ("64", torch.float64)):
            lg = torch.from_numpy(g["frcnn_%s_logits" % tag]).to(dt).requires_grad_(True)
            rg = torch.from_numpy(g["frcnn_%s_reg" % tag]).to(dt).requires_grad_(True)
            cls, box = oll.fastrcnn_loss(lg, rg, torch.from_numpy(g["frcnn_%s_labels" % tag]),
                                         torch.from_numpy(g["frcnn_%s_targets" % tag]).to(dt), n_old, bool(agn))
            (2 * cls + 3 * box).backward()
            assert cls.item() == g["frcnn_%s_cls%s" % (tag, name)] and box.item() == g["frcnn_%s_box%s" % (tag, name)]
            assert np.array_equal(lg.grad.numpy(), g["frcnn_%s_gl%s" % (tag, name)])
            assert np.array_equal(rg.grad.numpy(), g["frcnn_%s_gr%s" % (tag, name)])


def test_match_oracle_matches_reference_prepare_targets(golden):
    """oracle/match.py against FastRCNNLossComputation.prepare_targets / match_targets_to_proposals run by
    make_golden.py (reference Matcher, boxlist_iou, BoxCoder.encode): matched indices, labels and targets bit for bit."""
    from oracle import match as om

    g = golden("match.npz")
    ro = np.concatenate([[0], np.cumsum(g["n"])])
    go = np.concatenate([[0], np.cumsum(g["g"])])
    seen = set()
    for ci in range(2):
        high, low, *wts = g["c%d_cfg" % ci]
        for i in range(len(g["n"])):
            m, lab, t = om.prepare_targets(g["proposals"][ro[i]:ro[i + 1]], g["gt_boxes"][go[i]:go[i + 1]],
                                           g["gt_labels"][go[i]:go[i + 1]], high, low, wts)
            sl = slice(ro[i], ro[i + 1])
            assert np.array_equal(m, g["c%d_matched" % ci][sl]) and np.array_equal(lab, g["c%d_labels" % ci][sl])
            assert np.array_equal(t, g["c%d_targets" % ci][sl])
            seen |= set(np.minimum(m, 0).tolist())
    assert seen == {0, -1, -2}  # matched, background and ignored proposals all occur


def test_prototype_oracle_matches_reference_mem_sampling(golden):
    """oracle/prototype.py against Mem.mean_feature_sampling (tools/extract_memory.py) run by make_golden.py, including a
    class that has to be topped up with copies; descriptors bit for bit (same torch CPU mean)."""
    from oracle import prototype as op

    g = golden("prototype.npz")
    for c in range(3):
        assert np.array_equal(op.descriptors(g["pooled_%d" % c]), g["desc_%d" % c])
        order, _, source = op.mean_feature_ranking(list(g["desc_%d" % c]), int(g["per_cls"]))
        assert np.array_equal(source[order], g["selected_%d" % c])
    assert sorted(g["selected_1"].tolist()) == [0, 0, 1, 1, 2, 2]
