"""GPU parity of the RPN proposal path (sigmoid -> top-k -> decode -> clip -> small-box filter -> NMS -> cut) through
the reference-shaped Python API (-> C ABI -> sm_100a kernels) against the golden vectors produced by the reference's own
RPNPostProcessor and against the numpy oracle.

Index work (which anchors are selected, their order, which survive) is checked exactly.  Box coordinates and scores are
fp32 results of expf on the device vs torch's CPU exp, which differ by an ulp: a coordinate is `centre -+ exp(dw)*side/2`,
so its absolute error is an ulp of the LARGEST intermediate, not of the (clipped) result.  Stated tolerance:
|a-b| <= 4 * 2^-23 * S with S = e^bbox_xform_clip * (largest anchor side) -- 4 ulp of the largest value the decode can
form -- and |a-b| <= 1e-6 for the sigmoid scores."""
import numpy as np
import pytest
import torch

from inputs import make_anchors
from oracle import rpn as orpn

pytestmark = pytest.mark.gpu


def boxes_close(a, ref, anchors):
    a, ref = np.asarray(a, np.float64), np.asarray(ref, np.float64)
    assert a.shape == ref.shape
    side = float(max((anchors[:, 2] - anchors[:, 0]).max(), (anchors[:, 3] - anchors[:, 1]).max())) + 1.0
    tol = 4 * 2.0 ** -23 * np.exp(orpn.BBOX_XFORM_CLIP) * side
    err = np.abs(a - ref)
    assert (err <= tol).all(), "max err %g > %g" % (err.max(), tol)


def dev(x, channels_last=False):
    t = torch.as_tensor(x).cuda()
    return t.contiguous(memory_format=torch.channels_last) if channels_last else t


@pytest.mark.parametrize("channels_last", [False, True])
def test_rpn_postprocessor_golden_vs_reference_python(golden, channels_last):
    from abr_iod_b200.modeling.box_coder import BoxCoder
    from abr_iod_b200.modeling.rpn import rpn_proposals

    g = golden("rpn.npz")
    sizes = [tuple(int(v) for v in s) for s in g["image_sizes"]]
    obj, reg = dev(g["objectness"], channels_last), dev(g["box_regression"], channels_last)
    for ci, (pre, post, thr, min_size) in enumerate(g["cases"]):
        wts = tuple(g["case_weights"][ci])
        # the golden run used the reference's CPU NMS (IoU >= thr)
        props, scores, n_out = rpn_proposals(obj, reg, dev(g["anchors"]), sizes, int(pre), int(post), thr, min_size, wts,
                                             BoxCoder(wts).bbox_xform_clip, cpu_tie_rule=True)
        counts = n_out.tolist()
        for n in range(len(sizes)):
            gb, gs = g["c%d_i%d_boxes" % (ci, n)], g["c%d_i%d_scores" % (ci, n)]
            assert counts[n] == len(gb)
            boxes_close(props[n, : counts[n]].cpu().numpy(), gb, g["anchors"])
            assert np.abs(scores[n, : counts[n]].cpu().numpy() - gs).max() <= 1e-6
            assert not props[n, counts[n]:].any()  # zero padding


@pytest.mark.parametrize("channels_last", [False, True])
@pytest.mark.parametrize("N,A,H,W,pre,post,min_size", [(2, 15, 25, 38, 6000, 1000, 0), (3, 3, 40, 50, 2000, 2000, 16),
                                                        (1, 15, 50, 76, 12000, 2000, 0), (2, 9, 7, 9, 1000, 300, 0)])
def test_rpn_module_vs_oracle(channels_last, N, A, H, W, pre, post, min_size):
    from abr_iod_b200.modeling.rpn import RPNPostProcessor
    from abr_iod_b200.structures.bounding_box import BoxList

    rng = np.random.default_rng(N * 1000 + A * 10 + H)
    sizes_a = (32, 64, 128, 256, 512)[: max(1, A // 3)]
    anchors = make_anchors(H, W, 16, sizes=sizes_a, ratios=(0.5, 1.0, 2.0)[: A // len(sizes_a)])
    assert anchors.shape[0] == A * H * W
    logits = (rng.standard_normal((N, A, H, W)) * 2).astype(np.float32)
    reg = (rng.standard_normal((N, 4 * A, H, W)) * 0.3).astype(np.float32)
    sizes = [(W * 16 - 7 * i, H * 16 - 5 * i) for i in range(N)]
    pp = RPNPostProcessor(pre, post, 0.7, min_size)
    a_dev = dev(anchors)
    res = pp.forward_for_single_feature_map([BoxList(a_dev, s, "xyxy") for s in sizes], dev(logits, channels_last),
                                            dev(reg, channels_last))
    ref = orpn.rpn_proposals(logits, reg, [anchors], sizes, pre, post, 0.7, min_size)
    for n in range(N):
        rb, rs, _ = ref[n]
        assert len(res[n]) == len(rb) and res[n].size == sizes[n] and res[n].mode == "xyxy"
        boxes_close(res[n].bbox.cpu().numpy(), rb, anchors)
        assert np.abs(res[n].get_field("objectness").cpu().numpy() - rs).max() <= 1e-6


def test_rpn_selection_is_exact_and_ties_break_by_anchor_index():
    """Quantised logits: thousands of equal values straddle the top-k cut, so the index digits of the radix select run.
    The selected anchors and their order must equal the oracle's (logit desc, anchor index asc) exactly."""
    from abr_iod_b200.modeling.rpn import rpn_proposals

    rng = np.random.default_rng(5)
    N, A, H, W = 2, 6, 20, 30
    anchors = make_anchors(H, W, 16, sizes=(64, 256), ratios=(0.5, 1.0, 2.0))
    logits = np.round(rng.standard_normal((N, A, H, W)) * 2).astype(np.float32)  # ~10 distinct values
    logits[1] = 0.0                                                              # one image: every logit equal
    reg = (rng.standard_normal((N, 4 * A, H, W)) * 0.2).astype(np.float32)
    sizes = [(W * 16, H * 16)] * N
    for channels_last in (False, True):
        # nms_thresh <= 0: no NMS, so the output is the candidate list itself (boxlist_ops.py:22-23)
        _, scores, n_out, idx = rpn_proposals(dev(logits, channels_last), dev(reg, channels_last), dev(anchors), sizes, 1500,
                                              1000, 0.0, 0, return_anchor_index=True)
        cand = orpn.candidates(logits, reg, [anchors], sizes, 1500, 0)
        for n in range(N):
            assert int(n_out[n]) == 1500
            assert np.array_equal(idx[n].cpu().numpy(), cand[n][2])


def test_rpn_multilevel_forward_and_gt_proposals():
    from abr_iod_b200.modeling.rpn import RPNPostProcessor
    from abr_iod_b200.structures.bounding_box import BoxList

    rng = np.random.default_rng(9)
    N, A = 2, 3
    shapes = [(24, 32, 8), (12, 16, 16)]
    size = (256, 192)
    anchors, objs, regs = [], [], []
    for H, W, stride in shapes:
        anchors.append(make_anchors(H, W, stride, sizes=(stride * 4,), ratios=(0.5, 1.0, 2.0)))
        objs.append((rng.standard_normal((N, A, H, W)) * 2).astype(np.float32))
        regs.append((rng.standard_normal((N, 4 * A, H, W)) * 0.3).astype(np.float32))
    pp = RPNPostProcessor(500, 200, 0.7, 0, fpn_post_nms_top_n=250)
    pp.eval()
    per_image_anchors = [[BoxList(dev(a), size, "xyxy") for a in anchors] for _ in range(N)]
    out = pp(per_image_anchors, [dev(o) for o in objs], [dev(r) for r in regs])
    for n in range(N):
        ref = [orpn.rpn_proposals(o, r, [a], [size] * N, 500, 200, 0.7, 0)[n] for o, r, a in zip(objs, regs, anchors)]
        rb = np.concatenate([x[0] for x in ref], 0)
        rs = np.concatenate([x[1] for x in ref], 0)
        order = np.argsort(-rs, kind="stable")[:250]
        assert len(out[n]) == len(order)
        got = out[n].get_field("objectness").cpu().numpy()
        assert np.abs(got - rs[order]).max() <= 1e-6
        boxes_close(out[n].bbox.cpu().numpy(), rb[order], anchors[0])
    pp.train()
    targets = [BoxList(torch.tensor([[10.0, 20.0, 100.0, 120.0]]), size, "xyxy") for _ in range(N)]
    out_t = pp([[per_image_anchors[n][0]] for n in range(N)], [dev(objs[0])], [dev(regs[0])], targets)
    for n in range(N):
        assert out_t[n].bbox[-1].tolist() == [10.0, 20.0, 100.0, 120.0]
        assert float(out_t[n].get_field("objectness")[-1]) == 1.0


def test_rpn_argument_errors():
    from abr_iod_b200.modeling.rpn import rpn_proposals

    obj = torch.zeros((1, 3, 4, 5), device="cuda")
    reg = torch.zeros((1, 12, 4, 5), device="cuda")
    anchors = torch.zeros((60, 4), device="cuda")
    with pytest.raises(RuntimeError):
        rpn_proposals(obj.cpu(), reg, anchors, [(80, 64)], 10, 5, 0.7, 0)
    with pytest.raises(RuntimeError):
        rpn_proposals(obj, reg[:, :8], anchors, [(80, 64)], 10, 5, 0.7, 0)
    with pytest.raises(RuntimeError):
        rpn_proposals(obj, reg, anchors[:50], [(80, 64)], 10, 5, 0.7, 0)
    with pytest.raises(RuntimeError):  # more candidates than one CTA sorts
        big = torch.zeros((1, 3, 100, 100), device="cuda")
        rpn_proposals(big, torch.zeros((1, 12, 100, 100), device="cuda"), torch.zeros((30000, 4), device="cuda"), [(1600, 1600)],
                      20000, 5, 0.7, 0)


def test_rpn_many_images_and_per_image_anchors():
    """70 images (more than one launch group of the selection kernels and of the batched NMS), every image with its own
    anchor tensor (anchor_image_stride != 0) and its own size."""
    from abr_iod_b200.modeling.rpn import RPNPostProcessor
    from abr_iod_b200.structures.bounding_box import BoxList

    rng = np.random.default_rng(70)
    N, A, H, W = 70, 3, 6, 8
    base = make_anchors(H, W, 16, sizes=(48,), ratios=(0.5, 1.0, 2.0))
    anchors = [base + rng.uniform(-2, 2, base.shape).astype(np.float32) for _ in range(N)]
    logits = (rng.standard_normal((N, A, H, W)) * 2).astype(np.float32)
    reg = (rng.standard_normal((N, 4 * A, H, W)) * 0.3).astype(np.float32)
    sizes = [(W * 16 - (i % 5), H * 16 - (i % 3)) for i in range(N)]
    pp = RPNPostProcessor(100, 30, 0.7, 0)
    res = pp.forward_for_single_feature_map([BoxList(dev(a), s, "xyxy") for a, s in zip(anchors, sizes)], dev(logits), dev(reg))
    ref = orpn.rpn_proposals(logits, reg, anchors, sizes, 100, 30, 0.7, 0)
    assert len(res) == N
    for n in range(N):
        assert len(res[n]) == len(ref[n][0])
        boxes_close(res[n].bbox.cpu().numpy(), ref[n][0], base)
        assert np.abs(res[n].get_field("objectness").cpu().numpy() - ref[n][1]).max() <= 1e-6
