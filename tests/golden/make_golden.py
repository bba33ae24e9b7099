"""Generate tests/golden/*.npz by RUNNING THE REFERENCE (dev container only).

    python tests/golden/make_golden.py            # needs /root/reference and `make -C oracle ref`

What runs:
  * the reference's compiled CPU ops (oracle/_ref/libabr_ref_cpu.so, built in place from
    csrc/cpu/ROIAlign_cpu.cpp and csrc/cpu/nms_cpu.cpp) for ROIAlign forward and NMS;
  * the reference's own Python, imported from /root/reference through stub modules for the
    packages this image lacks (apex, maskrcnn_benchmark._C, tools.extract_memory, ...):
    distillation.calculate_attentive_roi_feature_distillation (+ torch autograd),
    modeling.poolers.Pooler / LevelMapper, structures.boxlist_ops.boxlist_nms, and
    PascalVOCDataset_ABR._start_mixup / _start_boxes_mosaic / transform_current_data_with_ABR.
The fixtures are small (KBs) and committed; this script is never imported by tests.
"""
import importlib.util
import io
import os
import random
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REFERENCE = os.environ.get("ABR_REFERENCE", "/root/reference")
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))

import oracle  # noqa: E402


# --------------------------------------------------------------------------- stubs
class _AnyAttr(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)

        def _missing(*a, **k):
            raise RuntimeError("stubbed native op %s called" % name)

        return _missing


def install_reference_stubs():
    sys.path.insert(0, REFERENCE)
    apex = types.ModuleType("apex")
    amp = types.ModuleType("apex.amp")
    amp.float_function = lambda f: f
    apex.amp = amp
    sys.modules["apex"], sys.modules["apex.amp"] = apex, amp

    C = _AnyAttr("maskrcnn_benchmark._C")

    def roi_align_forward(inp, rois, scale, ph, pw, ratio):
        out = oracle.roi_align_forward(inp.numpy(), rois.numpy(), scale, ph, pw, ratio, use_ref=True)
        return torch.from_numpy(out)

    def nms(dets, scores, thr):
        return torch.from_numpy(oracle.nms(dets.numpy(), scores.numpy(), thr, "cpu", use_ref=True))

    C.roi_align_forward = roi_align_forward
    C.nms = nms
    sys.modules["maskrcnn_benchmark._C"] = C
    import maskrcnn_benchmark

    maskrcnn_benchmark._C = C


def load_voc_abr():
    """voc_abr.py imports maskrcnn_benchmark.data (which needs `imp`, gone in 3.12) and
    tools.extract_memory; load it by path with those two stubbed."""
    data = types.ModuleType("maskrcnn_benchmark.data")
    tr = types.ModuleType("maskrcnn_benchmark.data.transforms")
    tr.Compose = object
    data.transforms = tr
    sys.modules["maskrcnn_benchmark.data"] = data
    sys.modules["maskrcnn_benchmark.data.transforms"] = tr
    tools = types.ModuleType("tools")
    em = types.ModuleType("tools.extract_memory")
    em.Mem = object
    tools.extract_memory = em
    sys.modules["tools"], sys.modules["tools.extract_memory"] = tools, em
    path = os.path.join(REFERENCE, "maskrcnn_benchmark/data/datasets/voc_abr.py")
    spec = importlib.util.spec_from_file_location("ref_voc_abr", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


from inputs import make_boxes, make_rois  # noqa: E402  (tests/inputs.py, shared with the tests)


def gen_roi_align(out):
    rng = np.random.default_rng(11)
    B, C, H, W = 2, 5, 19, 31
    x = rng.standard_normal((B, C, H, W)).astype(np.float32)
    rois = make_rois(rng, 40, B, W * 16, H * 16)
    d = {"input": x, "rois": rois}
    for P, ratio in [(7, 0), (7, 2), (14, 0), (3, 1)]:
        d["out_p%d_r%d" % (P, ratio)] = oracle.roi_align_forward(x, rois, 1 / 16, P, P, ratio, use_ref=True)
    np.savez_compressed(os.path.join(out, "roi_align_fwd.npz"), **d)


def gen_nms(out):
    rng = np.random.default_rng(12)
    d = {}
    for n in (1, 63, 64, 65, 300, 1500):
        b, s = make_boxes(rng, n)
        d["boxes_%d" % n], d["scores_%d" % n] = b, s
        for thr in (0.5, 0.7):
            d["keep_%d_t%d" % (n, int(thr * 10))] = oracle.nms(b, s, thr, "cpu", use_ref=True)
    np.savez_compressed(os.path.join(out, "nms_cpu.npz"), **d)


def gen_ard(out):
    from maskrcnn_benchmark.distillation.distillation import calculate_attentive_roi_feature_distillation as ref_ard

    rng = np.random.default_rng(13)
    d = {}
    for tag, (N, C, P, scale) in {"a": (3, 16, 7, 1.0), "b": (2, 24, 14, 0.5), "c": (2, 10, 7, 3.0)}.items():
        fo = (scale * rng.standard_normal((N, C, P, P))).astype(np.float32)
        fn = (fo + 0.1 * scale * rng.standard_normal((N, C, P, P))).astype(np.float32)
        for gamma in (1.0, 0.25):
            t_o = torch.from_numpy(fo)
            t_n = torch.from_numpy(fn).requires_grad_(True)
            loss = ref_ard(t_o, t_n, gamma)  # call-site order: (teacher, student), train_incremental.py:115
            loss.backward()
            t_n64 = torch.from_numpy(fn).double().requires_grad_(True)
            loss64 = ref_ard(t_o.double(), t_n64, gamma)
            loss64.backward()
            k = "%s_g%d" % (tag, int(gamma * 100))
            d["fo_" + tag], d["fn_" + tag] = fo, fn
            d["loss32_" + k], d["grad32_" + k] = np.float32(loss.item()), t_n.grad.numpy()
            d["loss64_" + k], d["grad64_" + k] = np.float64(loss64.item()), t_n64.grad.numpy()
    np.savez_compressed(os.path.join(out, "ard.npz"), **d)


def gen_pooler(out):
    from maskrcnn_benchmark.modeling.poolers import Pooler
    from maskrcnn_benchmark.structures.bounding_box import BoxList

    rng = np.random.default_rng(14)
    B, C = 2, 2
    im_w, im_h = 800, 640
    scales = (0.25, 0.125, 0.0625, 0.03125)
    feats = [rng.standard_normal((B, C, int(im_h * s), int(im_w * s))).astype(np.float32) for s in scales]
    boxes = []
    for b in range(B):
        n = 20 + 5 * b
        x1, y1 = rng.uniform(0, im_w - 8, n), rng.uniform(0, im_h - 8, n)
        side = np.exp(rng.uniform(np.log(6), np.log(900), n))
        x2, y2 = np.minimum(x1 + side, im_w - 1), np.minimum(y1 + side * rng.uniform(0.5, 2, n), im_h - 1)
        bx = np.stack([x1, y1, x2, y2], 1).astype(np.float32)
        big = np.array([[0, 0, im_w - 1, im_h - 1], [100, 50, 700, 600], [40, 30, 420, 390], [300, 200, 560, 470]],
                       np.float32)
        boxes.append(np.concatenate([bx, big[b::2]], 0))
    d = {"boxes_%d" % b: bx for b, bx in enumerate(boxes)}
    d.update({"feat_%d" % i: f for i, f in enumerate(feats)})
    d["scales"] = np.array(scales, np.float64)
    d["image_size"] = np.array([im_w, im_h])
    lists = [BoxList(torch.from_numpy(bx), (im_w, im_h), mode="xyxy") for bx in boxes]
    for ratio in (2, 0):
        multi = Pooler((7, 7), scales, ratio)
        d["multi_r%d" % ratio] = multi([torch.from_numpy(f) for f in feats], lists).numpy()
        single = Pooler((7, 7), (scales[2],), ratio)
        d["single_r%d" % ratio] = single([torch.from_numpy(feats[2])], lists).numpy()
    d["levels"] = multi.map_levels(lists).numpy()
    np.savez_compressed(os.path.join(out, "pooler.npz"), **d)


def gen_boxlist_nms(out):
    from maskrcnn_benchmark.structures.bounding_box import BoxList
    from maskrcnn_benchmark.structures.boxlist_ops import boxlist_nms

    rng = np.random.default_rng(15)
    b, s = make_boxes(rng, 400)
    labels = rng.integers(1, 21, 400)
    d = {"boxes": b, "scores": s, "labels": labels}
    for mode in ("xyxy", "xywh"):
        bl = BoxList(torch.from_numpy(b), (1000, 600), "xyxy").convert(mode)
        bl.add_field("scores", torch.from_numpy(s))
        bl.add_field("labels", torch.from_numpy(labels))
        for thr, maxp in ((0.7, 50), (0.5, -1), (0.0, -1)):
            r = boxlist_nms(bl, thr, max_proposals=maxp, score_field="scores")
            k = "%s_t%d_m%d" % (mode, int(thr * 10), maxp)
            d["bbox_" + k], d["scores_" + k], d["labels_" + k] = (
                r.bbox.numpy(), r.get_field("scores").numpy(), r.get_field("labels").numpy())
    np.savez_compressed(os.path.join(out, "boxlist_nms.npz"), **d)


# --------------------------------------------------------------------------- paste
def _pattern(rng, h, w):
    """Compressible but non-trivial RGB content: per-channel ramps modulo 256 plus a few flat blocks."""
    yy, xx = np.mgrid[0:h, 0:w]
    a, b, c0 = rng.integers(1, 7, 3), rng.integers(1, 7, 3), rng.integers(0, 256, 3)
    arr = ((xx[..., None] * a + yy[..., None] * b + c0) % 256).astype(np.uint8)
    for _ in range(4):
        y, x = int(rng.integers(0, h)), int(rng.integers(0, w))
        arr[y:y + h // 4, x:x + w // 4] = rng.integers(0, 256, 3)
    return arr


def make_prototypes(rng, n, lo=15, hi=130):
    """Synthetic Box-Rehearsal memory: ``{cls}_{idx:05d}.png`` -> RGB uint8 array."""
    protos = []
    for i in range(n):
        w, h = int(rng.integers(lo, hi)), int(rng.integers(lo, hi))
        protos.append(("%d_%05d.png" % (int(rng.integers(1, 16)), i), _pattern(rng, h, w)))
    return protos


def make_scene(rng, W=250, H=188, n_gt=2, big_single=False):
    img = _pattern(rng, H, W)
    if big_single:
        gts = np.array([[10.0, 8.0, W - 12.0, H - 9.0, 17.0]])
    else:
        x1, y1 = rng.uniform(0, W * 0.6, n_gt), rng.uniform(0, H * 0.6, n_gt)
        gts = np.stack([x1, y1, x1 + rng.uniform(15, W * 0.35, n_gt), y1 + rng.uniform(15, H * 0.35, n_gt),
                        rng.integers(16, 21, n_gt)], 1).astype(np.float64)
        gts[:, 2] = np.minimum(gts[:, 2], W - 1)
        gts[:, 3] = np.minimum(gts[:, 3], H - 1)
    return img, gts


def gen_paste(out):
    from PIL import Image

    voc_abr = load_voc_abr()
    from maskrcnn_benchmark.structures.bounding_box import BoxList
    import tempfile

    rng = np.random.default_rng(16)
    protos = make_prototypes(rng, 24)
    tmp = tempfile.mkdtemp(prefix="abr_protos_")
    for name, arr in protos:
        Image.fromarray(arr).save(os.path.join(tmp, name))

    def new_dataset(batch_size=4):
        ds = voc_abr.PascalVOCDataset_ABR.__new__(voc_abr.PascalVOCDataset_ABR)
        ds.PrototypeBoxSelection = types.SimpleNamespace(current_mem_path=tmp, first_mem_path=tmp)
        ds.BoxRehearsal_path = [n for n, _ in protos]
        ds.boxes_index = list(range(len(protos)))
        ds.batch_size = batch_size
        ds.bg_size = 0
        return ds

    d = {"proto_names": np.array([n for n, _ in protos])}
    for i, (_, arr) in enumerate(protos):
        d["proto_%02d" % i] = arr
    cases = []
    # a stream of calls on ONE dataset object so that boxes_index shrinks / refills like in training
    ds = new_dataset()
    for case in range(14):
        seed = 100 + case
        W, H = (250, 188) if case % 3 else (167, 250)
        img, gts = make_scene(rng, W, H, n_gt=1 + case % 3, big_single=(case == 5))
        if case in (8, 9):  # crowded image: forces the retry / bottom-right branches
            gts = np.array([[5.0, 5.0, W * 0.7, H * 0.55, 16.0], [W * 0.3, H * 0.3, W - 5.0, H - 5.0, 18.0],
                            [0.0, H * 0.5, W * 0.6, H - 1.0, 17.0]])
        target = BoxList(torch.tensor(gts[:, :4]), (W, H), mode="xyxy")
        target.add_field("labels", torch.tensor(gts[:, 4]).long())
        random.seed(seed)
        torch.manual_seed(seed)
        if case < 6 or case in (8, 9):
            kind = "mixup"
            o_img, o_t = ds._start_mixup(Image.fromarray(img), target)
        elif case < 8 or case in (10, 11):
            kind = "mosaic"
            o_img, o_t = ds._start_boxes_mosaic(Image.fromarray(img), [], num_boxes=4)
        else:
            kind = "auto"
            o_img, o_t = ds.transform_current_data_with_ABR(Image.fromarray(img), target)
        k = "case%02d" % case
        d[k + "_img"], d[k + "_gts"] = img, gts
        d[k + "_out_img"] = np.array(o_img)
        d[k + "_out_bbox"] = o_t.bbox.numpy()
        d[k + "_out_labels"] = o_t.get_field("labels").numpy()
        d[k + "_out_size"] = np.array(o_t.size)
        d[k + "_index_after"] = np.array(ds.boxes_index)
        cases.append("%s:%s:%d" % (k, kind, seed))
    d["cases"] = np.array(cases)
    np.savez_compressed(os.path.join(out, "paste.npz"), **d)


def gen_rpn(out):
    """RPNPostProcessor.forward_for_single_feature_map (modeling/rpn/inference.py:76-118) run on CPU; its NMS is the
    reference's compiled nms_cpu (IoU >= thr flavour)."""
    from maskrcnn_benchmark.modeling.box_coder import BoxCoder
    from maskrcnn_benchmark.modeling.rpn.inference import RPNPostProcessor
    from maskrcnn_benchmark.structures.bounding_box import BoxList
    from inputs import make_anchors
    from oracle import rpn as orpn

    rng = np.random.default_rng(11)
    N, A, H, W = 2, 15, 12, 17
    sizes = [(272, 192), (250, 180)]  # (width, height); the second image is smaller than the padded batch
    anchors = make_anchors(H, W, 16)
    logits = (rng.standard_normal((N, A, H, W)) * 2).astype(np.float32)
    reg = (rng.standard_normal((N, A * 4, H, W)) * 0.3).astype(np.float32)
    reg[0, 2::4][rng.random((A, H, W)) < 0.01] = 6.0  # dw beyond bbox_xform_clip
    reg[1, 3::4][rng.random((A, H, W)) < 0.01] = -5.0  # tiny boxes for the min_size filter
    assert len(np.unique(orpn.sigmoid(logits[0]))) == logits[0].size  # no ties: topk order is unambiguous
    assert len(np.unique(orpn.sigmoid(logits[1]))) == logits[1].size
    d = {"objectness": logits, "box_regression": reg, "anchors": anchors, "image_sizes": np.asarray(sizes, np.int64)}
    cases = [(1000, 150, 0.7, 0, (1.0, 1.0, 1.0, 1.0)), (600, 2000, 0.5, 8, (1.0, 1.0, 1.0, 1.0)),
             (5000, 300, 0.7, 0, (10.0, 10.0, 5.0, 5.0))]
    d["cases"] = np.asarray([c[:4] for c in cases], np.float64)
    d["case_weights"] = np.asarray([c[4] for c in cases], np.float64)
    for ci, (pre, post, thr, min_size, wts) in enumerate(cases):
        pp = RPNPostProcessor(pre, post, thr, min_size, box_coder=BoxCoder(weights=wts))
        boxlists = [BoxList(torch.from_numpy(anchors.copy()), s, "xyxy") for s in sizes]
        res = pp.forward_for_single_feature_map(boxlists, torch.from_numpy(logits), torch.from_numpy(reg))
        for n, r in enumerate(res):
            d["c%d_i%d_boxes" % (ci, n)] = r.bbox.numpy().astype(np.float32)
            d["c%d_i%d_scores" % (ci, n)] = r.get_field("objectness").numpy().astype(np.float32)
    np.savez_compressed(os.path.join(out, "rpn.npz"), **d)


def box_post_inputs(rng, sizes, counts, C):
    """Clustered proposals with peaky class logits, so that the score threshold, the per-class NMS and the
    detections_per_img cut all have something to do."""
    props, logits, regs = [], [], []
    for (w, h), n in zip(sizes, counts):
        centers = rng.uniform([0.2 * w, 0.2 * h], [0.8 * w, 0.8 * h], (6, 2))
        cls = rng.integers(1, C, 6)
        which = rng.integers(0, 6, n)
        c = centers[which] + rng.normal(0, 6, (n, 2))
        wh = rng.uniform(30, 90, (n, 2))
        b = np.stack([c[:, 0] - wh[:, 0] / 2, c[:, 1] - wh[:, 1] / 2, c[:, 0] + wh[:, 0] / 2, c[:, 1] + wh[:, 1] / 2], 1)
        b = np.clip(b, 0, [w - 1, h - 1, w - 1, h - 1]).astype(np.float32)
        lg = rng.normal(0, 1, (n, C)).astype(np.float32)
        lg[np.arange(n), cls[which]] += rng.uniform(0, 5, n).astype(np.float32)
        lg[:, 0] += rng.uniform(-1, 3, n).astype(np.float32)
        props.append(b)
        logits.append(lg)
        regs.append((rng.normal(0, 0.5, (n, 4 * C))).astype(np.float32))
    return props, np.concatenate(logits, 0), np.concatenate(regs, 0)


def gen_box_post(out):
    """PostProcessor.forward (modeling/roi_heads/box_head/inference.py:42-151) run on CPU; NMS = the reference's nms_cpu."""
    from maskrcnn_benchmark.modeling.box_coder import BoxCoder
    from maskrcnn_benchmark.modeling.roi_heads.box_head.inference import PostProcessor
    from maskrcnn_benchmark.structures.bounding_box import BoxList

    rng = np.random.default_rng(21)
    sizes = [(320, 240), (300, 200)]
    counts = [120, 90]
    C = 6
    props, logits, reg = box_post_inputs(rng, sizes, counts, C)
    d = {"image_sizes": np.asarray(sizes, np.int64), "counts": np.asarray(counts, np.int64), "class_logits": logits,
         "box_regression": reg, "proposals": np.concatenate(props, 0)}
    cases = [(0.05, 0.5, 100, False), (0.05, 0.5, 12, False), (0.3, 0.3, 100, False), (0.05, 0.5, 100, True)]
    d["cases"] = np.asarray([[c[0], c[1], c[2], float(c[3])] for c in cases], np.float64)
    for ci, (st, nt, det, agn) in enumerate(cases):
        pp = PostProcessor(st, nt, det, BoxCoder(weights=(10.0, 10.0, 5.0, 5.0)), agn)
        boxlists = [BoxList(torch.from_numpy(p.copy()), s, "xyxy") for p, s in zip(props, sizes)]
        r = reg[:, -4:] if False else reg
        results, bg = pp((torch.from_numpy(logits), torch.from_numpy(r)), boxlists)
        for n, res in enumerate(results):
            d["c%d_i%d_boxes" % (ci, n)] = res.bbox.numpy().astype(np.float32)
            d["c%d_i%d_scores" % (ci, n)] = res.get_field("scores").numpy().astype(np.float32)
            d["c%d_i%d_labels" % (ci, n)] = res.get_field("labels").numpy().astype(np.int64)
        d["c%d_bg_boxes" % ci] = bg.bbox.numpy().astype(np.float32)
        d["c%d_bg_scores" % ci] = bg.get_field("scores").numpy().astype(np.float32)
    np.savez_compressed(os.path.join(out, "box_post.npz"), **d)


def gen_logit_losses(out):
    """calculate_roi_distillation_losses(dist='id') (distillation/distillation.py:164-241) and
    FastRCNNLossComputation.__call__ (modeling/roi_heads/box_head/loss.py:122-184), fp32 and fp64, autograd gradients."""
    from maskrcnn_benchmark.distillation.distillation import calculate_roi_distillation_losses
    from maskrcnn_benchmark.modeling.roi_heads.box_head.loss import FastRCNNLossComputation
    from maskrcnn_benchmark.structures.bounding_box import BoxList

    rng = np.random.default_rng(31)
    d = {}
    # --- inclusive distillation: teacher 16 classes (15 + bg), student 21; and a 11 -> 21 case
    for tag, (R, Co, Ct) in {"a": (96, 16, 21), "b": (40, 11, 21), "c": (7, 2, 5)}.items():
        ss = (rng.standard_normal((R, Co)) * 3).astype(np.float32)
        sb = rng.standard_normal((R, Co, 4)).astype(np.float32)
        ts = (rng.standard_normal((R, Ct)) * 3).astype(np.float32)
        tb = rng.standard_normal((R, Ct, 4)).astype(np.float32)
        d["id_%s_ss" % tag], d["id_%s_sb" % tag], d["id_%s_ts" % tag], d["id_%s_tb" % tag] = ss, sb, ts, tb
        for dt, name in ((torch.float32, "32"), (torch.float64, "64")):
            t_s = torch.from_numpy(ts).to(dt).requires_grad_(True)
            t_b = torch.from_numpy(tb).to(dt).requires_grad_(True)
            loss = calculate_roi_distillation_losses((torch.from_numpy(ss).to(dt), torch.from_numpy(sb).to(dt)), (t_s, t_b), dist="id")
            loss.backward()
            d["id_%s_loss%s" % (tag, name)] = loss.detach().numpy()
            d["id_%s_gs%s" % (tag, name)] = t_s.grad.numpy()
            d["id_%s_gb%s" % (tag, name)] = t_b.grad.numpy()
    # --- box-head loss: 'id' (15 old classes) and plain cross-entropy, class-specific and class-agnostic regression
    for tag, (R, C, n_old, dist, agn) in {"a": (128, 21, 15, "id", False), "b": (64, 21, 15, "none", False),
                                          "c": (50, 11, 5, "id", True), "d": (9, 4, 0, "id", False)}.items():
        logits = (rng.standard_normal((R, C)) * 2).astype(np.float32)
        reg = rng.standard_normal((R, 8 if agn else 4 * C)).astype(np.float32) * 0.8
        allowed = np.asarray([0] + list(range(n_old + 1, C))) if dist == "id" else np.arange(C)
        labels = allowed[rng.integers(0, len(allowed), R)].astype(np.int64)
        labels[rng.random(R) < 0.5] = 0
        targets = (rng.standard_normal((R, 4)) * 0.9).astype(np.float32)
        d["frcnn_%s_logits" % tag], d["frcnn_%s_reg" % tag], d["frcnn_%s_labels" % tag], d["frcnn_%s_targets" % tag] = logits, reg, labels, targets
        d["frcnn_%s_cfg" % tag] = np.asarray([n_old if dist == "id" else -1, int(agn)], np.int64)
        for dt, name in ((torch.float32, "32"), (torch.float64, "64")):
            ev = FastRCNNLossComputation(None, None, None, agn, dist, old_classes=list(range(n_old)))
            half = R // 2
            props = []
            for a, b in ((0, half), (half, R)):
                bl = BoxList(torch.zeros((b - a, 4)), (100, 100), "xyxy")
                bl.add_field("labels", torch.from_numpy(labels[a:b]))
                bl.add_field("regression_targets", torch.from_numpy(targets[a:b]).to(dt))
                props.append(bl)
            ev._proposals = props
            t_l = torch.from_numpy(logits).to(dt).requires_grad_(True)
            t_r = torch.from_numpy(reg).to(dt).requires_grad_(True)
            cls, box = ev([t_l], [t_r])
            (2.0 * cls + 3.0 * box).backward()  # distinct upstream gradients for the two outputs
            d["frcnn_%s_cls%s" % (tag, name)], d["frcnn_%s_box%s" % (tag, name)] = cls.detach().numpy(), box.detach().numpy()
            d["frcnn_%s_gl%s" % (tag, name)], d["frcnn_%s_gr%s" % (tag, name)] = t_l.grad.numpy(), t_r.grad.numpy()
    np.savez_compressed(os.path.join(out, "logit_losses.npz"), **d)


def match_inputs(rng, size, n, G):
    """Ground-truth boxes and proposals scattered around them (so that all three Matcher outcomes occur)."""
    w, h = size
    c = rng.uniform([0.2 * w, 0.2 * h], [0.8 * w, 0.8 * h], (G, 2))
    wh = rng.uniform(30, 160, (G, 2))
    gt = np.clip(np.concatenate([c - wh / 2, c + wh / 2], 1), 0, [w - 1, h - 1, w - 1, h - 1]).astype(np.float32)
    which = rng.integers(0, G, n)
    jitter = rng.normal(0, 1, (n, 4)) * rng.choice([3.0, 15.0, 60.0], (n, 1))
    props = np.clip(gt[which] + jitter, 0, [w - 1, h - 1, w - 1, h - 1])
    props = np.stack([np.minimum(props[:, 0], props[:, 2]), np.minimum(props[:, 1], props[:, 3]),
                      np.maximum(props[:, 0], props[:, 2]) + 1, np.maximum(props[:, 1], props[:, 3]) + 1], 1).astype(np.float32)
    props[: min(G, n)] = gt[: min(G, n)]  # add_gt_proposals: exact copies of the ground truth (IoU = 1)
    return props, gt, rng.integers(1, 21, G).astype(np.int64)


def gen_match(out):
    """FastRCNNLossComputation.prepare_targets (modeling/roi_heads/box_head/loss.py:57-84) with the reference's Matcher,
    boxlist_iou and BoxCoder, on CPU."""
    from maskrcnn_benchmark.modeling.box_coder import BoxCoder
    from maskrcnn_benchmark.modeling.matcher import Matcher
    from maskrcnn_benchmark.modeling.roi_heads.box_head.loss import FastRCNNLossComputation
    from maskrcnn_benchmark.structures.bounding_box import BoxList

    rng = np.random.default_rng(41)
    sizes = [(640, 480), (500, 375), (320, 200)]
    d = {"image_sizes": np.asarray(sizes, np.int64)}
    props, gts, labs = [], [], []
    for (n, G), size in zip([(300, 5), (257, 1), (64, 12)], sizes):
        p, g, l = match_inputs(rng, size, n, G)
        props.append(p); gts.append(g); labs.append(l)
    d["n"], d["g"] = np.asarray([len(p) for p in props]), np.asarray([len(g) for g in gts])
    d["proposals"], d["gt_boxes"], d["gt_labels"] = np.concatenate(props), np.concatenate(gts), np.concatenate(labs)
    for ci, (high, low, wts) in enumerate([(0.5, 0.5, (10.0, 10.0, 5.0, 5.0)), (0.7, 0.3, (1.0, 1.0, 1.0, 1.0))]):
        ev = FastRCNNLossComputation(Matcher(high, low, allow_low_quality_matches=False), None, BoxCoder(weights=wts))
        pl = [BoxList(torch.from_numpy(p.copy()), s, "xyxy") for p, s in zip(props, sizes)]
        tl = []
        for g, l, s in zip(gts, labs, sizes):
            t = BoxList(torch.from_numpy(g.copy()), s, "xyxy")
            t.add_field("labels", torch.from_numpy(l))
            tl.append(t)
        labels, targets = ev.prepare_targets(pl, tl)
        matched = [ev.match_targets_to_proposals(p, t).get_field("matched_idxs") for p, t in zip(pl, tl)]
        d["c%d_cfg" % ci] = np.asarray([high, low] + list(wts), np.float64)
        d["c%d_labels" % ci] = torch.cat(labels).numpy()
        d["c%d_targets" % ci] = torch.cat(targets).numpy()
        d["c%d_matched" % ci] = torch.cat(matched).numpy()
    np.savez_compressed(os.path.join(out, "match.npz"), **d)


def gen_prototype(out):
    """Mem.mean_feature_sampling (tools/extract_memory.py:111-161) of the reference, loaded by file path with the config
    import stubbed, run on a hand-built Mem object; creat_and_save_box_image is replaced by a recorder."""
    cfgmod = types.ModuleType("maskrcnn_benchmark.config")
    cfgmod.cfg = None
    sys.modules["maskrcnn_benchmark.config"] = cfgmod
    spec = importlib.util.spec_from_file_location("ref_extract_memory", os.path.join(REFERENCE, "tools", "extract_memory.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    import tempfile

    rng = np.random.default_rng(51)
    d = {}
    counts = [40, 3, 17]     # the middle class has fewer boxes than num_bbox_per_cls: the top-up path
    per_cls = 6
    pooled = [(rng.standard_normal((n, 32, 7, 7)) + rng.uniform(0, 2, (1, 1, 7, 7))).astype(np.float32) for n in counts]
    saved = []
    mem = mod.Mem.__new__(mod.Mem)
    mem.num_current_classes = len(counts)
    mem.num_bbox_per_cls = per_cls
    mem.mem_size = 0
    mem.current_mem_path = tempfile.mkdtemp()
    mem.creat_and_save_box_image = lambda info, ind: saved.append((info["cls"], info["idx"], ind))
    mem.current_mem_info, mem.current_features, mem.current_logits = [], [], []
    for c, x in enumerate(pooled):
        desc = torch.mean(torch.from_numpy(x), dim=1).tolist()  # prototype_box_selection.py:97
        mem.current_mem_info.append([{"cls": c, "idx": i} for i in range(len(desc))])
        mem.current_features.append(list(desc))
        mem.current_logits.append([0] * len(desc))
        d["pooled_%d" % c] = x
        d["desc_%d" % c] = torch.mean(torch.from_numpy(x), dim=1).numpy()
    mem.mean_feature_sampling()
    d["per_cls"] = np.asarray(per_cls)
    for c in range(len(counts)):
        d["selected_%d" % c] = np.asarray([idx for cls, idx, _ in saved if cls == c], np.int64)
    np.savez_compressed(os.path.join(out, "prototype.npz"), **d)


def main():
    assert os.path.isdir(REFERENCE), "the reference tree is needed to (re)generate golden vectors"
    assert oracle.ref_available(), "run `make -C oracle ref` first"
    install_reference_stubs()
    for fn in (gen_roi_align, gen_nms, gen_ard, gen_pooler, gen_boxlist_nms, gen_paste, gen_rpn, gen_box_post, gen_logit_losses, gen_match, gen_prototype):
        fn(HERE)
        print("wrote", fn.__name__)


if __name__ == "__main__":
    main()
