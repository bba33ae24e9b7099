"""Randomised parity sweep, part 2 (run on a GPU box): ROIPool, proposal matching, the two fused logit losses, the RPN
proposal path and the box-head post-processing against their oracles on random shapes.  Index work is compared exactly;
a run reports how many cases it checked per family and fails loudly on the first mismatch.

    python tools/fuzz_parity2.py [cases]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

import oracle
from abr_iod_b200.distillation.distillation import calculate_roi_distillation_losses
from abr_iod_b200.layers import roi_pool
from abr_iod_b200.modeling.box_coder import BoxCoder
from abr_iod_b200.modeling.matcher import Matcher
from abr_iod_b200.modeling.roi_heads.box_head import box_postprocess
from abr_iod_b200.modeling.roi_heads.box_head.loss import fastrcnn_loss, match_proposals
from abr_iod_b200.modeling.rpn import rpn_proposals
from abr_iod_b200.structures.bounding_box import BoxList
from inputs import make_anchors, make_rois
from oracle import box_post as obp
from oracle import logit_losses as oll
from oracle import match as om
from oracle import rpn as orpn


def fail(msg):
    raise SystemExit("MISMATCH " + msg)


def close(a, ref, rel, what):
    a, ref = np.asarray(a, np.float64), np.asarray(ref, np.float64)
    scale = np.abs(ref).max() if ref.size else 1.0
    err = np.abs(a - ref)
    if not (err <= rel * scale + rel * np.abs(ref)).all():
        fail("%s: max err %g (scale %g)" % (what, err.max(), scale))


def dev(x):
    return torch.as_tensor(x).cuda()


def main():
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    rng = np.random.default_rng(77)
    # ---- ROIPool
    for i in range(cases):
        B, C = int(rng.integers(1, 4)), int(rng.choice([1, 5, 16, 70]))
        H, W, P = int(rng.integers(4, 40)), int(rng.integers(4, 50)), int(rng.choice([1, 2, 7, 9]))
        R = int(rng.integers(1, 60))
        cl = bool(rng.integers(0, 2))
        x = rng.standard_normal((B, C, H, W)).astype(np.float32)
        rois = make_rois(rng, R, B, W * 16, H * 16)
        xt = dev(x)
        if cl:
            xt = xt.contiguous(memory_format=torch.channels_last)
        xt.requires_grad_(True)
        out = roi_pool(xt, dev(rois), (P, P), 1 / 16)
        ref, arg = oracle.roi_pool_forward(x, rois, 1 / 16, P, P)
        if not np.array_equal(out.detach().cpu().numpy(), ref):
            fail("roi_pool forward case %d" % i)
        g = rng.standard_normal(ref.shape).astype(np.float32)
        out.backward(dev(g))
        close(xt.grad.cpu().numpy(), oracle.roi_pool_backward(g, arg, rois, B, C, H, W), 1e-5, "roi_pool backward case %d" % i)
    print("roi_pool: %d random cases match the oracle (forward exactly)" % cases)
    # ---- multi-level Pooler (FPN), all ROIAlign layouts
    from abr_iod_b200 import _lib
    from abr_iod_b200.modeling.poolers import Pooler
    from oracle import pooler as opooler

    for i in range(cases):
        B, C = int(rng.integers(1, 3)), int(rng.choice([4, 8, 36, 128]))
        L = int(rng.integers(2, 5))
        scales = tuple(0.25 / 2 ** l for l in range(L))
        im_w, im_h = int(rng.integers(8, 40)) * 32, int(rng.integers(8, 30)) * 32
        P, ratio = int(rng.choice([3, 7, 7, 14])), int(rng.choice([0, 2]))
        feats_np = [rng.standard_normal((B, C, int(im_h * sc), int(im_w * sc))).astype(np.float32) for sc in scales]
        boxes_np = []
        for _ in range(B):
            n = int(rng.integers(1, 40))
            x1, y1 = rng.uniform(0, im_w - 8, n), rng.uniform(0, im_h - 8, n)
            side = np.exp(rng.uniform(np.log(4), np.log(1200), n))
            boxes_np.append(np.stack([x1, y1, np.minimum(x1 + side, im_w - 1), np.minimum(y1 + side * rng.uniform(0.3, 3, n), im_h - 1)], 1).astype(np.float32))
        route = int(rng.integers(0, 4))  # channels-last | contiguous staged | contiguous direct | contiguous + channels-last pooled
        _lib.NCHW_STAGING = route != 2
        _lib.POOLED_CHANNELS_LAST = route == 3
        feats = [dev(f).contiguous(memory_format=torch.channels_last) if route == 0 else dev(f) for f in feats_np]
        feats = [f.requires_grad_(True) for f in feats]
        out = Pooler((P, P), scales, ratio)(feats, [BoxList(dev(b), (im_w, im_h), "xyxy") for b in boxes_np])
        close(out.detach().cpu().numpy(), opooler.pooler(feats_np, boxes_np, P, scales, ratio), 1e-5, "pooler case %d route %d forward" % (i, route))
        g = rng.standard_normal(out.shape).astype(np.float32)
        out.backward(dev(g))
        rois = opooler.to_roi_format(boxes_np)
        k_min, k_max = -np.log2(np.float32(scales[0])), -np.log2(np.float32(scales[-1]))
        levels = opooler.map_levels(rois[:, 1:], k_min, k_max)
        for lvl in range(L):
            sel = np.nonzero(levels == lvl)[0]
            gref = oracle.roi_align_backward(g[sel], rois[sel], scales[lvl], P, P, *feats_np[lvl].shape, ratio)
            close(feats[lvl].grad.cpu().numpy(), gref, 1e-5, "pooler case %d route %d backward level %d" % (i, route, lvl))
    _lib.NCHW_STAGING, _lib.POOLED_CHANNELS_LAST = True, False
    print("pooler: %d random multi-level cases (all four layout routes) match the oracle" % cases)
    # ---- matching
    for i in range(cases):
        n_img = int(rng.integers(1, 5))
        size = (int(rng.integers(200, 1300)), int(rng.integers(200, 900)))
        props, targets, raw = [], [], []
        for _ in range(n_img):
            G, n = int(rng.integers(1, 20)), int(rng.integers(1, 600))
            c = rng.uniform([0, 0], size, (G, 2))
            wh = rng.uniform(10, 300, (G, 2))
            gt = np.clip(np.concatenate([c - wh / 2, c + wh / 2], 1), 0, [size[0] - 1, size[1] - 1] * 2).astype(np.float32)
            p = gt[rng.integers(0, G, n)] + rng.normal(0, 1, (n, 4)) * rng.choice([2.0, 20.0, 100.0], (n, 1))
            p = np.stack([np.minimum(p[:, 0], p[:, 2]), np.minimum(p[:, 1], p[:, 3]), np.maximum(p[:, 0], p[:, 2]) + 1,
                          np.maximum(p[:, 1], p[:, 3]) + 1], 1).astype(np.float32)
            p[: min(G, n)] = gt[: min(G, n)]
            lab = rng.integers(1, 21, G).astype(np.int64)
            raw.append((p, gt, lab))
            props.append(BoxList(dev(p), size, "xyxy"))
            t = BoxList(dev(gt), size, "xyxy")
            t.add_field("labels", dev(lab))
            targets.append(t)
        high = float(rng.choice([0.5, 0.7]))
        low = float(rng.choice([0.3, high]))
        wts = (10.0, 10.0, 5.0, 5.0)
        labels, reg, matched = match_proposals(props, targets, Matcher(high, low), BoxCoder(wts))
        for k, (p, gt, lab) in enumerate(raw):
            m, l_ref, t_ref = om.prepare_targets(p, gt, lab, high, low, wts)
            if not (np.array_equal(matched[k].cpu().numpy(), m) and np.array_equal(labels[k].cpu().numpy(), l_ref)):
                fail("match case %d image %d" % (i, k))
            err = np.abs(reg[k].cpu().numpy() - t_ref)
            if not (err <= 1e-5 * np.maximum(1.0, np.abs(t_ref))).all():
                fail("match targets case %d image %d: %g" % (i, k, err.max()))
    print("match: %d random batches match the oracle (indices and labels exactly)" % cases)
    # ---- logit losses
    for i in range(cases):
        R = int(rng.integers(1, 700))
        Co = int(rng.integers(1, 40))
        Ct = Co + int(rng.integers(1, 40))
        sc = float(rng.choice([0.5, 3.0, 10.0]))
        ss, ts = (rng.standard_normal((R, Co)) * sc).astype(np.float32), (rng.standard_normal((R, Ct)) * sc).astype(np.float32)
        sb, tb = rng.standard_normal((R, Co, 4)).astype(np.float32), rng.standard_normal((R, Ct, 4)).astype(np.float32)
        if Co > 1:
            t_s, t_b = dev(ts).requires_grad_(True), dev(tb).requires_grad_(True)
            loss = calculate_roi_distillation_losses((dev(ss), dev(sb)), (t_s, t_b), dist="id")
            loss.backward()
            o_s, o_b = torch.from_numpy(ts).double().requires_grad_(True), torch.from_numpy(tb).double().requires_grad_(True)
            ref, _, _ = oll.roi_distillation_id(torch.from_numpy(ss).double(), torch.from_numpy(sb).double(), o_s, o_b)
            ref.backward()
            close(loss.item(), ref.item(), 1e-5, "roi_distillation case %d loss" % i)
            close(t_s.grad.cpu().numpy(), o_s.grad.numpy(), 1e-5, "roi_distillation case %d dscores" % i)
            close(t_b.grad.cpu().numpy(), o_b.grad.numpy(), 1e-5, "roi_distillation case %d dboxes" % i)
        C = Ct
        n_old = int(rng.integers(-1, C - 1))
        agn = bool(rng.integers(0, 2))
        logits = (rng.standard_normal((R, C)) * sc).astype(np.float32)
        regr = rng.standard_normal((R, 8 if agn else 4 * C)).astype(np.float32)
        allowed = np.asarray([0] + list(range(n_old + 1, C))) if n_old >= 0 else np.arange(C)
        lab = allowed[rng.integers(0, len(allowed), R)].astype(np.int64)
        lab[rng.random(R) < 0.4] = 0
        tg = rng.standard_normal((R, 4)).astype(np.float32) * 1.5
        l_t, r_t = dev(logits).requires_grad_(True), dev(regr).requires_grad_(True)
        cls, box = fastrcnn_loss(l_t, r_t, dev(lab), dev(tg), n_old=n_old, cls_agnostic_bbox_reg=agn)
        (cls * 0.7 + box * 1.3).backward()
        ol, orr = torch.from_numpy(logits).double().requires_grad_(True), torch.from_numpy(regr).double().requires_grad_(True)
        rc, rb = oll.fastrcnn_loss(ol, orr, torch.from_numpy(lab), torch.from_numpy(tg).double(), n_old, agn)
        (rc * 0.7 + rb * 1.3).backward()
        close(cls.item(), rc.item(), 1e-5, "fastrcnn case %d cls" % i)
        close(box.item(), rb.item(), 1e-5, "fastrcnn case %d box" % i)
        close(l_t.grad.cpu().numpy(), ol.grad.numpy(), 1e-5, "fastrcnn case %d dlogits" % i)
        close(r_t.grad.cpu().numpy(), orr.grad.numpy(), 1e-5, "fastrcnn case %d dreg" % i)
    print("logit losses: %d random cases match the float64 oracle" % cases)
    # ---- RPN proposals: candidate selection (exact) and the full path
    flips = 0
    for i in range(cases):
        N, A = int(rng.integers(1, 5)), int(rng.choice([3, 6, 9, 15]))
        H, W = int(rng.integers(3, 40)), int(rng.integers(3, 50))
        sizes_a = (32, 64, 128, 256, 512)[: max(1, A // 3)]
        anchors = make_anchors(H, W, 16, sizes=sizes_a, ratios=(0.5, 1.0, 2.0)[: A // len(sizes_a)])
        logits = (rng.standard_normal((N, A, H, W)) * 2).astype(np.float32)
        if rng.random() < 0.3:
            logits = np.round(logits)  # heavy ties
        reg = (rng.standard_normal((N, 4 * A, H, W)) * 0.3).astype(np.float32)
        sizes = [(W * 16 - int(rng.integers(0, 8)), H * 16 - int(rng.integers(0, 8))) for _ in range(N)]
        pre, post = int(rng.integers(1, 3000)), int(rng.integers(1, 500))
        min_size = float(rng.choice([0, 0, 4, 16]))
        cl = bool(rng.integers(0, 2))
        lo, rg = dev(logits), dev(reg)
        if cl:
            lo, rg = lo.contiguous(memory_format=torch.channels_last), rg.contiguous(memory_format=torch.channels_last)
        cand = orpn.candidates(logits, reg, [anchors], sizes, pre, min_size)
        _, sc_dev, n_out, idx = rpn_proposals(lo, rg, dev(anchors), sizes, pre, post, 0.0, min_size, return_anchor_index=True)
        for n in range(N):
            k = int(n_out[n])
            if k != len(cand[n][2]) or not np.array_equal(idx[n, :k].cpu().numpy(), cand[n][2]):
                # the min_size filter compares decoded sides with a threshold: an ulp of expf can flip it
                if min_size > 0:
                    flips += 1
                else:
                    fail("rpn candidate selection case %d image %d" % (i, n))
        props, scores, n_out = rpn_proposals(lo, rg, dev(anchors), sizes, pre, post, 0.7, min_size)
        ref = orpn.rpn_proposals(logits, reg, [anchors], sizes, pre, post, 0.7, min_size)
        for n in range(N):
            if int(n_out[n]) != len(ref[n][0]):
                flips += 1  # an IoU within an ulp of the threshold (decode differs by ulps from torch's CPU exp)
    print("rpn: %d random cases; candidate selection exact; %d near-threshold flips in the size filter / NMS" % (cases, flips))
    # ---- ABR paste: whole batches in one launch against the numpy restatement of the reference's three methods
    import random

    from PIL import Image

    from abr_iod_b200.data.abr_paste import BoxRehearsalPaster
    from oracle import paste as opaste

    kinds_seen = {}
    for i in range(max(4, cases // 6)):
        protos = []
        for j in range(int(rng.integers(6, 40))):
            h, w = int(rng.integers(8, 320)), int(rng.integers(8, 320))
            protos.append(("%d_%05d.jpg" % (int(rng.integers(1, 16)), j), rng.integers(0, 256, (h, w, 3), dtype=np.uint8)))
        bs = int(rng.choice([2, 4, 8]))
        paster = BoxRehearsalPaster([(n, a) for n, a in protos], bs)
        st = opaste.BoxRehearsalState([(n, opaste.as_pil(a)) for n, a in protos], bs)
        images, targets = [], []
        for _ in range(int(rng.integers(1, 12))):
            h, w = int(rng.integers(60, 420)), int(rng.integers(60, 520))
            images.append(Image.fromarray(rng.integers(0, 256, (h, w, 3), dtype=np.uint8)))
            ng = int(rng.integers(1, 5))
            x1, y1 = rng.uniform(0, w * 0.6, ng), rng.uniform(0, h * 0.6, ng)
            targets.append(np.stack([x1, y1, x1 + rng.uniform(5, w * 0.4, ng), y1 + rng.uniform(5, h * 0.4, ng), rng.integers(16, 21, ng)], 1))
        seed = int(rng.integers(0, 1 << 30))
        random.seed(seed); torch.manual_seed(seed)
        ref = [opaste.transform_current_data_with_abr(st, im, t) for im, t in zip(images, targets)]
        random.seed(seed); torch.manual_seed(seed)
        outs, gts, kinds = paster.paste_batch(images, targets)
        if kinds != [r[0] for r in ref] or paster.boxes_index != st.boxes_index:
            fail("paste round %d: decisions or rehearsal index differ" % i)
        for k, (o, gt, (kind, rimg, rgt)) in enumerate(zip(outs, gts, ref)):
            kinds_seen[kind] = kinds_seen.get(kind, 0) + 1
            if not (np.array_equal(o.cpu().numpy(), rimg) and np.array_equal(gt, rgt)):
                fail("paste round %d image %d (%s): pixels or boxes differ" % (i, k, kind))
    print("paste: %d random batches bit-exact (%s)" % (max(4, cases // 6), ", ".join("%s x%d" % kv for kv in sorted(kinds_seen.items()))))
    # ---- box-head post-processing
    flips = 0
    for i in range(cases):
        n_img = int(rng.integers(1, 5))
        C = int(rng.choice([2, 5, 21, 81]))
        counts = [int(rng.integers(0, 300)) for _ in range(n_img)]
        if sum(counts) == 0:
            counts[0] = 5
        sizes = [(int(rng.integers(200, 1300)), int(rng.integers(200, 900))) for _ in range(n_img)]
        props, logits, regs = [], [], []
        for (w, h), n in zip(sizes, counts):
            c = rng.uniform([0, 0], [w, h], (n, 2))
            wh = rng.uniform(10, 200, (n, 2))
            props.append(np.clip(np.concatenate([c - wh / 2, c + wh / 2], 1), 0, [w - 1, h - 1, w - 1, h - 1]).astype(np.float32))
            logits.append((rng.standard_normal((n, C)) * 2).astype(np.float32))
            regs.append((rng.standard_normal((n, 4 * C)) * 0.5).astype(np.float32))
        lg, rg = np.concatenate(logits, 0), np.concatenate(regs, 0)
        det = int(rng.choice([0, 10, 100]))
        out = box_postprocess(dev(lg), dev(rg), dev(np.concatenate(props, 0)), counts, sizes, 0.05, 0.5, det)
        ref, _ = obp.box_postprocess(lg, rg, props, sizes, 0.05, 0.5, det)
        for n in range(n_img):
            k = out["n_host"][n]
            same = k == len(ref[n][1]) and np.array_equal(out["labels"][n, :k].cpu().numpy(), ref[n][2]) and \
                np.array_equal(out["rows"][n, :k].cpu().numpy(), ref[n][3])
            if not same:
                flips += 1  # a softmax score within an ulp of score_thresh / kthvalue, or an IoU within an ulp of 0.5
    print("box_post: %d random cases; %d images differ from the oracle by near-threshold flips" % (cases, flips))


if __name__ == "__main__":
    main()
