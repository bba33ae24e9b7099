"""GPU parity of the two fused logit-level losses (inclusive RoI distillation; box-head classification + box loss)
through the reference-shaped Python API against golden vectors produced by the reference's own functions (fp32 and fp64
runs with autograd gradients).  Tolerance (north_star, fp32): |a-b| <= 1e-5 * max|ref| + 1e-5 * |ref| against the fp64
run of the reference; the reference's own fp32 run is held to the same bound, so the kernel is as close as the reference
is to itself."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def close(a, ref, rel=1e-5):
    a, ref = np.asarray(a, np.float64), np.asarray(ref, np.float64)
    scale = np.abs(ref).max() if ref.size else 1.0
    err = np.abs(a - ref)
    assert (err <= rel * scale + rel * np.abs(ref)).all(), "max err %g (scale %g)" % (err.max(), scale)


def dev(x):
    return torch.as_tensor(x).cuda()


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_roi_distillation_id_golden(golden, tag):
    from abr_iod_b200.distillation.distillation import calculate_roi_distillation_losses, roi_distillation_id_terms

    g = golden("logit_losses.npz")
    ss, sb = dev(g["id_%s_ss" % tag]), dev(g["id_%s_sb" % tag])
    ts = dev(g["id_%s_ts" % tag]).requires_grad_(True)
    tb = dev(g["id_%s_tb" % tag]).requires_grad_(True)
    loss = calculate_roi_distillation_losses((ss, sb), (ts, tb), dist="id")
    (0.5 * loss).backward()  # the call site multiplies by cfg.DIST.ALPHA (train_incremental.py:103)
    close(loss.item(), g["id_%s_loss64" % tag])
    close(g["id_%s_loss32" % tag], g["id_%s_loss64" % tag])
    close(ts.grad.cpu().numpy(), 0.5 * g["id_%s_gs64" % tag])
    close(tb.grad.cpu().numpy(), 0.5 * g["id_%s_gb64" % tag])
    _, terms = roi_distillation_id_terms((ss, sb), (ts.detach(), tb.detach()))
    t = terms.cpu().numpy()
    assert abs(t[0] - (t[1] + t[2])) <= 1e-6 * max(1.0, abs(t[0]))


@pytest.mark.parametrize("tag", ["a", "b", "c", "d"])
def test_fastrcnn_loss_golden(golden, tag):
    from abr_iod_b200.modeling.roi_heads.box_head.loss import FastRCNNLossComputation
    from abr_iod_b200.structures.bounding_box import BoxList

    g = golden("logit_losses.npz")
    n_old, agn = (int(v) for v in g["frcnn_%s_cfg" % tag])
    labels, targets = g["frcnn_%s_labels" % tag], g["frcnn_%s_targets" % tag]
    R = len(labels)
    ev = FastRCNNLossComputation(None, None, None, bool(agn), "id" if n_old >= 0 else None, old_classes=list(range(max(n_old, 0))))
    props = []
    for a, b in ((0, R // 2), (R // 2, R)):
        bl = BoxList(torch.zeros((b - a, 4)).cuda(), (100, 100), "xyxy")
        bl.add_field("labels", dev(labels[a:b]))
        bl.add_field("regression_targets", dev(targets[a:b]))
        props.append(bl)
    ev._proposals = props
    logits = dev(g["frcnn_%s_logits" % tag]).requires_grad_(True)
    reg = dev(g["frcnn_%s_reg" % tag]).requires_grad_(True)
    cls, box = ev([logits], [reg])
    (2.0 * cls + 3.0 * box).backward()  # distinct upstream gradients, as in the golden run
    close(cls.item(), g["frcnn_%s_cls64" % tag])
    close(box.item(), g["frcnn_%s_box64" % tag])
    close(logits.grad.cpu().numpy(), g["frcnn_%s_gl64" % tag])
    close(reg.grad.cpu().numpy(), g["frcnn_%s_gr64" % tag])
    close(g["frcnn_%s_gl32" % tag], g["frcnn_%s_gl64" % tag])


def test_logit_losses_vs_oracle_larger_and_edge_cases():
    from abr_iod_b200.distillation.distillation import calculate_roi_distillation_losses
    from abr_iod_b200.modeling.roi_heads.box_head.loss import fastrcnn_loss
    from oracle import logit_losses as oll

    rng = np.random.default_rng(17)
    # 512 RoIs, COCO-sized heads (41 old + 40 new classes), large logits (softmax saturation)
    R, Co, Ct = 512, 41, 81
    ss = (rng.standard_normal((R, Co)) * 8).astype(np.float32)
    sb = rng.standard_normal((R, Co, 4)).astype(np.float32)
    ts = (rng.standard_normal((R, Ct)) * 8).astype(np.float32)
    tb = rng.standard_normal((R, Ct, 4)).astype(np.float32)
    t_s, t_b = dev(ts).requires_grad_(True), dev(tb).requires_grad_(True)
    loss = calculate_roi_distillation_losses((dev(ss), dev(sb)), (t_s, t_b), dist="id")
    loss.backward()
    o_s = torch.from_numpy(ts).double().requires_grad_(True)
    o_b = torch.from_numpy(tb).double().requires_grad_(True)
    ref, _, _ = oll.roi_distillation_id(torch.from_numpy(ss).double(), torch.from_numpy(sb).double(), o_s, o_b)
    ref.backward()
    close(loss.item(), ref.item())
    close(t_s.grad.cpu().numpy(), o_s.grad.numpy())
    close(t_b.grad.cpu().numpy(), o_b.grad.numpy())
    # box-head loss: ignored rows (-100), no positives at all, teacher-student identical heads
    C = 21
    logits = (rng.standard_normal((R, C)) * 4).astype(np.float32)
    reg = rng.standard_normal((R, 4 * C)).astype(np.float32)
    targets = rng.standard_normal((R, 4)).astype(np.float32) * 2
    for labels in (np.where(rng.random(R) < 0.2, -100, rng.integers(16, C, R)).astype(np.int64), np.zeros(R, np.int64)):
        l_t, r_t = dev(logits).requires_grad_(True), dev(reg).requires_grad_(True)
        cls, box = fastrcnn_loss(l_t, r_t, dev(labels), dev(targets), n_old=15)
        (cls + box).backward()
        ol = torch.from_numpy(logits).double().requires_grad_(True)
        orr = torch.from_numpy(reg).double().requires_grad_(True)
        rc, rb = oll.fastrcnn_loss(ol, orr, torch.from_numpy(labels), torch.from_numpy(targets).double(), 15)
        (rc + rb).backward()
        close(cls.item(), rc.item())
        close(box.item(), rb.item())
        close(l_t.grad.cpu().numpy(), ol.grad.numpy())
        close(r_t.grad.cpu().numpy(), orr.grad.numpy())
    with pytest.raises(NotImplementedError):
        calculate_roi_distillation_losses((dev(ss), dev(sb)), (dev(ts), dev(tb)), dist="l2")
    with pytest.raises(RuntimeError):
        calculate_roi_distillation_losses((dev(ss), dev(sb)), (dev(ss), dev(sb)), dist="id")  # student must know more classes
