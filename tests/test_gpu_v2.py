"""GPU parity of the gather-form ("v2") ROIAlign kernels (csrc/roi_v2.cu) and of the fused ARD step
(``abr_roi_ard_fused``: teacher pooling + student pooling + ARD loss + backward into the student's map) against the CPU
oracle, through the Python API -> C ABI.

Tolerances, fp32 (north_star: 1e-5 relative):
* ``close``      |a-b| <= 1e-5*max|ref| + 1e-5*|ref|                       (the round-1 criterion)
* ``close_cond`` |a-b| <= 1e-5 * cond, ELEMENT-WISE, where cond is the sum of the absolute values of the terms that make
  up that element (ROIAlign's weights are non-negative, so cond = ROIAlign(|x|), resp. ROIAlign_backward(|g|)): the
  standard backward-error form -- an element that is small because large terms cancel cannot be resolved better than this
  by ANY fp32 summation order, the reference's included.
"""
import numpy as np
import pytest
import torch

import oracle
from inputs import make_rois

pytestmark = pytest.mark.gpu


def close(a, ref, rel=1e-5):
    a, ref = np.asarray(a, np.float64), np.asarray(ref, np.float64)
    scale = np.abs(ref).max() if ref.size else 1.0
    err = np.abs(a - ref)
    ok = err <= rel * scale + rel * np.abs(ref)
    assert ok.all(), "max err %g (scale %g) at %s" % (err.max(), scale, np.unravel_index(err.argmax(), err.shape))


def close_cond(a, ref, cond, rel=1e-5):
    a, ref, cond = np.asarray(a, np.float64), np.asarray(ref, np.float64), np.asarray(cond, np.float64)
    err = np.abs(a - ref)
    ok = err <= rel * cond + 1e-30
    assert ok.all(), "element-wise: err %g vs cond %g at %s" % (
        err[~ok].max(), cond[~ok].min(), np.unravel_index(np.argmax(err - rel * cond), err.shape))


def dev(x, channels_last=False):
    t = torch.as_tensor(x).cuda()
    return t.contiguous(memory_format=torch.channels_last) if channels_last else t


@pytest.fixture
def force_v2():
    from abr_iod_b200 import _lib

    _lib.set_option("roi_v2", 1)
    yield
    _lib.set_option("roi_v2", -1)


@pytest.mark.parametrize("channels_last", [True, False])
@pytest.mark.parametrize("C", [3, 8, 64, 260])
@pytest.mark.parametrize("P,ratio", [(7, 0), (7, 2), (14, 0), (2, 3), (16, 0), (1, 1)])
def test_v2_forward_backward_vs_oracle(force_v2, channels_last, C, P, ratio):
    from abr_iod_b200.layers import roi_align

    rng = np.random.default_rng(C * 100 + P * 10 + ratio)
    B, H, W = 3, 25, 38
    x = rng.standard_normal((B, C, H, W)).astype(np.float32)
    rois = make_rois(rng, 50, B, W * 16, H * 16)
    xt = dev(x, channels_last).requires_grad_(True)
    out = roi_align(xt, dev(rois), (P, P), 1 / 16, ratio)
    ref = oracle.roi_align_forward(x, rois, 1 / 16, P, P, ratio)
    close(out.detach().cpu().numpy(), ref)
    close_cond(out.detach().cpu().numpy(), ref, oracle.roi_align_forward(np.abs(x), rois, 1 / 16, P, P, ratio))
    gout = rng.standard_normal(out.shape).astype(np.float32)
    out.backward(dev(gout, channels_last))
    gref = oracle.roi_align_backward(gout, rois, 1 / 16, P, P, B, C, H, W, ratio)
    close(xt.grad.cpu().numpy(), gref)
    close_cond(xt.grad.cpu().numpy(), gref, oracle.roi_align_backward(np.abs(gout), rois, 1 / 16, P, P, B, C, H, W, ratio))


@pytest.mark.parametrize("P,ratio", [(7, 0), (14, 0), (7, 2), (14, 3)])
def test_v2_every_plan_mode(force_v2, P, ratio):
    """EMPTY, thin, fat, border-hugging RoIs and RoIs that overflow the records (bins wider than 15 map pixels, footprints
    wider than 64: the per-sample path inside the same kernels)."""
    from abr_iod_b200.layers import roi_align

    rng = np.random.default_rng(P * 10 + ratio)
    B, C, H, W = 2, 12, 120, 200
    x = rng.standard_normal((B, C, H, W)).astype(np.float32)
    s = 16.0
    rois = np.array([
        [0, -900, -900, -500, -500], [1, 100, 100, 900, 700], [0, 0, 0, W * s - 1, H * s - 1], [1, 320, 160, 360, 190],
        [0, 50.5, 60.25, 51.0, 60.5], [1, 10, 10, 2000, 40], [0, 3000, 100, 3400, 1800], [1, -200, -200, 300, 250],
    ], np.float32)
    rois = np.concatenate([rois, make_rois(rng, 24, B, int(W * s), int(H * s), adversarial=False)], 0)
    xt = dev(x, True).requires_grad_(True)
    out = roi_align(xt, dev(rois), (P, P), 1 / s, ratio)
    close(out.detach().cpu().numpy(), oracle.roi_align_forward(x, rois, 1 / s, P, P, ratio))
    gout = rng.standard_normal(out.shape).astype(np.float32)
    out.backward(dev(gout, True))
    close(xt.grad.cpu().numpy(), oracle.roi_align_backward(gout, rois, 1 / s, P, P, B, C, H, W, ratio))


def test_v2_bf16(force_v2):
    from abr_iod_b200.layers import roi_align

    rng = np.random.default_rng(3)
    B, C, H, W, P = 2, 64, 20, 30, 14
    x = torch.from_numpy(rng.standard_normal((B, C, H, W)).astype(np.float32)).bfloat16()
    rois = make_rois(rng, 40, B, W * 16, H * 16)
    ref = oracle.roi_align_forward(x.float().numpy(), rois, 1 / 16, P, P, 0)
    xt = dev(x, True).requires_grad_(True)
    out = roi_align(xt, dev(rois), (P, P), 1 / 16, 0)
    assert out.dtype == torch.bfloat16
    assert np.abs(out.detach().float().cpu().numpy() - ref).max() <= 2e-2 * np.abs(ref).max()  # stated bf16 tolerance
    gout = torch.from_numpy(rng.standard_normal(out.shape).astype(np.float32)).bfloat16()
    out.backward(dev(gout, True))
    gref = oracle.roi_align_backward(gout.float().numpy(), rois, 1 / 16, P, P, B, C, H, W, 0)
    assert np.abs(xt.grad.float().cpu().numpy() - gref).max() <= 4e-2 * np.abs(gref).max()


@pytest.mark.parametrize("channels_last", [True, False])
def test_v2_pooler_multilevel(force_v2, channels_last):
    from abr_iod_b200.modeling.poolers import Pooler
    from abr_iod_b200.structures.bounding_box import BoxList
    from oracle import pooler as opooler

    rng = np.random.default_rng(17)
    B, C, im_w, im_h = 2, 136, 640, 512
    scales = (0.25, 0.125, 0.0625, 0.03125)
    feats_np = [rng.standard_normal((B, C, int(im_h * s), int(im_w * s))).astype(np.float32) for s in scales]
    boxes_np = []
    for b in range(B):
        n = 30
        x1, y1 = rng.uniform(0, im_w - 8, n), rng.uniform(0, im_h - 8, n)
        side = np.exp(rng.uniform(np.log(6), np.log(700), n))
        bx = np.stack([x1, y1, np.minimum(x1 + side, im_w - 1), np.minimum(y1 + side * rng.uniform(0.5, 2, n), im_h - 1)], 1)
        big = np.array([[0, 0, im_w - 1, im_h - 1], [60, 40, 600, 500], [20, 30, 420, 390]], np.float32)
        boxes_np.append(np.concatenate([bx.astype(np.float32), big], 0))
    feats = [dev(f, channels_last).requires_grad_(True) for f in feats_np]
    boxes = [BoxList(dev(b), (im_w, im_h), "xyxy") for b in boxes_np]
    for ratio in (2, 0):
        for f in feats:
            f.grad = None
        out = Pooler((7, 7), scales, ratio)(feats, boxes)
        close(out.detach().cpu().numpy(), opooler.pooler(feats_np, boxes_np, 7, scales, ratio))
        gout = rng.standard_normal(out.shape).astype(np.float32)
        out.backward(dev(gout, channels_last))
        rois = opooler.to_roi_format(boxes_np)
        levels = opooler.map_levels(rois[:, 1:], 2.0, 5.0)
        for lvl in range(4):
            idx = np.nonzero(levels == lvl)[0]
            gref = oracle.roi_align_backward(gout[idx], rois[idx], scales[lvl], 7, 7, *feats_np[lvl].shape, ratio)
            close(feats[lvl].grad.cpu().numpy(), gref)


def bench_like_rois(rng, R, B, im_w, im_h):
    """bench.py's RoI distribution: centres uniform, sides U(16,400) px, clipped, 5 % degenerate."""
    cx, cy = rng.uniform(0, im_w, R), rng.uniform(0, im_h, R)
    bw, bh = rng.uniform(16, 400, R), rng.uniform(16, 400, R)
    deg = rng.uniform(0, 1, R) < 0.05
    bw[deg] = rng.uniform(0, 1, deg.sum())
    x1, x2 = np.clip(cx - bw / 2, 0, im_w - 1), np.clip(cx + bw / 2, 0, im_w - 1)
    y1, y2 = np.clip(cy - bh / 2, 0, im_h - 1), np.clip(cy + bh / 2, 0, im_h - 1)
    img = np.repeat(np.arange(B), -(-R // B))[:R]
    return np.stack([img, x1, y1, x2, y2], 1).astype(np.float32)


@pytest.mark.parametrize("P", [14, 7])
def test_full_size_config1_parity(P):
    """BASELINE.json configs[0] at FULL size -- [2,1024,38,63] fp32, 1024 RoIs (512 / image), sampling_ratio 0, P = 14 (and
    the shipped P = 7) -- forward and backward on the default kernel route, channels-last:
    a 64-channel slice of every RoI against the oracle (element-wise criterion), and adjointness over the whole tensor."""
    from abr_iod_b200.layers import roi_align

    rng = np.random.default_rng(P)
    B, C, H, W, R = 2, 1024, 38, 63, 1024
    x = rng.standard_normal((B, C, H, W)).astype(np.float32)
    rois = bench_like_rois(rng, R, B, 1000, 600)
    xt = dev(x, True).requires_grad_(True)
    out = roi_align(xt, dev(rois), (P, P), 1 / 16, 0)
    sl = slice(480, 544)  # spans two 128-channel slices
    ref = oracle.roi_align_forward(x[:, sl], rois, 1 / 16, P, P, 0)
    got = out.detach()[:, sl].cpu().numpy()
    close(got, ref)
    close_cond(got, ref, oracle.roi_align_forward(np.abs(x[:, sl]), rois, 1 / 16, P, P, 0))
    g = torch.randn(out.shape, device="cuda").contiguous(memory_format=torch.channels_last)
    out.backward(g)
    gs = g[:, sl].cpu().numpy()
    gref = oracle.roi_align_backward(gs, rois, 1 / 16, P, P, B, sl.stop - sl.start, H, W, 0)
    ggot = xt.grad[:, sl].cpu().numpy()
    close(ggot, gref)
    close_cond(ggot, gref, oracle.roi_align_backward(np.abs(gs), rois, 1 / 16, P, P, B, sl.stop - sl.start, H, W, 0), 2e-5)
    lhs = (out.detach().double() * g.double()).sum().item()
    rhs = (torch.as_tensor(x).cuda().double() * xt.grad.double()).sum().item()
    assert abs(lhs - rhs) <= 1e-5 * max(abs(lhs), abs(rhs), 1.0) + 1e-3


# ------------------------------------------------------------------------------------------------ fused ARD step
def fused(t_np, s_np, rois, P, ratio, gamma, channels_last=True, beta=None, head_grad=None):
    from abr_iod_b200.distillation.distillation import pooled_attentive_roi_distillation as fused_op

    t = dev(t_np, channels_last)
    s = dev(s_np, channels_last).requires_grad_(True)
    f_old, f_new, loss = fused_op(t, s, dev(rois), (P, P), 1 / 16, ratio, gamma)
    total = loss if beta is None else loss * beta
    if head_grad is not None:
        total = total + (f_new * dev(head_grad)).sum()
    total.backward()
    return f_old.detach(), f_new.detach(), loss.item(), s.grad


def oracle_chain(t, s, rois, P, ratio, gamma):
    B, C, H, W = s.shape
    ro, rn = oracle.roi_align_forward(t, rois, 1 / 16, P, P, ratio), oracle.roi_align_forward(s, rois, 1 / 16, P, P, ratio)
    loss, afd, pad, dfn = oracle.ard(ro, rn, gamma)
    return ro, rn, loss, dfn, oracle.roi_align_backward(dfn, rois, 1 / 16, P, P, B, C, H, W, ratio)


@pytest.mark.parametrize("channels_last", [True, False])
@pytest.mark.parametrize("C,P,ratio", [(256, 7, 0), (130, 14, 0), (64, 7, 2), (1024, 14, 0), (36, 3, 1), (6, 7, 0)])
def test_fused_vs_oracle_chain(channels_last, C, P, ratio):
    rng = np.random.default_rng(C + P)
    B, H, W = 2, 25, 38
    t = rng.standard_normal((B, C, H, W)).astype(np.float32)
    s = (t + 0.2 * rng.standard_normal(t.shape)).astype(np.float32)
    rois = make_rois(rng, 40, B, W * 16, H * 16)
    f_old, f_new, loss, grad = fused(t, s, rois, P, ratio, 0.7, channels_last)
    ro, rn, rl, _, rg = oracle_chain(t, s, rois, P, ratio, 0.7)
    assert f_old.shape == (40, C, P, P) and f_old.is_contiguous(memory_format=torch.channels_last)
    close(f_old.cpu().numpy(), ro)
    close(f_new.cpu().numpy(), rn)
    assert abs(loss - rl) <= 1e-5 * abs(rl), (loss, rl)
    close(grad.cpu().numpy(), rg, 2e-5)


def test_fused_upstream_scale_head_gradient_and_no_grad():
    from abr_iod_b200.distillation.distillation import pooled_attentive_roi_distillation as fused_op

    rng = np.random.default_rng(5)
    B, C, H, W, P = 2, 64, 20, 30, 7
    t = rng.standard_normal((B, C, H, W)).astype(np.float32)
    s = (t + 0.3 * rng.standard_normal(t.shape)).astype(np.float32)
    rois = make_rois(rng, 30, B, W * 16, H * 16)
    _, _, _, _, rg = oracle_chain(t, s, rois, P, 0, 1.0)
    # upstream gradient != 1 (cfg.DIST.BETA, amp loss scale) plus a gradient arriving from the box head at the pooled features
    hg = rng.standard_normal((30, C, P, P)).astype(np.float32)
    _, _, _, grad = fused(t, s, rois, P, 0, 1.0, beta=0.375, head_grad=hg)
    extra = oracle.roi_align_backward(hg, rois, 1 / 16, P, P, B, C, H, W, 0)
    close(grad.cpu().numpy(), 0.375 * rg + extra, 2e-5)
    # identical maps: loss exactly 0 and (sign(0) = 0) a zero gradient
    _, _, loss, grad = fused(t, t.copy(), rois, P, 0, 1.0)
    assert loss == 0.0 and not grad.any()
    # no gradient requested: plain tensors, no backward kernel
    f_old, f_new, loss_t = fused_op(dev(t, True), dev(s, True), dev(rois), (P, P), 1 / 16, 0, 1.0)
    assert not loss_t.requires_grad and not f_new.requires_grad
    # the teacher carries no gradient in the reference; asking for one is an error
    tt = dev(t, True).requires_grad_(True)
    ss = dev(s, True).requires_grad_(True)
    with pytest.raises(RuntimeError):
        fused_op(tt, ss, dev(rois), (P, P), 1 / 16, 0, 1.0)[2].backward()
    # a second backward over the same graph gives the same gradient (nothing is consumed in place)
    ss = dev(s, True).requires_grad_(True)
    loss_t = fused_op(dev(t, True), ss, dev(rois), (P, P), 1 / 16, 0, 1.0)[2]
    loss_t.backward(retain_graph=True)
    g1 = ss.grad.clone()
    ss.grad = None
    loss_t.backward()
    assert torch.equal(g1, ss.grad)


def ard_sign_allowance(f_old, f_new, sl, gamma, thr=3e-5):
    """The pad term of ARD has sign(A_new - A_old) in its gradient (distillation.py:115-116: L1 loss).  Where that
    difference is below what fp32 resolves (|d| < thr * A_old; the attention values carry ~1e-6..1e-5 relative error in
    ANY fp32 evaluation, the reference's included), either sign is a correct fp32 answer.  Returns, for the channel slice
    ``sl``, the exact bound [R,c,P,P] on how much dL/df_new changes when those signs flip:
        |delta dF[c,j]| <= (4*gamma*|F[c,j]| / (C*N)) * s_j * ([j ambiguous] + sum_{i ambiguous} s_i)
    (from dF = (2F/C)*HW*s_j*(g_j - sum_k g_k s_k), g = gamma*sign(d)/(N*HW)).  float64 torch is the checker here."""
    N, C, PH, PW = f_new.shape
    HW = PH * PW
    mo = (f_old.double() ** 2).mean(1).flatten(1)
    mn = (f_new.double() ** 2).mean(1).flatten(1)
    a_old = HW * torch.softmax(mo, 1)
    s_new = torch.softmax(mn, 1)
    amb = ((HW * s_new - a_old).abs() < thr * a_old).double()
    factor = s_new * (amb + (amb * s_new).sum(1, keepdim=True))  # [N, HW]
    allow = (4.0 * gamma / (C * N)) * f_new[:, sl].double().abs() * factor.view(N, 1, PH, PW)
    return allow.float().cpu().numpy(), int(amb.sum().item())


def close_allow(a, ref, allow, rel):
    a, ref = np.asarray(a, np.float64), np.asarray(ref, np.float64)
    scale = np.abs(ref).max()
    err = np.abs(a - ref)
    ok = err <= rel * scale + rel * np.abs(ref) + 1.01 * np.asarray(allow, np.float64)
    assert ok.all(), "max err %g (scale %g) at %s" % (err[~ok].max(), scale, np.unravel_index(np.argmax(err * ~ok), err.shape))


@pytest.mark.parametrize("P", [14, 7])
def test_fused_full_size_config1(P):
    """configs[0] at full size through the fused call; compared with (a) the separate ops of this library over all
    elements and (b) the oracle chain on the loss and on a slice of the map gradient."""
    from abr_iod_b200.distillation.distillation import calculate_attentive_roi_feature_distillation as ard
    from abr_iod_b200.layers import roi_align

    rng = np.random.default_rng(100 + P)
    B, C, H, W, R = 2, 1024, 38, 63, 1024
    t = rng.standard_normal((B, C, H, W)).astype(np.float32)
    s = (t + np.float32(0.1) * rng.standard_normal(t.shape).astype(np.float32)).astype(np.float32)
    rois = bench_like_rois(rng, R, B, 1000, 600)
    f_old, f_new, loss, grad = fused(t, s, rois, P, 0, 1.0)
    st = dev(s, True).requires_grad_(True)
    u_old = roi_align(dev(t, True), dev(rois), (P, P), 1 / 16, 0)
    u_new = roi_align(st, dev(rois), (P, P), 1 / 16, 0)
    u_loss = ard(u_old, u_new, 1.0)
    u_loss.backward()
    close(f_old.cpu().numpy(), u_old.detach().cpu().numpy(), 2e-6)
    close(f_new.cpu().numpy(), u_new.detach().cpu().numpy(), 2e-6)
    assert abs(loss - u_loss.item()) <= 1e-5 * abs(loss)
    # oracle: the loss over everything (pooled tensors from the GPU, themselves checked above and in test_full_size_*),
    # the gradient on a 64-channel slice of the map
    rl, _, _, dfn = oracle.ard(u_old.detach().cpu().numpy(), u_new.detach().cpu().numpy(), 1.0)
    assert abs(loss - rl) <= 1e-5 * abs(rl), (loss, rl)
    sl = slice(96, 160)
    rg = oracle.roi_align_backward(np.ascontiguousarray(dfn[:, sl]), rois, 1 / 16, P, P, B, sl.stop - sl.start, H, W, 0)
    # positions whose sign(A_new - A_old) fp32 cannot resolve: their exact effect on the map gradient is allowed on top
    allow, n_amb = ard_sign_allowance(u_old.detach(), u_new.detach(), sl, 1.0)
    assert n_amb < 0.03 * R * P * P
    allow_map = oracle.roi_align_backward(allow, rois, 1 / 16, P, P, B, sl.stop - sl.start, H, W, 0)
    close_allow(grad[:, sl].cpu().numpy(), rg, allow_map, 2e-5)
    close_allow(st.grad[:, sl].cpu().numpy(), rg, allow_map, 2e-5)


def test_stage_timing_samples_every_nth_fused_call():
    """abr_stage_timing_begin_every: events around the four stages of every n-th abr_roi_ard_fused call on the caller's stream."""
    from abr_iod_b200 import _lib
    from abr_iod_b200.distillation.distillation import pooled_attentive_roi_distillation as fused_op

    rng = np.random.default_rng(2)
    B, C, H, W, P = 1, 32, 12, 16, 7
    t = rng.standard_normal((B, C, H, W)).astype(np.float32)
    rois = make_rois(rng, 12, B, W * 16, H * 16)
    tt, rr = dev(t, True), dev(rois)
    for every, calls, want in ((1, 3, 3), (4, 9, 3), (2, 5, 3)):
        _lib.stage_timing_begin(16, every)
        for _ in range(calls):
            ss = dev(t + 0.1, True).requires_grad_(True)
            fused_op(tt, ss, rr, (P, P), 1 / 16, 0, 1.0)
        n, stages = _lib.stage_timing_end()
        assert n == want, (every, calls, n)
        assert set(stages) == set(_lib.FUSED_STAGES) and all(v > 0 for v in stages.values())
    # outside begin/end nothing is recorded
    fused_op(tt, dev(t + 0.1, True).requires_grad_(True), rr, (P, P), 1 / 16, 0, 1.0)
    _lib.stage_timing_begin(4)
    assert _lib.stage_timing_end()[0] == 0
