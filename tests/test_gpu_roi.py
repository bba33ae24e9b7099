"""GPU parity of ROIAlign / ROIPool / Pooler through the reference-shaped Python API (-> C ABI -> sm_100a kernels)
against the CPU oracle and the golden vectors produced by the reference itself.

Tolerance (north_star: "within 1e-5 relative error in fp32"): |a-b| <= 1e-5 * max|ref| + 1e-5 * |ref|.
bf16 (stated tolerance): inputs are rounded to bf16 first, then |a-b| <= 2e-2 * max|ref|."""
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


def dev(x, channels_last=False):
    t = torch.as_tensor(x).cuda()
    return t.contiguous(memory_format=torch.channels_last) if channels_last else t


@pytest.mark.parametrize("channels_last", [False, True])
def test_roi_align_forward_golden(golden, channels_last):
    from abr_iod_b200.layers import ROIAlign

    g = golden("roi_align_fwd.npz")
    x, rois = dev(g["input"], channels_last), dev(g["rois"])
    for P, ratio in [(7, 0), (7, 2), (14, 0), (3, 1)]:
        out = ROIAlign((P, P), 1 / 16, ratio)(x, rois)
        assert out.shape == (len(g["rois"]), x.shape[1], P, P) and out.dtype == torch.float32
        assert out.is_contiguous(memory_format=torch.channels_last if channels_last else torch.contiguous_format)
        close(out.cpu().numpy(), g["out_p%d_r%d" % (P, ratio)])


@pytest.mark.parametrize("channels_last", [False, True])
@pytest.mark.parametrize("C", [3, 8, 64, 260])
@pytest.mark.parametrize("P,ratio", [(7, 0), (7, 2), (14, 0), (2, 3)])
def test_roi_align_forward_backward_vs_oracle(channels_last, C, P, ratio):
    from abr_iod_b200.layers import roi_align

    rng = np.random.default_rng(C * 100 + P * 10 + ratio)
    B, H, W = 3, 25, 38
    x = rng.standard_normal((B, C, H, W)).astype(np.float32)
    rois = make_rois(rng, 50, B, W * 16, H * 16)
    xt = dev(x, channels_last).requires_grad_(True)
    out = roi_align(xt, dev(rois), (P, P), 1 / 16, ratio)
    close(out.detach().cpu().numpy(), oracle.roi_align_forward(x, rois, 1 / 16, P, P, ratio))
    gout = rng.standard_normal(out.shape).astype(np.float32)
    out.backward(dev(gout, channels_last))
    assert xt.grad.shape == xt.shape
    close(xt.grad.cpu().numpy(), oracle.roi_align_backward(gout, rois, 1 / 16, P, P, B, C, H, W, ratio))


def test_roi_align_empty_and_errors():
    from abr_iod_b200.layers import ROIAlign

    x = torch.randn(1, 4, 8, 8, device="cuda", requires_grad=True)
    out = ROIAlign((7, 7), 0.25, 2)(x, torch.zeros((0, 5), device="cuda"))
    assert out.shape == (0, 4, 7, 7)
    out.sum().backward()
    assert x.grad is not None and not x.grad.any()
    with pytest.raises(RuntimeError):
        ROIAlign((7, 7), 0.25, 2)(x, torch.zeros((3, 4), device="cuda"))


@pytest.mark.parametrize("channels_last", [False, True])
def test_roi_align_bf16_and_fp16(channels_last):
    from abr_iod_b200.layers import roi_align

    rng = np.random.default_rng(3)
    B, C, H, W, P = 2, 64, 20, 30, 7
    x = torch.from_numpy(rng.standard_normal((B, C, H, W)).astype(np.float32)).bfloat16()
    rois = make_rois(rng, 40, B, W * 16, H * 16)
    ref = oracle.roi_align_forward(x.float().numpy(), rois, 1 / 16, P, P, 0)
    xt = dev(x, channels_last).requires_grad_(True)
    out = roi_align(xt, dev(rois), (P, P), 1 / 16, 0)
    assert out.dtype == torch.bfloat16  # native bf16 path
    assert np.abs(out.detach().float().cpu().numpy() - ref).max() <= 2e-2 * np.abs(ref).max()
    gout = torch.from_numpy(rng.standard_normal(out.shape).astype(np.float32)).bfloat16()
    out.backward(dev(gout, channels_last))
    gref = oracle.roi_align_backward(gout.float().numpy(), rois, 1 / 16, P, P, B, C, H, W, 0)
    assert np.abs(xt.grad.float().cpu().numpy() - gref).max() <= 4e-2 * np.abs(gref).max()
    # fp16 is upcast like amp.float_function does (layers/roi_align.py:58): fp32 out
    out16 = roi_align(dev(x.half().float().half(), channels_last), dev(rois), (P, P), 1 / 16, 0)
    assert out16.dtype == torch.float32
    close(out16.cpu().numpy(), oracle.roi_align_forward(x.half().float().numpy(), rois, 1 / 16, P, P, 0))


@pytest.mark.parametrize("channels_last", [False, True])
def test_roi_align_full_size_adjoint_property(channels_last):
    """Config-1 shapes (B=2, C=1024, 38x63, 1024 RoIs, P=7): <fwd(x), g> == <x, bwd(g)> (the op is linear and the
    backward is its transpose), plus a slice checked against the oracle."""
    from abr_iod_b200.layers import roi_align

    rng = np.random.default_rng(8)
    B, C, H, W, P, R = 2, 1024, 38, 63, 7, 1024
    x = torch.randn(B, C, H, W, device="cuda")
    rois = make_rois(rng, R, B, 1000, 600)
    xt = (x.contiguous(memory_format=torch.channels_last) if channels_last else x.clone()).requires_grad_(True)
    out = roi_align(xt, dev(rois), (P, P), 1 / 16, 0)
    g = torch.randn_like(out)
    out.backward(g)
    lhs = (out.detach().double() * g.double()).sum().item()
    rhs = (x.double() * xt.grad.double()).sum().item()
    assert abs(lhs - rhs) <= 1e-5 * max(abs(lhs), abs(rhs), 1.0) + 1e-3
    sl = slice(100, 116)
    ref = oracle.roi_align_forward(x[:, sl].cpu().numpy(), rois, 1 / 16, P, P, 0)
    close(out.detach()[:, sl].cpu().numpy(), ref)


@pytest.mark.parametrize("channels_last", [False, True])
def test_roi_pool_vs_oracle(channels_last):
    from abr_iod_b200.layers import ROIPool
    from abr_iod_b200.layers.roi_pool import roi_pool_forward

    rng = np.random.default_rng(4)
    B, C, H, W = 2, 10, 22, 33
    x = rng.standard_normal((B, C, H, W)).astype(np.float32)
    rois = make_rois(rng, 60, B, W * 16, H * 16)
    for P in (7, 3):
        ref_out, ref_arg = oracle.roi_pool_forward(x, rois, 1 / 16, P, P)
        out, arg = roi_pool_forward(dev(x, channels_last), dev(rois), 1 / 16, P, P)
        assert arg.dtype == torch.int32
        assert np.array_equal(out.cpu().numpy(), ref_out)  # max pooling is exact
        assert np.array_equal(arg.cpu().numpy(), ref_arg)
        xt = dev(x, channels_last).requires_grad_(True)
        o = ROIPool((P, P), 1 / 16)(xt, dev(rois))
        gout = rng.standard_normal(o.shape).astype(np.float32)
        o.backward(dev(gout, channels_last))
        close(xt.grad.cpu().numpy(), oracle.roi_pool_backward(gout, ref_arg, rois, B, C, H, W))


@pytest.mark.parametrize("channels_last", [False, True])
def test_pooler_golden_single_and_multi_level(golden, channels_last):
    from abr_iod_b200.modeling.poolers import Pooler
    from abr_iod_b200.structures.bounding_box import BoxList

    g = golden("pooler.npz")
    feats = [dev(g["feat_%d" % i], channels_last) for i in range(4)]
    size = tuple(int(v) for v in g["image_size"])
    boxes = [BoxList(dev(g["boxes_%d" % b]), size, "xyxy") for b in range(2)]
    scales = tuple(g["scales"].tolist())
    for ratio in (2, 0):
        multi = Pooler((7, 7), scales, ratio)
        assert np.array_equal(multi.map_levels(boxes).cpu().numpy(), g["levels"].astype(np.int64))
        close(multi(feats, boxes).cpu().numpy(), g["multi_r%d" % ratio])
        single = Pooler((7, 7), (scales[2],), ratio)
        close(single([feats[2]], boxes).cpu().numpy(), g["single_r%d" % ratio])


def test_pooler_multilevel_backward_vs_per_level_oracle(golden):
    from abr_iod_b200.modeling.poolers import Pooler
    from abr_iod_b200.structures.bounding_box import BoxList
    from oracle import pooler as opooler

    g = golden("pooler.npz")
    size = tuple(int(v) for v in g["image_size"])
    boxes_np = [g["boxes_0"], g["boxes_1"]]
    boxes = [BoxList(dev(b), size, "xyxy") for b in boxes_np]
    scales = tuple(g["scales"].tolist())
    feats = [dev(g["feat_%d" % i]).requires_grad_(True) for i in range(4)]
    out = Pooler((7, 7), scales, 2)(feats, boxes)
    rng = np.random.default_rng(2)
    gout = rng.standard_normal(out.shape).astype(np.float32)
    out.backward(dev(gout))
    rois = opooler.to_roi_format(boxes_np)
    levels = g["levels"].astype(np.int64)
    for lvl in range(4):
        idx = np.nonzero(levels == lvl)[0]
        f = g["feat_%d" % lvl]
        ref = oracle.roi_align_backward(gout[idx], rois[idx], scales[lvl], 7, 7, *f.shape, 2)
        close(feats[lvl].grad.cpu().numpy(), ref)


@pytest.mark.parametrize("use_workspace", [True, False])
@pytest.mark.parametrize("P,ratio", [(7, 0), (14, 0), (7, 2), (14, 3)])
def test_roi_align_nhwc_every_plan_mode(monkeypatch, use_workspace, P, ratio):
    """Channels-last kernels on a large map: RoIs chosen to hit every plan mode of csrc/roi_align.cu -- EMPTY (outside
    the map), ROLLING (thick bins), THIN (bins thinner than a pixel), GENERIC (columns wider than 16 pixels, thin bins
    with P > 8) -- and the self-contained kernels that run when the caller passes no workspace."""
    from abr_iod_b200 import _lib
    from abr_iod_b200.layers import roi_align

    if not use_workspace:
        monkeypatch.setattr(_lib, "roi_align_workspace", lambda *a, **k: (None, 0))
    rng = np.random.default_rng(P * 10 + ratio)
    B, C, H, W = 2, 12, 120, 200
    x = rng.standard_normal((B, C, H, W)).astype(np.float32)
    s = 16.0
    rois = np.array([
        [0, -900, -900, -500, -500],            # EMPTY
        [1, 100, 100, 900, 700],                # ROLLING, moderately wide
        [0, 0, 0, W * s - 1, H * s - 1],        # whole map: columns ~28 px wide -> GENERIC
        [1, 320, 160, 360, 190],                # THIN: ~2.5 x 1.9 feature px
        [0, 50.5, 60.25, 51.0, 60.5],           # degenerate, forced to 1x1
        [1, 10, 10, 2000, 40],                  # very wide, very flat: GENERIC columns + thin rows
        [0, 3000, 100, 3400, 1800],             # partly off the right/bottom edge
        [1, -200, -200, 300, 250],              # partly off the top-left
    ], np.float32)
    rois = np.concatenate([rois, make_rois(rng, 24, B, int(W * s), int(H * s), adversarial=False)], 0)
    xt = dev(x, True).requires_grad_(True)
    out = roi_align(xt, dev(rois), (P, P), 1 / s, ratio)
    close(out.detach().cpu().numpy(), oracle.roi_align_forward(x, rois, 1 / s, P, P, ratio))
    gout = rng.standard_normal(out.shape).astype(np.float32)
    out.backward(dev(gout, True))
    close(xt.grad.cpu().numpy(), oracle.roi_align_backward(gout, rois, 1 / s, P, P, B, C, H, W, ratio))


@pytest.mark.parametrize("route", ["nchw_staged", "nchw_direct", "nhwc"])
@pytest.mark.parametrize("C,ratio", [(8, 2), (136, 0)])
def test_pooler_multilevel_vector_paths_vs_oracle(monkeypatch, route, C, ratio):
    """Four FPN levels with C % 4 == 0, so channels-last inputs take the 16-byte-lane kernels (one tensor map per level in
    the TMA-staged forward, the level table in the backward); C=136 has a ragged last 128-channel slice."""
    from abr_iod_b200.modeling.poolers import Pooler
    from abr_iod_b200.structures.bounding_box import BoxList
    from oracle import pooler as opooler

    from abr_iod_b200 import _lib

    channels_last = route == "nhwc"
    monkeypatch.setattr(_lib, "NCHW_STAGING", route != "nchw_direct")
    rng = np.random.default_rng(C + ratio)
    B, im_w, im_h = 2, 640, 512
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
    out = Pooler((7, 7), scales, ratio)(feats, boxes)
    ref = opooler.pooler(feats_np, boxes_np, 7, scales, ratio)
    close(out.detach().cpu().numpy(), ref)
    gout = rng.standard_normal(out.shape).astype(np.float32)
    out.backward(dev(gout, channels_last))
    rois = opooler.to_roi_format(boxes_np)
    levels = opooler.map_levels(rois[:, 1:], 2.0, 5.0)
    assert len(set(levels.tolist())) == 4
    for lvl in range(4):
        idx = np.nonzero(levels == lvl)[0]
        f = feats_np[lvl]
        gref = oracle.roi_align_backward(gout[idx], rois[idx], scales[lvl], 7, 7, *f.shape, ratio)
        close(feats[lvl].grad.cpu().numpy(), gref)


@pytest.mark.parametrize("C,P,ratio", [(8, 7, 0), (260, 7, 2), (30, 14, 0)])
def test_roi_align_nchw_direct_kernels(monkeypatch, C, P, ratio):
    """Contiguous-NCHW callers normally run through the channels-last kernels (staging copies in the workspace); without
    the staging room the C ABI falls back to its direct NCHW kernels.  Both must match the oracle."""
    from abr_iod_b200 import _lib
    from abr_iod_b200.layers import roi_align

    rng = np.random.default_rng(C + P)
    B, H, W = 2, 25, 38
    x = rng.standard_normal((B, C, H, W)).astype(np.float32)
    rois = make_rois(rng, 40, B, W * 16, H * 16)
    gout = rng.standard_normal((40, C, P, P)).astype(np.float32)
    ref = oracle.roi_align_forward(x, rois, 1 / 16, P, P, ratio)
    gref = oracle.roi_align_backward(gout, rois, 1 / 16, P, P, B, C, H, W, ratio)
    for staging in (False, True):
        monkeypatch.setattr(_lib, "NCHW_STAGING", staging)
        xt = dev(x).requires_grad_(True)
        out = roi_align(xt, dev(rois), (P, P), 1 / 16, ratio)
        assert out.is_contiguous()
        close(out.detach().cpu().numpy(), ref)
        out.backward(dev(gout))
        close(xt.grad.cpu().numpy(), gref)


def test_roi_align_nchw_staged_accumulates_when_not_zero_init():
    """zero_init=0 through the C ABI adds into the caller's gradient map in both NCHW routes."""
    from abr_iod_b200 import _lib

    rng = np.random.default_rng(77)
    B, C, H, W, P, R = 2, 16, 20, 31, 7, 30
    rois = make_rois(rng, R, B, W * 16, H * 16)
    gout = rng.standard_normal((R, C, P, P)).astype(np.float32)
    base = rng.standard_normal((B, C, H, W)).astype(np.float32)
    gref = base + oracle.roi_align_backward(gout, rois, 1 / 16, P, P, B, C, H, W, 0)
    g, r = dev(gout), dev(rois)
    for staged in (True, False):
        gin = dev(base.copy())
        ws, n = (_lib.roi_align_workspace(R, P, P, H, g.device, False, nchw_staging=(B, C, H * W, 0)) if staged else (None, 0))
        _lib.check(_lib.lib().abr_roi_align_backward(
            g.data_ptr(), r.data_ptr(), gin.data_ptr(), B, C, H, W, R, P, P, 1 / 16, 0, 0, _lib.ABR_NCHW, 0,
            ws.data_ptr() if ws is not None else None, n, 0, _lib.stream_ptr(g.device)))
        close(gin.cpu().numpy(), gref)


def test_roi_align_contiguous_maps_channels_last_pooled(monkeypatch):
    """_lib.POOLED_CHANNELS_LAST: contiguous feature maps in, channels-last RoI features out (same logical tensor), the
    gradient map contiguous again -- single level through ROIAlign and four levels through the Pooler."""
    from abr_iod_b200 import _lib
    from abr_iod_b200.layers import roi_align
    from abr_iod_b200.modeling.poolers import Pooler
    from abr_iod_b200.structures.bounding_box import BoxList
    from oracle import pooler as opooler

    monkeypatch.setattr(_lib, "POOLED_CHANNELS_LAST", True)
    rng = np.random.default_rng(12)
    B, C, H, W, P = 2, 40, 25, 38, 7
    x = rng.standard_normal((B, C, H, W)).astype(np.float32)
    rois = make_rois(rng, 50, B, W * 16, H * 16)
    gout = rng.standard_normal((50, C, P, P)).astype(np.float32)
    xt = dev(x).requires_grad_(True)
    out = roi_align(xt, dev(rois), (P, P), 1 / 16, 2)
    assert out.is_contiguous(memory_format=torch.channels_last) and not out.is_contiguous()
    close(out.detach().cpu().numpy(), oracle.roi_align_forward(x, rois, 1 / 16, P, P, 2))
    out.backward(dev(gout))  # a contiguous upstream gradient is accepted
    assert xt.grad.is_contiguous()
    close(xt.grad.cpu().numpy(), oracle.roi_align_backward(gout, rois, 1 / 16, P, P, B, C, H, W, 2))
    # four levels
    scales = (0.25, 0.125, 0.0625, 0.03125)
    im_w, im_h = 320, 256
    feats_np = [rng.standard_normal((B, 16, int(im_h * s), int(im_w * s))).astype(np.float32) for s in scales]
    boxes_np = []
    for b in range(B):
        x1, y1 = rng.uniform(0, im_w - 8, 25), rng.uniform(0, im_h - 8, 25)
        side = np.exp(rng.uniform(np.log(6), np.log(300), 25))
        boxes_np.append(np.stack([x1, y1, np.minimum(x1 + side, im_w - 1), np.minimum(y1 + side, im_h - 1)], 1).astype(np.float32))
    feats = [dev(f).requires_grad_(True) for f in feats_np]
    out = Pooler((7, 7), scales, 2)(feats, [BoxList(dev(b), (im_w, im_h), "xyxy") for b in boxes_np])
    assert out.is_contiguous(memory_format=torch.channels_last)
    close(out.detach().cpu().numpy(), opooler.pooler(feats_np, boxes_np, 7, scales, 2))
    g2 = rng.standard_normal(out.shape).astype(np.float32)
    out.backward(dev(g2))
    rois2 = opooler.to_roi_format(boxes_np)
    levels = opooler.map_levels(rois2[:, 1:], 2.0, 5.0)
    for lvl in range(4):
        idx = np.nonzero(levels == lvl)[0]
        assert feats[lvl].grad.is_contiguous()
        close(feats[lvl].grad.cpu().numpy(), oracle.roi_align_backward(g2[idx], rois2[idx], scales[lvl], 7, 7, *feats_np[lvl].shape, 2))


def test_roi_align_forward_plan_shared_between_two_maps():
    """The teacher / student pair pools the same RoIs from maps of the same shape: the second forward reuses the first's
    plans (workspace_has_plan) and must give the same result as planning afresh."""
    from abr_iod_b200.layers.roi_align import roi_align_forward

    rng = np.random.default_rng(21)
    B, C, H, W, P = 2, 32, 25, 38, 7
    a = rng.standard_normal((B, C, H, W)).astype(np.float32)
    b = rng.standard_normal((B, C, H, W)).astype(np.float32)
    rois = make_rois(rng, 60, B, W * 16, H * 16)
    for cl in (True, False):
        fa, plan = roi_align_forward(dev(a, cl), dev(rois), 1 / 16, P, P, 0, return_plan=True)
        fb = roi_align_forward(dev(b, cl), dev(rois), 1 / 16, P, P, 0, plan=plan)
        close(fa.cpu().numpy(), oracle.roi_align_forward(a, rois, 1 / 16, P, P, 0))
        close(fb.cpu().numpy(), oracle.roi_align_forward(b, rois, 1 / 16, P, P, 0))


def test_roi_align_wide_rois_on_the_staged_path():
    """RoIs whose footprint is 33..64 map pixels wide (the ring slots of the TMA-staged kernels hold rows of up to 64
    pixels) and a few wider ones (left to the generic kernel), channels-last, forward and backward against the oracle."""
    from abr_iod_b200.layers import roi_align

    rng = np.random.default_rng(64)
    B, C, H, W, P = 2, 16, 50, 76, 7
    x = rng.standard_normal((B, C, H, W)).astype(np.float32)
    n = 40
    w = np.concatenate([rng.uniform(520, 1000, n - 6), rng.uniform(1030, 1215, 6)])
    h = rng.uniform(40, 790, n)
    x1 = rng.uniform(0, 1215 - w)
    y1 = rng.uniform(0, 799 - h)
    rois = np.stack([rng.integers(0, B, n), x1, y1, x1 + w, y1 + h], 1).astype(np.float32)
    gout = rng.standard_normal((n, C, P, P)).astype(np.float32)
    for ratio in (0, 2):
        xt = dev(x, True).requires_grad_(True)
        out = roi_align(xt, dev(rois), (P, P), 1 / 16, ratio)
        close(out.detach().cpu().numpy(), oracle.roi_align_forward(x, rois, 1 / 16, P, P, ratio))
        out.backward(dev(gout, True))
        close(xt.grad.cpu().numpy(), oracle.roi_align_backward(gout, rois, 1 / 16, P, P, B, C, H, W, ratio))


def test_pooler_accepts_int_output_size_and_context_keeps_only_the_plans():
    """Pooler(7, ...) like the reference (its ROIAlign goes through _pair); and for a contiguous-NCHW call the autograd
    context keeps the per-RoI plans only, not the channels-last staging copies."""
    from abr_iod_b200 import _lib
    from abr_iod_b200.layers.roi_align import _ROIAlign
    from abr_iod_b200.modeling.poolers import Pooler
    from abr_iod_b200.structures.bounding_box import BoxList

    rng = np.random.default_rng(3)
    scales = (0.25, 0.125)
    feats = [torch.randn(1, 16, 64, 80, device="cuda", requires_grad=True), torch.randn(1, 16, 32, 40, device="cuda", requires_grad=True)]
    x1, y1 = rng.uniform(0, 200, 20), rng.uniform(0, 150, 20)
    b = np.stack([x1, y1, x1 + rng.uniform(8, 110, 20), y1 + rng.uniform(8, 100, 20)], 1).astype(np.float32)
    boxes = [BoxList(torch.from_numpy(b).cuda(), (320, 256), "xyxy")]
    a = Pooler(7, scales, 2)(feats, boxes)
    c = Pooler((7, 7), scales, 2)(feats, boxes)
    assert torch.equal(a, c)
    a.sum().backward()
    assert all(f.grad is not None for f in feats)

    class Ctx:  # a stand-in for the autograd context: what does forward keep?
        def save_for_backward(self, *t):
            self.saved = t

    x = torch.randn(2, 256, 50, 76, device="cuda")
    rois = torch.from_numpy(make_rois(rng, 300, 2, 76 * 16, 50 * 16)).cuda()
    ctx = Ctx()
    out = _ROIAlign.forward(ctx, x, rois, (7, 7), 1 / 16, 0)
    plan_bytes = int(_lib.lib().abr_roi_align_workspace_bytes(300, 7, 7, 50))
    assert ctx.plan.numel() == plan_bytes < out.numel() * 4  # no map / pooled staging copies kept alive
    ctx.saved_tensors = ctx.saved
    g = torch.randn_like(out)
    gin = _ROIAlign.backward(ctx, g)[0]
    ref = oracle.roi_align_backward(g.cpu().numpy(), rois.cpu().numpy(), 1 / 16, 7, 7, 2, 256, 50, 76, 0)
    close(gin.cpu().numpy(), ref)
