"""CPU check of the v2 (gather-form) ROIAlign device logic: ``abr_iod_b200/csrc/roi_v2.cuh`` -- the same source the
sm_100a kernels are built from -- is compiled for the host by ``tools/emu`` and run lane by lane in the kernels' task
decomposition, then compared with the CPU oracle (``oracle/abr_oracle.c``).  Catches plan / cache / edge-case logic
errors before any GPU time is spent.  Tolerance as in the GPU parity tests: |a-b| <= 1e-5*max|ref| + 1e-5*|ref|."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import oracle
from inputs import make_rois

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_DIR = os.path.join(ROOT, "tools", "emu")
_f32p = ctypes.POINTER(ctypes.c_float)
_i32p = ctypes.POINTER(ctypes.c_int32)
_int = ctypes.c_int


@pytest.fixture(scope="module")
def emu():
    out = os.path.join(EMU_DIR, "_build", "libroi_v2_emu.so")
    srcs = [os.path.join(EMU_DIR, "roi_v2_emu.cpp"), os.path.join(EMU_DIR, "emu_shim.h"),
            os.path.join(ROOT, "abr_iod_b200", "csrc", "roi_v2.cuh")]
    if not os.path.exists(out) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in srcs):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-w", "-o", out, srcs[0]])
    L = ctypes.CDLL(out)
    L.emu_plan_words.argtypes = [_int, _int]
    L.emu_plan.argtypes = [_f32p, _int, _int, _int, ctypes.c_float, _int, _int, _int, _i32p, _int]
    L.emu_fwd.argtypes = [_i32p, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p] + [_int] * 6 + [ctypes.c_float, _int, _int]
    L.emu_bwd.argtypes = [_i32p, _f32p, _f32p, _f32p, _f32p, _f32p] + [_int] * 6 + [ctypes.c_float, _int, _int, _int]
    return L


def _p(a, t=_f32p):
    return a.ctypes.data_as(t) if a is not None else None


def close(a, ref, rel=1e-5):
    a, ref = np.asarray(a, np.float64), np.asarray(ref, np.float64)
    scale = np.abs(ref).max() if ref.size else 1.0
    err = np.abs(a - ref)
    ok = err <= rel * scale + rel * np.abs(ref)
    assert ok.all(), "max err %g (scale %g) at %s" % (err.max(), scale, np.unravel_index(err.argmax(), err.shape))


def nhwc(x):
    return np.ascontiguousarray(np.transpose(x, (0, 2, 3, 1)))


def nchw(x):
    return np.ascontiguousarray(np.transpose(x, (0, 3, 1, 2)))


def plans_for(emu, rois, H, W, scale, P, ratio, nth=1):
    words = emu.emu_plan_words(P, P)
    plans = np.zeros((len(rois), words), np.int32)
    emu.emu_plan(_p(rois), len(rois), H, W, scale, P, P, ratio, _p(plans, _i32p), nth)
    return plans


CASES = [(7, 0), (7, 2), (14, 0), (14, 2), (2, 3), (3, 1), (16, 0), (1, 0)]


@pytest.mark.parametrize("P,ratio", CASES)
@pytest.mark.parametrize("C,V", [(8, 4), (36, 4), (5, 1)])
def test_forward_backward(emu, P, ratio, C, V):
    rng = np.random.default_rng(P * 100 + ratio * 10 + C)
    B, H, W = 2, 25, 38
    x = rng.standard_normal((B, C, H, W)).astype(np.float32)
    rois = make_rois(rng, 60, B, W * 16, H * 16)
    # thin RoIs (several bins inside one map pixel) and RoIs hugging every border
    extra = np.array([[0, 100, 100, 103, 230], [1, 5, 5, 300, 9], [0, 0, 0, 15, 15], [1, W * 16 - 20, H * 16 - 20, W * 16 + 5, H * 16 + 5],
                      [0, -40, 50, 30, 120], [1, 200, -30, 330, 12]], np.float32)
    rois = np.concatenate([rois, extra]).astype(np.float32)
    R = len(rois)
    plans = plans_for(emu, rois, H, W, 1 / 16, P, ratio, nth=1 + (P % 3))
    out = np.full((R, P, P, C), np.nan, np.float32)
    emu.emu_fwd(_p(plans, _i32p), _p(rois), _p(nhwc(x)), None, _p(out), None, None, R, C, H, W, P, P, 1 / 16, ratio, V)
    ref = oracle.roi_align_forward(x, rois, 1 / 16, P, P, ratio)
    close(nchw(out), ref)
    gout = rng.standard_normal(ref.shape).astype(np.float32)
    gmap = np.zeros((B, H, W, C), np.float32)
    emu.emu_bwd(_p(plans, _i32p), _p(rois), _p(gmap), _p(nhwc(gout)), None, None, R, C, H, W, P, P, 1 / 16, ratio, V, 0)
    close(nchw(gmap), oracle.roi_align_backward(gout, rois, 1 / 16, P, P, B, C, H, W, ratio))


def test_generic_and_wide(emu):
    """RoIs that overflow the records (bins wider than 15 map pixels, footprints wider than 64) take the per-sample path."""
    rng = np.random.default_rng(5)
    B, C, H, W, P = 1, 8, 60, 90, 2
    x = rng.standard_normal((B, C, H, W)).astype(np.float32)
    rois = np.array([[0, 0, 0, 89, 59], [0, 3, 4, 80, 20], [0, 10, 10, 30, 55], [0, 5, 5, 9, 9]], np.float32)
    plans = plans_for(emu, rois, H, W, 1.0, P, 0)
    assert plans[0, 0] == 3 and plans[3, 0] == 1
    out = np.full((4, P, P, C), np.nan, np.float32)
    emu.emu_fwd(_p(plans, _i32p), _p(rois), _p(nhwc(x)), None, _p(out), None, None, 4, C, H, W, P, P, 1.0, 0, 4)
    ref = oracle.roi_align_forward(x, rois, 1.0, P, P, 0)
    close(nchw(out), ref)
    gout = rng.standard_normal(ref.shape).astype(np.float32)
    gmap = np.zeros((B, H, W, C), np.float32)
    emu.emu_bwd(_p(plans, _i32p), _p(rois), _p(gmap), _p(nhwc(gout)), None, None, 4, C, H, W, P, P, 1.0, 0, 4, 0)
    close(nchw(gmap), oracle.roi_align_backward(gout, rois, 1.0, P, P, B, C, H, W, 0))


def ard_coefficients(sums, N, C, HW, gamma):
    """Per-position coefficients of dL/df_new = ka*(f_new - f_old) + kb*f_new from the channel sums (the closed form of
    oracle.ard / distillation.py:86-130), in float64."""
    so, sn = sums[..., 0].astype(np.float64) / C, sums[..., 1].astype(np.float64) / C
    eo, en = np.exp(so - so.max(1, keepdims=True)), np.exp(sn - sn.max(1, keepdims=True))
    a_old = HW * eo / eo.sum(1, keepdims=True)
    s = en / en.sum(1, keepdims=True)
    d = HW * s - a_old
    g = gamma * np.sign(d) / (N * HW)
    gs = (g * s).sum(1, keepdims=True)
    ka = 2.0 * a_old / (N * C * HW)
    kb = (2.0 / C) * HW * s * (g - gs)
    return np.stack([ka, kb], -1).astype(np.float32)


@pytest.mark.parametrize("P", [7, 14])
def test_fused_pool_ard(emu, P):
    rng = np.random.default_rng(P)
    B, C, H, W, V = 2, 40, 20, 30, 4
    t = rng.standard_normal((B, C, H, W)).astype(np.float32)
    s = (t + 0.1 * rng.standard_normal(t.shape)).astype(np.float32)
    rois = make_rois(rng, 24, B, W * 16, H * 16)
    R = len(rois)
    plans = plans_for(emu, rois, H, W, 1 / 16, P, 0)
    nsl = -(-C // (32 * V))
    fo = np.full((R, P, P, C), np.nan, np.float32)
    fn = np.full((R, P, P, C), np.nan, np.float32)
    sums = np.zeros((R, nsl, P * P, 3), np.float32)
    emu.emu_fwd(_p(plans, _i32p), _p(rois), _p(nhwc(t)), _p(nhwc(s)), _p(fo), _p(fn), _p(sums), R, C, H, W, P, P, 1 / 16, 0, V)
    ro, rn = oracle.roi_align_forward(t, rois, 1 / 16, P, P, 0), oracle.roi_align_forward(s, rois, 1 / 16, P, P, 0)
    close(nchw(fo), ro)
    close(nchw(fn), rn)
    tot = sums.sum(1)
    close(tot[..., 0], (ro.astype(np.float64) ** 2).sum(1).reshape(R, -1), 2e-5)
    close(tot[..., 2], ((rn.astype(np.float64) - ro) ** 2).sum(1).reshape(R, -1), 2e-5)
    coef = ard_coefficients(tot, R, C, P * P, 1.0)
    gmap = np.zeros((B, H, W, C), np.float32)
    emu.emu_bwd(_p(plans, _i32p), _p(rois), _p(gmap), _p(fo), _p(fn), _p(coef), R, C, H, W, P, P, 1 / 16, 0, V, 1)
    _, _, _, dfn = oracle.ard(ro, rn, 1.0)
    # RoIs whose attention difference is below fp32 resolution somewhere have an ill-defined sign(): leave them out
    close(nchw(gmap), oracle.roi_align_backward(dfn, rois, 1 / 16, P, P, B, C, H, W, 0), 2e-5)


def test_fused_tall_bins_take_the_per_sample_path(emu):
    """The two-tensor forward keeps 12 map rows per strip; a RoI with a taller bin (13..15 rows) runs the per-sample path
    inside the same kernel, and the fused backward serves it from its plan as usual."""
    rng = np.random.default_rng(9)
    B, C, H, W, P, V = 1, 8, 40, 40, 2, 4
    t = rng.standard_normal((B, C, H, W)).astype(np.float32)
    s = (t + 0.1 * rng.standard_normal(t.shape)).astype(np.float32)
    rois = np.array([[0, 2, 3, 20, 27.5], [0, 5, 1, 30, 28.9], [0, 1, 10, 12, 36.2], [0, 3, 3, 9, 9]], np.float32)
    R = len(rois)
    plans = plans_for(emu, rois, H, W, 1.0, P, 0)
    tallest = [max(int(plans[r, 16 + (P + ph) * 16]) >> 16 for ph in range(P)) for r in range(R)]
    assert max(tallest) in (13, 14, 15) and min(tallest) <= 12 and all(plans[:, 0] == 1)
    fo = np.full((R, P, P, C), np.nan, np.float32)
    fn = np.full((R, P, P, C), np.nan, np.float32)
    sums = np.zeros((R, 1, P * P, 3), np.float32)
    emu.emu_fwd(_p(plans, _i32p), _p(rois), _p(nhwc(t)), _p(nhwc(s)), _p(fo), _p(fn), _p(sums), R, C, H, W, P, P, 1.0, 0, V)
    ro, rn = oracle.roi_align_forward(t, rois, 1.0, P, P, 0), oracle.roi_align_forward(s, rois, 1.0, P, P, 0)
    close(nchw(fo), ro)
    close(nchw(fn), rn)
    close(sums.sum(1)[..., 1], (rn.astype(np.float64) ** 2).sum(1).reshape(R, -1), 2e-5)
    gout = rng.standard_normal(ro.shape).astype(np.float32)
    gmap = np.zeros((B, H, W, C), np.float32)
    emu.emu_bwd(_p(plans, _i32p), _p(rois), _p(gmap), _p(nhwc(gout)), None, None, R, C, H, W, P, P, 1.0, 0, V, 0)
    close(nchw(gmap), oracle.roi_align_backward(gout, rois, 1.0, P, P, B, C, H, W, 0))
