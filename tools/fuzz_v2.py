"""Randomised parity sweep of the round-2 kernels (run on a GPU box): the gather-form ROIAlign forward / backward forced
onto every output size up to 16x16 (incl. non-square), random shapes / sampling ratios / layouts / adversarial RoIs, and
the fused ARD step (pooled tensors, loss, map gradient) against the CPU oracle chain; the device bicubic resize against
PIL.  Prints one line per family and fails loudly on a mismatch.

    python tools/fuzz_v2.py [cases]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

import oracle
from abr_iod_b200 import _lib
from abr_iod_b200.distillation.distillation import pooled_attentive_roi_distillation
from abr_iod_b200.layers.roi_align import roi_align_backward, roi_align_forward
from inputs import make_rois


def close(a, ref, rel=1e-5, what="", allow=None):
    a, ref = np.asarray(a, np.float64), np.asarray(ref, np.float64)
    scale = np.abs(ref).max() if ref.size else 1.0
    err = np.abs(a - ref)
    bound = rel * scale + rel * np.abs(ref) + (0 if allow is None else 1.01 * np.asarray(allow, np.float64))
    if not (err <= bound).all():
        raise SystemExit("MISMATCH %s: max err %g (scale %g)" % (what, err.max(), scale))


def sign_allowance(f_old, f_new, gamma, thr=3e-5):
    """Exact bound on what fp32-unresolvable signs of (A_new - A_old) change in dL/df_new (see tests/test_gpu_v2.py)."""
    N, C, PH, PW = f_new.shape
    HW = PH * PW
    a_old = HW * torch.softmax((f_old.double() ** 2).mean(1).flatten(1), 1)
    s_new = torch.softmax((f_new.double() ** 2).mean(1).flatten(1), 1)
    amb = ((HW * s_new - a_old).abs() < thr * a_old).double()
    factor = s_new * (amb + (amb * s_new).sum(1, keepdim=True))
    return ((4.0 * gamma / (C * N)) * f_new.double().abs() * factor.view(N, 1, PH, PW)).float().cpu().numpy()


def main():
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    rng = np.random.default_rng(2027)
    _lib.set_option("roi_v2", 1)
    for i in range(cases):
        B = int(rng.integers(1, 4))
        C = int(rng.choice([3, 4, 8, 20, 64, 132, 260]))
        H, W = int(rng.integers(3, 70)), int(rng.integers(3, 100))
        PH, PW = int(rng.integers(1, 17)), int(rng.integers(1, 17))
        if rng.integers(0, 2):
            PW = PH
        ratio = int(rng.choice([0, 0, 1, 2, 3]))
        R = int(rng.integers(1, 60))
        x = rng.standard_normal((B, C, H, W)).astype(np.float32)
        rois = make_rois(rng, R, B, W * 16, H * 16, adversarial=bool(rng.integers(0, 2)))
        xt = torch.from_numpy(x).cuda().contiguous(memory_format=torch.channels_last)
        rt = torch.from_numpy(rois).cuda()
        what = "roi_align v2 B%d C%d %dx%d P%dx%d r%d R%d" % (B, C, H, W, PH, PW, ratio, R)
        out, plan = roi_align_forward(xt, rt, 1 / 16, PH, PW, ratio, return_plan=True)
        close(out.cpu().numpy(), oracle.roi_align_forward(x, rois, 1 / 16, PH, PW, ratio), what=what + " fwd")
        g = rng.standard_normal((R, C, PH, PW)).astype(np.float32)
        gt = torch.from_numpy(g).cuda().contiguous(memory_format=torch.channels_last)
        gin = roi_align_backward(gt, rt, 1 / 16, PH, PW, B, C, H, W, ratio, layout=_lib.ABR_NHWC, plan=plan)
        close(gin.cpu().numpy(), oracle.roi_align_backward(g, rois, 1 / 16, PH, PW, B, C, H, W, ratio), what=what + " bwd")
    print("roi_align v2 forward/backward: %d random cases ok (outputs 1x1 .. 16x16, non-square, ratios 0-3, adversarial RoIs)" % cases)
    _lib.set_option("roi_v2", -1)
    n_fused = max(cases // 2, 1)
    for i in range(n_fused):
        B = int(rng.integers(1, 4))
        C = int(rng.choice([4, 8, 36, 64, 132, 256]))
        H, W = int(rng.integers(4, 50)), int(rng.integers(4, 70))
        P = int(rng.choice([2, 3, 7, 7, 14, 14, 16]))
        ratio = int(rng.choice([0, 0, 2]))
        R = int(rng.integers(1, 50))
        gamma = float(rng.choice([0.3, 1.0, 2.0]))
        t = rng.standard_normal((B, C, H, W)).astype(np.float32)
        s = (t + rng.choice([0.05, 0.3]) * rng.standard_normal(t.shape)).astype(np.float32)
        rois = make_rois(rng, R, B, W * 16, H * 16, adversarial=bool(rng.integers(0, 2)))
        tt = torch.from_numpy(t).cuda().contiguous(memory_format=torch.channels_last)
        st = torch.from_numpy(s).cuda().contiguous(memory_format=torch.channels_last).requires_grad_(True)
        what = "fused B%d C%d %dx%d P%d r%d R%d" % (B, C, H, W, P, ratio, R)
        f_old, f_new, loss = pooled_attentive_roi_distillation(tt, st, torch.from_numpy(rois).cuda(), (P, P), 1 / 16, ratio, gamma)
        loss.backward()
        ro, rn = oracle.roi_align_forward(t, rois, 1 / 16, P, P, ratio), oracle.roi_align_forward(s, rois, 1 / 16, P, P, ratio)
        rl, _, _, dfn = oracle.ard(ro, rn, gamma)
        close(f_old.cpu().numpy(), ro, what=what + " f_old")
        close(f_new.detach().cpu().numpy(), rn, what=what + " f_new")
        if abs(loss.item() - rl) > 2e-5 * abs(rl):
            raise SystemExit("MISMATCH %s loss %g vs %g" % (what, loss.item(), rl))
        allow = oracle.roi_align_backward(sign_allowance(f_old, f_new.detach(), gamma), rois, 1 / 16, P, P, B, C, H, W, ratio)
        close(st.grad.cpu().numpy(), oracle.roi_align_backward(dfn, rois, 1 / 16, P, P, B, C, H, W, ratio), 2e-5, what + " grad", allow)
    print("fused ARD step (pooled tensors, loss, map gradient): %d random cases ok" % n_fused)
    from PIL import Image

    from abr_iod_b200.data.resample import bicubic_taps

    for i in range(cases):
        sh, sw, dh, dw = (int(v) for v in rng.integers(2, 400, 4))
        if sh > 100 * sw:  # Pillow 12 switches to vertical-then-horizontal for sources more than 100x taller than wide
            continue       # (measured: exact threshold sh >= 100*sw + 1 for sw in 2..10); the paster sends those to PIL itself
        src = rng.integers(0, 256, (sh, sw, 3), dtype=np.uint8)
        xt, xk = bicubic_taps(sw, dw)
        yt, yk = bicubic_taps(sh, dh)
        job = _lib.ResizeJob(0, 0, dh * dw * 3, sh, sw, dh, dw, 0, xk, xt.size, yk)
        d_out = torch.zeros(dh * dw * 3 + sh * dw * 3, dtype=torch.uint8, device="cuda")
        d_pool = torch.from_numpy(src.reshape(-1)).cuda()
        d_job = torch.frombuffer(bytearray(bytes(job)), dtype=torch.uint8).cuda()
        d_taps = torch.from_numpy(np.concatenate([xt.reshape(-1), yt.reshape(-1)]).astype(np.int32)).cuda()
        _lib.check(_lib.lib().abr_resize_bicubic_batch(d_pool.data_ptr(), d_out.data_ptr(), d_job.data_ptr(), 1, d_taps.data_ptr(),
                                                       max(sh * dw, dh * dw), _lib.stream_ptr(d_out.device)))
        want = np.asarray(Image.fromarray(src).resize((dw, dh)))
        got = d_out[: dh * dw * 3].cpu().numpy().reshape(dh, dw, 3)
        if not np.array_equal(got, want):
            raise SystemExit("MISMATCH resize %dx%d -> %dx%d: max diff %d" % (sh, sw, dh, dw, np.abs(got.astype(int) - want.astype(int)).max()))
    print("device bicubic resize vs PIL: %d random sizes (2..400 px per side) bit-exact" % cases)


if __name__ == "__main__":
    main()
