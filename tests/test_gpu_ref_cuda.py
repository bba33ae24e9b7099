"""Third parity opinion on the GPU box: the REFERENCE's own CUDA kernels (csrc/cuda/ROIAlign_cuda.cu, ROIPool_cuda.cu,
nms.cu compiled in place for sm_100a into oracle/_ref/libabr_ref_cuda.so by oracle/ref_cuda_shim.cu) against this
library on the same device tensors -- for the rows whose reference has no CPU implementation (ROIAlign backward, ROIPool
forward/backward, the '>' NMS) and, for completeness, ROIAlign forward.  Skipped where the comparator was not built
(it needs /root/reference at build time; the built .so travels to the GPU box)."""
import numpy as np
import pytest
import torch

import oracle
from inputs import make_boxes, make_rois

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not oracle.ref_cuda_available(), reason="oracle/_ref/libabr_ref_cuda.so not built")]


def close(a, ref, rel=1e-5):
    a, ref = np.asarray(a, np.float64), np.asarray(ref, np.float64)
    scale = np.abs(ref).max() if ref.size else 1.0
    err = np.abs(a - ref)
    assert (err <= rel * scale + rel * np.abs(ref)).all(), "max err %g (scale %g)" % (err.max(), scale)


@pytest.mark.parametrize("channels_last", [False, True])
@pytest.mark.parametrize("P,ratio", [(7, 0), (7, 2), (14, 0), (3, 1)])
def test_roi_align_forward_backward_vs_reference_cuda(channels_last, P, ratio):
    from abr_iod_b200.layers import roi_align

    rng = np.random.default_rng(P * 10 + ratio)
    B, C, H, W = 2, 40, 25, 38
    x = torch.from_numpy(rng.standard_normal((B, C, H, W)).astype(np.float32)).cuda()
    rois = torch.from_numpy(make_rois(rng, 60, B, W * 16, H * 16)).cuda()
    xt = (x.contiguous(memory_format=torch.channels_last) if channels_last else x.clone()).requires_grad_(True)
    out = roi_align(xt, rois, (P, P), 1 / 16, ratio)
    ref = oracle.ref_cuda_roi_align_forward(x, rois, 1 / 16, P, P, ratio)
    close(out.detach().cpu().numpy(), ref.cpu().numpy())
    g = torch.from_numpy(rng.standard_normal(tuple(out.shape)).astype(np.float32)).cuda()
    out.backward(g.contiguous(memory_format=torch.channels_last) if channels_last else g)
    gref = oracle.ref_cuda_roi_align_backward(g, rois, 1 / 16, P, P, B, C, H, W, ratio)
    close(xt.grad.cpu().numpy(), gref.cpu().numpy())  # the reference's atomicAdd order is not deterministic: 1e-5, not bits


def test_roi_align_full_size_configs0_vs_reference_cuda():
    """BASELINE.json configs[0] at full size, P = 14: forward and backward over ALL elements against the reference kernels."""
    from abr_iod_b200.layers import roi_align
    from test_gpu_v2 import bench_like_rois

    rng = np.random.default_rng(14)
    B, C, H, W, R, P = 2, 1024, 38, 63, 1024, 14
    x = torch.randn(B, C, H, W, device="cuda")
    rois = torch.from_numpy(bench_like_rois(rng, R, B, 1000, 600)).cuda()
    xt = x.contiguous(memory_format=torch.channels_last).requires_grad_(True)
    out = roi_align(xt, rois, (P, P), 1 / 16, 0)
    ref = oracle.ref_cuda_roi_align_forward(x, rois, 1 / 16, P, P, 0)
    err = (out.detach() - ref).abs().max().item()
    assert err <= 1e-5 * ref.abs().max().item(), err
    del ref
    g = torch.randn(R, C, P, P, device="cuda")
    out.backward(g.contiguous(memory_format=torch.channels_last))
    gref = oracle.ref_cuda_roi_align_backward(g, rois, 1 / 16, P, P, B, C, H, W, 0)
    err = (xt.grad - gref).abs().max().item()
    assert err <= 2e-5 * gref.abs().max().item(), err  # hundreds of float atomics per element in the reference


@pytest.mark.parametrize("channels_last", [False, True])
def test_roi_pool_vs_reference_cuda(channels_last):
    from abr_iod_b200.layers import roi_pool

    rng = np.random.default_rng(5)
    B, C, H, W, P = 2, 24, 25, 38, 7
    x = torch.from_numpy(rng.standard_normal((B, C, H, W)).astype(np.float32)).cuda()
    rois = torch.from_numpy(make_rois(rng, 50, B, W * 16, H * 16)).cuda()
    xt = (x.contiguous(memory_format=torch.channels_last) if channels_last else x.clone()).requires_grad_(True)
    out = roi_pool(xt, rois, (P, P), 1 / 16)
    ref, arg = oracle.ref_cuda_roi_pool_forward(x, rois, 1 / 16, P, P)
    assert torch.equal(out.detach().contiguous(), ref)  # a max: bit-exact
    g = torch.from_numpy(rng.standard_normal(tuple(out.shape)).astype(np.float32)).cuda()
    out.backward(g.contiguous(memory_format=torch.channels_last) if channels_last else g)
    gref = oracle.ref_cuda_roi_pool_backward(g, x, rois, arg, 1 / 16, P, P)
    close(xt.grad.cpu().numpy(), gref.cpu().numpy())


@pytest.mark.parametrize("n", [1, 63, 64, 65, 1000, 6000, 12000])
def test_nms_keep_indices_equal_reference_cuda(n):
    from abr_iod_b200.layers import nms

    rng = np.random.default_rng(n)
    b, s = make_boxes(rng, n)
    bt, st = torch.from_numpy(b).cuda(), torch.from_numpy(s).cuda()
    for thr in (0.5, 0.7):
        ours = nms(bt, st, thr)
        ref = oracle.ref_cuda_nms(bt, st, thr)
        assert torch.equal(ours.cpu(), ref.cpu())
