"""GPU parity of the fused Attentive-RoI-Distillation kernel (loss + gradient) against the reference's own Python
(golden vectors with autograd gradients) and the CPU oracle.  fp32 tolerance 1e-5 relative (north_star)."""
import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu


def ard_gpu(fo, fn, gamma, channels_last=False, beta=None):
    from abr_iod_b200.distillation.distillation import calculate_attentive_roi_feature_distillation as ard

    fmt = torch.channels_last if channels_last else torch.contiguous_format
    t_o = torch.as_tensor(fo).cuda().contiguous(memory_format=fmt)
    t_n = torch.as_tensor(fn).cuda().contiguous(memory_format=fmt).requires_grad_(True)
    loss = ard(t_o, t_n, gamma)  # call-site order (teacher, student): tools/train_incremental.py:115
    assert loss.dim() == 0 and loss.requires_grad
    (loss if beta is None else loss * beta).backward()
    return loss.item(), t_n.grad


@pytest.mark.parametrize("channels_last", [False, True])
def test_ard_golden_vs_reference_python(golden, channels_last):
    g = golden("ard.npz")
    for tag in "abc":
        for gamma in (1.0, 0.25):
            k = "%s_g%d" % (tag, int(gamma * 100))
            loss, grad = ard_gpu(g["fo_" + tag], g["fn_" + tag], gamma, channels_last)
            ref = float(g["loss64_" + k])
            assert abs(loss - ref) <= 1e-5 * abs(ref), (k, loss, ref)
            assert abs(loss - float(g["loss32_" + k])) <= 1e-5 * abs(ref)
            gref = g["grad64_" + k]
            err = np.abs(grad.cpu().numpy() - gref)
            assert (err <= 1e-5 * np.abs(gref).max() + 1e-5 * np.abs(gref)).all(), (k, err.max(), np.abs(gref).max())


@pytest.mark.parametrize("channels_last", [False, True])
@pytest.mark.parametrize("N,C,P", [(5, 1024, 7), (3, 256, 14), (2, 37, 7), (1, 8, 1), (2, 16, 28), (4, 130, 5),
                                   (3, 64, 7), (3, 2048, 7), (2, 900, 14)])
def test_ard_vs_oracle_shapes(channels_last, N, C, P):
    rng = np.random.default_rng(N * 1000 + C + P)
    fo = rng.standard_normal((N, C, P, P)).astype(np.float32)
    fn = (fo + 0.2 * rng.standard_normal((N, C, P, P))).astype(np.float32)
    loss, grad = ard_gpu(fo, fn, 0.7, channels_last)
    rl, _, _, rg = oracle.ard(fo, fn, 0.7)
    assert abs(loss - rl) <= 1e-5 * abs(rl)
    err = np.abs(grad.cpu().numpy() - rg)
    assert (err <= 1e-5 * np.abs(rg).max() + 1e-5 * np.abs(rg)).all(), (err.max(), np.abs(rg).max())


def test_ard_terms_upstream_gradient_and_edge_cases():
    from abr_iod_b200.distillation.distillation import attentive_roi_distillation_terms
    from abr_iod_b200.distillation.distillation import calculate_attentive_roi_feature_distillation as ard

    rng = np.random.default_rng(7)
    fo = rng.standard_normal((4, 64, 7, 7)).astype(np.float32)
    fn = (fo + 0.3 * rng.standard_normal(fo.shape)).astype(np.float32)
    rl, rafd, rpad, rg = oracle.ard(fo, fn, 2.0)
    terms = attentive_roi_distillation_terms(torch.from_numpy(fo).cuda(), torch.from_numpy(fn).cuda(), 2.0).cpu().numpy()
    np.testing.assert_allclose(terms, [rl, rafd, rpad], rtol=1e-5)
    # upstream gradient != 1 (cfg.DIST.BETA and amp loss scale, train_incremental.py:116,144)
    _, grad = ard_gpu(fo, fn, 2.0, beta=0.375)
    np.testing.assert_allclose(grad.cpu().numpy(), 0.375 * rg, rtol=1e-5, atol=1e-5 * np.abs(rg).max())
    # identical teacher and student: loss 0 and, sign(0) = 0, zero gradient
    loss, grad = ard_gpu(fo, fo.copy(), 1.0)
    assert loss == 0.0 and not grad.any()
    # large activations: the softmax must not overflow
    big = (30 * fo).astype(np.float32)
    bign = (big + rng.standard_normal(fo.shape)).astype(np.float32)
    loss, grad = ard_gpu(big, bign, 1.0)
    rl, _, _, rg = oracle.ard(big, bign, 1.0)
    assert np.isfinite(loss) and abs(loss - rl) <= 1e-5 * abs(rl)
    assert np.abs(grad.cpu().numpy() - rg).max() <= 1e-5 * np.abs(rg).max()
    # the teacher side carries no gradient in the reference; asking for one is an error, not silence
    t_o = torch.from_numpy(fo).cuda().requires_grad_(True)
    t_n = torch.from_numpy(fn).cuda().requires_grad_(True)
    with pytest.raises(RuntimeError):
        ard(t_o, t_n, 1.0).backward()
    # no grad requested -> plain scalar
    assert not ard(torch.from_numpy(fo).cuda(), torch.from_numpy(fn).cuda(), 1.0).requires_grad


@pytest.mark.parametrize("channels_last", [False, True])
def test_ard_bf16_stated_tolerance(channels_last):
    rng = np.random.default_rng(9)
    fo = torch.from_numpy(rng.standard_normal((4, 256, 7, 7)).astype(np.float32)).bfloat16()
    fn = (fo.float() + 0.2 * torch.from_numpy(rng.standard_normal(fo.shape).astype(np.float32))).bfloat16()
    loss, grad = ard_gpu(fo, fn, 1.0, channels_last)
    rl, _, _, rg = oracle.ard(fo.float().numpy(), fn.float().numpy(), 1.0)
    assert abs(loss - rl) <= 1e-2 * abs(rl)  # stated bf16 tolerance on the loss
    assert grad.dtype == torch.bfloat16
    assert np.abs(grad.float().cpu().numpy() - rg).max() <= 2e-2 * np.abs(rg).max()


def test_ard_full_size_config(golden):
    """Real step size: 64 RoIs/img x 4 img, C=1024, P=7, checked against the C oracle."""
    rng = np.random.default_rng(10)
    fo = rng.standard_normal((256, 1024, 7, 7)).astype(np.float32)
    fn = (fo + 0.1 * rng.standard_normal(fo.shape)).astype(np.float32)
    for cl in (False, True):
        loss, grad = ard_gpu(fo, fn, 1.0, cl)
        rl, _, _, rg = oracle.ard(fo, fn, 1.0)
        assert abs(loss - rl) <= 1e-5 * abs(rl)
        assert np.abs(grad.cpu().numpy() - rg).max() <= 1e-5 * np.abs(rg).max()
