"""GPU parity of the Prototype Box Selection scoring (channel-mean descriptors + nearest-to-class-mean ranking) through
the Python mirror against golden vectors produced by the reference's Mem.mean_feature_sampling and against the oracle.
Selected boxes (integer work) must be identical; descriptors are fp32 means over C channels (different summation order
than torch's CPU mean): |a-b| <= 1e-5 * max|ref|; distances are float64: relative 1e-12."""
import numpy as np
import pytest
import torch

from oracle import prototype as op

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("channels_last", [False, True])
def test_prototype_selection_golden_vs_reference_python(golden, channels_last):
    from abr_iod_b200.tools.prototype_box_selection import mean_feature_ranking, roi_descriptors

    g = golden("prototype.npz")
    per_cls = int(g["per_cls"])
    for c in range(3):
        x = torch.from_numpy(g["pooled_%d" % c]).cuda()
        if channels_last:
            x = x.contiguous(memory_format=torch.channels_last)
        desc = roi_descriptors(x)
        ref = g["desc_%d" % c]
        assert np.abs(desc.cpu().numpy() - ref).max() <= 1e-5 * np.abs(ref).max()
        # ranking of the reference's own descriptors: identical selection
        order, dist, source = mean_feature_ranking(torch.from_numpy(ref).cuda(), per_cls)
        assert np.array_equal(source[order].cpu().numpy(), g["selected_%d" % c])
        _, odist, _ = op.mean_feature_ranking(list(ref), per_cls)
        assert np.abs(dist.cpu().numpy() - odist).max() <= 1e-12 * odist.max()
        # end to end from the device descriptors: the same boxes (distances are well separated in this fixture)
        order2, _, source2 = mean_feature_ranking(desc, per_cls)
        assert np.array_equal(source2[order2].cpu().numpy(), g["selected_%d" % c])


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_roi_descriptors_full_size_and_bf16(dtype):
    from abr_iod_b200.tools.prototype_box_selection import roi_descriptors

    x = torch.randn(64, 1024, 7, 7, device="cuda").to(dtype)
    ref = x.float().mean(dim=1)
    for xx in (x, x.contiguous(memory_format=torch.channels_last)):
        got = roi_descriptors(xx)
        assert got.dtype == torch.float32 and got.shape == (64, 7, 7)
        assert (got - ref).abs().max().item() <= 1e-5 * ref.abs().max().item() + 1e-6
    assert roi_descriptors(torch.zeros((0, 8, 7, 7), device="cuda")).shape == (0, 7, 7)
    with pytest.raises(RuntimeError):
        roi_descriptors(torch.zeros((2, 8, 7, 7)))


def test_herding_selection_equals_the_numpy_restatement():
    """abr_prototype_herding against the float64 restatement of Mem.herding_feature_sampling's loop
    (tools/extract_memory.py:163-197; the reference's function itself stops with an unbound variable at :203 and cannot
    produce a golden run): the same boxes in the same order, incl. a topped-up class."""
    from abr_iod_b200.tools.prototype_box_selection import herding_ranking

    rng = np.random.default_rng(7)
    for n, per_cls in ((40, 6), (3, 6), (200, 25), (17, 17)):
        feats = (rng.standard_normal((n, 7, 7)) + rng.uniform(0, 2, (1, 7, 7))).astype(np.float32)
        order, source = herding_ranking(torch.from_numpy(feats).cuda(), per_cls)
        want, wsource = op.herding_ranking(list(feats), per_cls)
        assert np.array_equal(source.cpu().numpy(), wsource)
        assert np.array_equal(order.cpu().numpy(), want)


def test_device_fg_bg_sampling_counts_and_subset():
    """abr_sample_fg_bg: the reference's per-image counts (balanced_positive_negative_sampler.py:19-68) and, for the same
    keys, exactly the subset of the numpy restatement -- whole batch, one launch."""
    from abr_iod_b200.modeling.balanced_positive_negative_sampler import BalancedPositiveNegativeSampler

    rng = np.random.default_rng(2)
    sizes = [2003, 517, 1, 64, 1200]
    matched = [rng.integers(-1, 3, n) for n in sizes]
    matched[2][:] = 0
    matched[3][:] = 2      # positives only: min(#pos, 128) positives, no negatives
    keys = rng.uniform(0, 1, sum(sizes)).astype(np.float32)
    keys[5:9] = keys[4]    # equal keys: ties go to the lower index
    sampler = BalancedPositiveNegativeSampler(512, 0.25, device_sampling=True)
    pos, neg, counts = sampler.sample_on_device([torch.from_numpy(m).cuda() for m in matched], torch.from_numpy(keys).cuda())
    at = 0
    for i, (m, n) in enumerate(zip(matched, sizes)):
        wp, wn = op.sample_fg_bg(m, keys[at: at + n], 512, 0.25)
        assert np.array_equal(pos[i].cpu().numpy(), wp) and np.array_equal(neg[i].cpu().numpy(), wn)
        assert counts[i].tolist() == [int(wp.sum()), int(wn.sum())]
        at += n
    # through __call__ (own torch.rand draw): right counts, right classes
    p2, n2 = sampler([torch.from_numpy(m).cuda() for m in matched])
    for m, a, b in zip(matched, p2, n2):
        a, b = a.cpu().numpy(), b.cpu().numpy()
        assert a.sum() == min((m >= 1).sum(), 128) and b.sum() == min((m == 0).sum(), 512 - a.sum())
        assert (m[a == 1] >= 1).all() and (m[b == 1] == 0).all()
