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
