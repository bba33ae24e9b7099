"""GPU parity of NMS (bit-exact keep indices) through layers.nms / boxlist_nms / the batched entry points."""
import numpy as np
import pytest
import torch

import oracle
from oracle import pooler as opooler
from inputs import make_boxes

pytestmark = pytest.mark.gpu


def gpu_nms(b, s, thr):
    from abr_iod_b200.layers import nms

    return nms(torch.from_numpy(b).cuda(), torch.from_numpy(s).cuda(), thr)


@pytest.mark.parametrize("n", [1, 2, 63, 64, 65, 127, 128, 129, 1000, 6000, 12000])
def test_nms_keep_indices_bit_exact(n):
    rng = np.random.default_rng(n)
    b, s = make_boxes(rng, n)
    for thr in (0.7, 0.5) if n <= 6000 else (0.7,):
        keep = gpu_nms(b, s, thr)
        assert keep.dtype == torch.int64 and keep.is_cuda
        ref = oracle.nms(b, s, thr, "cuda")
        assert np.array_equal(keep.cpu().numpy(), ref), (n, thr, len(ref))


def test_nms_golden_from_compiled_reference(golden):
    """csrc/cpu/nms_cpu.cpp differs from csrc/cuda/nms.cu only on exact IoU == thr ties; cpu_tie_rule selects it."""
    from abr_iod_b200.layers import nms_batched

    g = golden("nms_cpu.npz")
    for n in (1, 63, 64, 65, 300, 1500):
        b, s = g["boxes_%d" % n], g["scores_%d" % n]
        for thr in (0.5, 0.7):
            keep, cnt = nms_batched([torch.from_numpy(b).cuda()], [torch.from_numpy(s).cuda()], thr, cpu_tie_rule=True)
            assert np.array_equal(keep[0, : int(cnt[0])].cpu().numpy(), g["keep_%d_t%d" % (n, int(thr * 10))])


def test_nms_exact_threshold_tie_and_duplicates():
    from abr_iod_b200.layers import nms_batched

    a = np.array([[0, 0, 9, 9], [0, 0, 9, 4]], np.float32)  # IoU exactly 0.5 with the +1 convention
    s = np.array([0.9, 0.8], np.float32)
    assert gpu_nms(a, s, 0.5).tolist() == [0, 1]  # '>' (csrc/cuda/nms.cu:60)
    keep, cnt = nms_batched([torch.from_numpy(a).cuda()], [torch.from_numpy(s).cuda()], 0.5, cpu_tie_rule=True)
    assert keep[0, : int(cnt[0])].tolist() == [0]  # '>=' (csrc/cpu/nms_cpu.cpp:60)
    # duplicates and equal scores: ties resolve by ascending index
    rng = np.random.default_rng(0)
    b, _ = make_boxes(rng, 500)
    b[250:] = b[:250]
    s = np.full(500, 0.5, np.float32)
    s[::3] = 0.75
    assert np.array_equal(gpu_nms(b, s, 0.7).cpu().numpy(), oracle.nms(b, s, 0.7, "cuda"))
    s2 = rng.uniform(0, 1, 500).astype(np.float32)
    s2[::7] = s2[3]
    assert np.array_equal(gpu_nms(b, s2, 0.3).cpu().numpy(), oracle.nms(b, s2, 0.3, "cuda"))


def test_nms_empty_returns_cpu_tensor_like_reference():
    from abr_iod_b200.layers import nms

    k = nms(torch.zeros((0, 4), device="cuda"), torch.zeros((0,), device="cuda"), 0.5)
    assert k.device.type == "cpu" and k.dtype == torch.int64 and k.numel() == 0  # csrc/nms.h:17-18


def test_nms_unsorted_input_ascending_output():
    b = np.array([[0, 0, 10, 10], [100, 100, 110, 110], [1, 1, 11, 11]], np.float32)
    s = np.array([0.1, 0.5, 0.9], np.float32)
    assert gpu_nms(b, s, 0.5).tolist() == [1, 2]


@pytest.mark.parametrize("max_proposals", [-1, 100, 2000])
def test_nms_batched_ragged_matches_per_image_oracle(max_proposals):
    from abr_iod_b200.layers import nms_batched

    rng = np.random.default_rng(21)
    sizes = [6000, 1, 0, 777, 64, 3000, 12000 if max_proposals == 2000 else 129]
    data = [make_boxes(rng, n) if n else (np.zeros((0, 4), np.float32), np.zeros((0,), np.float32)) for n in sizes]
    keep, cnt = nms_batched([torch.from_numpy(b).cuda() for b, _ in data], [torch.from_numpy(s).cuda() for _, s in data],
                            0.7, max_proposals)
    cnt = cnt.cpu().numpy()
    for i, (b, s) in enumerate(data):
        ref = opooler.boxlist_nms(b, s, 0.7, max_proposals) if len(b) else np.zeros(0, np.int64)
        assert cnt[i] == len(ref), (i, cnt[i], len(ref))
        row = keep[i].cpu().numpy()
        assert np.array_equal(row[: cnt[i]], ref)
        assert (row[cnt[i]:] == -1).all()


def test_boxlist_nms_golden(golden):
    from abr_iod_b200.structures.bounding_box import BoxList
    from abr_iod_b200.structures.boxlist_ops import boxlist_nms, boxlist_nms_batched

    g = golden("boxlist_nms.npz")
    b, s, lab = g["boxes"], g["scores"], g["labels"]
    for mode in ("xyxy", "xywh"):
        bl = BoxList(torch.from_numpy(b).cuda(), (1000, 600), "xyxy").convert(mode)
        bl.add_field("scores", torch.from_numpy(s).cuda())
        bl.add_field("labels", torch.from_numpy(lab).cuda())
        for thr, maxp in ((0.7, 50), (0.5, -1), (0.0, -1)):
            k = "%s_t%d_m%d" % (mode, int(thr * 10), maxp)
            for r in (boxlist_nms(bl, thr, max_proposals=maxp, score_field="scores"),
                      boxlist_nms_batched([bl, bl], thr, max_proposals=maxp, score_field="scores")[1]):
                assert r.mode == mode
                np.testing.assert_allclose(r.bbox.cpu().numpy(), g["bbox_" + k], rtol=0, atol=1e-4)
                assert np.array_equal(r.get_field("scores").cpu().numpy(), g["scores_" + k])
                assert np.array_equal(r.get_field("labels").cpu().numpy(), g["labels_" + k])


@pytest.mark.parametrize("n,max_proposals", [(6000, 1000), (12000, 2000), (3000, 2000), (700, 300)])
def test_nms_sorted_input_prefix_pass(n, max_proposals):
    """RPN-shaped call: scores already sorted (top-k output) and a post-NMS cut.  The kernel then settles the image
    from a prefix of the boxes; the result must still equal the full greedy NMS cut to max_proposals -- also when the
    prefix does not contain enough survivors (heavy duplication) and the full pass has to redo the image."""
    from abr_iod_b200.layers import nms_batched

    rng = np.random.default_rng(n + max_proposals)
    imgs = []
    for kind in ("plain", "duplicates", "unsorted"):
        b, s = make_boxes(rng, n, 1216, 800)
        if kind == "duplicates":  # ~everything suppressed: only a handful of distinct boxes, so the prefix is not enough
            b = b[rng.integers(0, 40, n)] + rng.normal(0, 0.5, (n, 4)).astype(np.float32)
        if kind != "unsorted":
            o = np.argsort(-s, kind="stable")
            b, s = np.ascontiguousarray(b[o]), np.ascontiguousarray(s[o])
        imgs.append((b, s))
    keep, cnt = nms_batched([torch.from_numpy(b).cuda() for b, _ in imgs], [torch.from_numpy(s).cuda() for _, s in imgs],
                            0.7, max_proposals)
    cnt = cnt.cpu().numpy()
    for i, (b, s) in enumerate(imgs):
        ref = opooler.boxlist_nms(b, s, 0.7, max_proposals)
        assert cnt[i] == len(ref), (i, cnt[i], len(ref))
        assert np.array_equal(keep[i, : cnt[i]].cpu().numpy(), ref)
