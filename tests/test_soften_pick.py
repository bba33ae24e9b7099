"""select_soften_proposals against a line-by-line restatement of generalized_rcnn.py:128-163 (sort, random.sample over
the first 128, per-box concatenation) under the same `random` seed: identical boxes and scores in identical order."""
import random

import numpy as np
import torch


def reference_pick(bbox, scores):
    inds = scores.sort(descending=True)[1]
    bbox, scores = bbox[inds], scores[inds]
    n = len(bbox)
    if n < 64:
        sel = random.sample(range(0, n, 1), n)
    elif n < 128:
        sel = random.sample(range(0, n, 1), 64)
    else:
        sel = random.sample(range(0, 128, 1), 64)
    b = torch.cat([bbox[e].view(-1, 4) for e in sel], 0)
    s = torch.cat([scores[e].view(-1, 1) for e in sel], 1).view(-1)
    return b, s


def test_soften_pick_matches_reference_stream():
    from abr_iod_b200.modeling.detector import select_soften_proposals
    from abr_iod_b200.structures.bounding_box import BoxList

    rng = np.random.default_rng(0)
    lists = []
    for n in (300, 100, 40, 1):
        b = torch.from_numpy(rng.uniform(0, 500, (n, 4)).astype(np.float32))
        s = torch.from_numpy(rng.permutation(n).astype(np.float32) / n)  # distinct scores
        bl = BoxList(b, (640, 480), "xyxy")
        bl.add_field("objectness", s)
        lists.append(bl)
    random.seed(123)
    ours = select_soften_proposals(lists)
    random.seed(123)
    for bl, got in zip(lists, ours):
        rb, rs = reference_pick(bl.bbox, bl.get_field("objectness"))
        assert torch.equal(got.bbox, rb) and torch.equal(got.get_field("objectness"), rs)
        assert got.size == bl.size and got.mode == bl.mode
    assert [len(x) for x in ours] == [64, 64, 40, 1]
