"""GPU parity of the ABR mixup / mosaic paste: bit-exact pixels, boxes, labels and Box-Rehearsal index state against
the golden stream recorded from the reference's own PascalVOCDataset_ABR (tests/golden/make_golden.py)."""
import random

import numpy as np
import pytest
import torch
from PIL import Image

from oracle import paste as opaste

pytestmark = pytest.mark.gpu


def make_paster(g, batch_size=4):
    from abr_iod_b200.data.abr_paste import BoxRehearsalPaster

    names = [str(n) for n in g["proto_names"]]
    return BoxRehearsalPaster([(n, g["proto_%02d" % i]) for i, n in enumerate(names)], batch_size)


def test_paste_golden_stream_bit_exact(golden):
    from abr_iod_b200.structures.bounding_box import BoxList

    g = golden("paste.npz")
    paster = make_paster(g)
    seen = set()
    for case in g["cases"]:
        key, kind, seed = str(case).split(":")
        random.seed(int(seed))
        torch.manual_seed(int(seed))
        img, gts = g[key + "_img"], g[key + "_gts"]
        target = BoxList(torch.tensor(gts[:, :4]), (img.shape[1], img.shape[0]), "xyxy")
        target.add_field("labels", torch.tensor(gts[:, 4]).long())
        pil = Image.fromarray(img)
        if kind == "mixup":
            o_img, o_t = paster._start_mixup(pil, target)
        elif kind == "mosaic":
            o_img, o_t = paster._start_boxes_mosaic(pil, [], num_boxes=4)
        else:
            o_img, o_t = paster.transform_current_data_with_ABR(pil, target)
        seen.add(kind)
        assert np.array_equal(np.array(o_img), g[key + "_out_img"]), key
        assert np.array_equal(o_t.bbox.numpy(), g[key + "_out_bbox"]), key
        labels = o_t.get_field("labels").numpy()
        assert np.array_equal(labels, g[key + "_out_labels"]) and labels.dtype == g[key + "_out_labels"].dtype, key
        assert tuple(o_t.size) == tuple(g[key + "_out_size"]), key
        assert paster.boxes_index == g[key + "_index_after"].tolist(), key
    assert seen == {"mixup", "mosaic", "auto"}


def test_paste_batch_one_launch_matches_oracle(golden):
    """A whole batch (mixed mixup / mosaic / untouched, prototypes that need a host-side resize) in one launch."""
    from abr_iod_b200 import _lib

    g = golden("paste.npz")
    names = [str(n) for n in g["proto_names"]]
    paster = make_paster(g, batch_size=8)
    st = opaste.BoxRehearsalState([(n, opaste.as_pil(g["proto_%02d" % i])) for i, n in enumerate(names)], 8)
    rng = np.random.default_rng(5)
    images, targets = [], []
    for i in range(16):
        h, w = int(rng.integers(120, 260)), int(rng.integers(120, 260))
        images.append(Image.fromarray(rng.integers(0, 256, (h, w, 3), dtype=np.uint8)))
        x1, y1 = rng.uniform(0, w * 0.5, 2), rng.uniform(0, h * 0.5, 2)
        targets.append(np.stack([x1, y1, x1 + rng.uniform(10, w * 0.4, 2), y1 + rng.uniform(10, h * 0.4, 2),
                                 rng.integers(16, 21, 2)], 1))
    random.seed(77); torch.manual_seed(77)
    ref = [opaste.transform_current_data_with_abr(st, im, t) for im, t in zip(images, targets)]
    random.seed(77); torch.manual_seed(77)
    before = _lib.launch_count()
    outs, gts, kinds = paster.paste_batch(images, targets)
    assert _lib.launch_count() - before == 1
    assert kinds == [r[0] for r in ref] and len(set(kinds)) == 3
    for o, gt, (_, rimg, rgt) in zip(outs, gts, ref):
        assert np.array_equal(o.cpu().numpy(), rimg)
        assert np.array_equal(gt, rgt)
    assert paster.boxes_index == st.boxes_index
