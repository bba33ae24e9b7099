"""The packed prototype store (CPU): round trip through a file, the naming / cropping rule of
Mem.creat_and_save_box_image (tools/extract_memory.py:213-236), and the numpy restatements of the other two selection
rules against hand-checkable cases."""
import random

import numpy as np
from PIL import Image

from abr_iod_b200.data.prototype_store import PrototypeStore
from oracle import prototype as op


def test_store_round_trip_and_crop_rule(tmp_path):
    rng = np.random.default_rng(0)
    image = rng.integers(0, 256, (120, 160, 3), dtype=np.uint8)
    picks = [(3, 0, image, [10.7, 20.2, 90.9, 100.5]), (3, 1, Image.fromarray(image), [0, 0, 160, 120]), (7, 0, image, [50, 60, 121, 119])]
    store = PrototypeStore.from_boxes(picks)
    assert store.names == ["3_00000.jpg", "3_00001.jpg", "7_00000.jpg"]
    want = np.asarray(Image.fromarray(image).crop((10, 20, 90, 100)))  # the reference: im.crop((int(x1), int(y1), int(x2), int(y2)))
    assert np.array_equal(store.crop(0), want) and store.crop(1).shape == (120, 160, 3)
    store.save(str(tmp_path / "mem.npz"))
    back = PrototypeStore.load(str(tmp_path / "mem.npz"))
    assert back.names == store.names and np.array_equal(back.pixels, store.pixels) and np.array_equal(back.offsets, store.offsets)
    protos = back.prototypes()
    assert protos[2][0] == "7_00000.jpg" and np.array_equal(protos[2][1], image[60:119, 50:121])


def test_herding_restatement_picks_the_mean_first():
    # three descriptors on a line: the middle one is closest to the class mean and is picked first; then the pair that
    # keeps the running centre nearest to it
    f = [np.full((2, 2), v) for v in (1.0, 2.0, 3.0, 10.0)]
    order, source = op.herding_ranking(f, 3)
    assert list(source) == [0, 1, 2, 3] and len(order) == 3 and len(set(order.tolist())) == 3


def test_random_ranking_follows_random_shuffle():
    from abr_iod_b200.tools.prototype_box_selection import random_ranking

    random.seed(3)
    got = random_ranking(7, 4)
    random.seed(3)
    ref = list(range(7))
    random.shuffle(ref)
    assert got == ref[:4]
    random.seed(4)
    got = random_ranking(2, 5)   # top-up: the shuffled list followed by its first entries
    random.seed(4)
    ref = [0, 1]
    random.shuffle(ref)
    assert got == (ref + ref[:3])[:5]


def test_sample_fg_bg_restatement_counts():
    rng = np.random.default_rng(1)
    m = rng.integers(-1, 4, 300)
    keys = rng.uniform(0, 1, 300).astype(np.float32)
    pos, neg = op.sample_fg_bg(m, keys, 64, 0.25)
    assert pos.sum() == min((m >= 1).sum(), 16) and neg.sum() == min((m == 0).sum(), 64 - pos.sum())
    assert not (pos & neg).any() and (m[pos == 1] >= 1).all() and (m[neg == 1] == 0).all()


def test_planner_runs_without_cuda_and_its_plans_pickle():
    """BoxRehearsalPlanner (the host half of the paste) needs only prototype names and sizes: it runs in a CPU-only worker
    and its plans travel between processes; a seeded planner reproduces its own draws."""
    import pickle

    import torch

    from abr_iod_b200.data.abr_paste import BoxRehearsalPlanner

    rng = np.random.default_rng(5)
    names = ["%d_%05d.jpg" % (1 + i % 15, i) for i in range(40)]
    sizes = [(int(rng.integers(71, 301)), int(rng.integers(71, 301))) for _ in names]
    images = [Image.fromarray(rng.integers(0, 256, (200, 260, 3), dtype=np.uint8)) for _ in range(12)]
    targets = [np.array([[20.0, 30.0, 120.0, 110.0, 17.0]]) for _ in images]
    runs = []
    for _ in range(2):
        planner = BoxRehearsalPlanner(names, sizes, batch_size=4)
        random.seed(9)
        torch.manual_seed(9)
        plans = [planner.plan_transform(im, g) for im, g in zip(images, targets)]
        runs.append(pickle.loads(pickle.dumps(plans)))
    assert {p.kind for p in runs[0]} == {"none", "mixup", "mosaic"}
    for a, b in zip(*runs):
        assert a.kind == b.kind and np.array_equal(a.gts, b.gts) and len(a.ops) == len(b.ops)
        assert all(x.dst == y.dst and x.proto == y.proto and x.resize == y.resize for x, y in zip(a.ops, b.ops))
