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
    # one paste launch for the whole batch (+ the two resampling passes when prototypes are rescaled on the device)
    assert _lib.launch_count() - before in (1, 3)
    assert kinds == [r[0] for r in ref] and len(set(kinds)) == 3
    for o, gt, (_, rimg, rgt) in zip(outs, gts, ref):
        assert np.array_equal(o.cpu().numpy(), rimg)
        assert np.array_equal(gt, rgt)
    assert paster.boxes_index == st.boxes_index


def test_device_bicubic_resize_is_bit_exact_with_pil():
    """abr_resize_bicubic_batch against PIL's Image.resize (the call the reference makes, voc_abr.py:548; default filter
    BICUBIC): up- and down-scaling, one axis unchanged, extreme aspect ratios -- every byte equal."""
    import ctypes

    from abr_iod_b200 import _lib
    from abr_iod_b200.data.resample import bicubic_taps

    rng = np.random.default_rng(21)
    shapes = [(120, 200, 60, 90), (71, 300, 150, 100), (250, 250, 133, 140), (80, 90, 200, 310), (100, 100, 100, 57),
              (64, 64, 31, 64), (300, 71, 37, 290), (90, 130, 91, 131)]
    shapes += [tuple(int(v) for v in rng.integers(20, 320, 4)) for _ in range(12)]
    srcs = [rng.integers(0, 256, (sh, sw, 3), dtype=np.uint8) for sh, sw, _, _ in shapes]
    pool = np.concatenate([s.reshape(-1) for s in srcs])
    jobs, tables, tap_at, cursor, src_at, max_pix = [], [], 0, 0, 0, 0
    dst_at = []
    for (sh, sw, dh, dw), src in zip(shapes, srcs):
        xt, xk = bicubic_taps(sw, dw)
        yt, yk = bicubic_taps(sh, dh)
        jobs.append(_lib.ResizeJob(src_at, cursor, cursor + dh * dw * 3, sh, sw, dh, dw, tap_at, xk, tap_at + xt.size, yk))
        tables += [xt.reshape(-1), yt.reshape(-1)]
        tap_at += xt.size + yt.size
        dst_at.append(cursor)
        cursor += dh * dw * 3 + sh * dw * 3
        src_at += src.size
        max_pix = max(max_pix, sh * dw, dh * dw)
    d_pool = torch.from_numpy(pool).cuda()
    d_out = torch.zeros(cursor, dtype=torch.uint8, device="cuda")
    d_jobs = torch.frombuffer(bytearray(b"".join(bytes(j) for j in jobs)), dtype=torch.uint8).cuda()
    d_taps = torch.from_numpy(np.concatenate(tables).astype(np.int32)).cuda()
    _lib.check(_lib.lib().abr_resize_bicubic_batch(d_pool.data_ptr(), d_out.data_ptr(), d_jobs.data_ptr(), len(jobs),
                                                   d_taps.data_ptr(), max_pix, _lib.stream_ptr(d_out.device)))
    out = d_out.cpu().numpy()
    for (sh, sw, dh, dw), src, at in zip(shapes, srcs, dst_at):
        want = np.asarray(Image.fromarray(src).resize((dw, dh)))
        got = out[at: at + dh * dw * 3].reshape(dh, dw, 3)
        assert np.array_equal(got, want), (sh, sw, dh, dw, np.abs(got.astype(int) - want.astype(int)).max())
    assert ctypes.sizeof(_lib.ResizeJob) == 56


def test_paste_with_device_resize_equals_host_resize():
    """A seeded batch with rescaled prototypes: resampling them on the GPU gives the same pixels and boxes as PIL on the host."""
    from abr_iod_b200.data.abr_paste import BoxRehearsalPaster

    rng = np.random.default_rng(8)
    protos = [("%d_%03d.jpg" % (1 + i % 15, i), rng.integers(0, 256, (int(rng.integers(71, 301)), int(rng.integers(71, 301)), 3), dtype=np.uint8))
              for i in range(60)]
    images = [Image.fromarray(rng.integers(0, 256, (375, 500, 3), dtype=np.uint8)) for _ in range(16)]
    targets = [np.array([[20.0, 30.0, 200.0, 180.0, 17.0], [250.0, 100.0, 420.0, 300.0, 18.0]]) for _ in range(16)]
    outs = []
    for device_resize in (True, False):
        paster = BoxRehearsalPaster(protos, batch_size=16, device="cuda", device_resize=device_resize)
        random.seed(5)
        torch.manual_seed(5)
        imgs, gts, kinds = paster.paste_batch(images, targets)
        outs.append(([i.cpu().numpy() for i in imgs], gts, kinds))
    assert outs[0][2] == outs[1][2] and {"mixup", "mosaic"} <= set(outs[0][2])
    for a, b in zip(outs[0][0], outs[1][0]):
        assert np.array_equal(a, b)
    for a, b in zip(outs[0][1], outs[1][1]):
        assert np.array_equal(a, b)
