"""World-size-2 checks (gloo, CPU) of the data-parallel plumbing around the hot path: shards partition the images and
the RoIs, and the cross-rank reductions bench.py relies on give every rank the same answer."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from abr_iod_b200.utils import shard


def test_epoch_shards_partition_the_dataset():
    for n, world in ((10, 2), (11, 4), (7, 8), (16, 1)):
        parts = [shard.epoch_shard(n, r, world, epoch=3) for r in range(world)]
        assert len({len(p) for p in parts}) == 1
        flat = [i for p in parts for i in p]
        assert set(flat) == set(range(n)) and len(flat) == -(-n // world) * world
    assert shard.epoch_shard(10, 0, 2, epoch=1) != shard.epoch_shard(10, 0, 2, epoch=2)
    assert shard.images_per_rank(16, 8) == 2
    with pytest.raises(ValueError):
        shard.images_per_rank(6, 4)


def test_shard_rois_rebases_image_index():
    rois = torch.tensor([[0, 1, 1, 2, 2], [3, 5, 5, 6, 6], [1, 0, 0, 9, 9], [3, 7, 7, 8, 8]], dtype=torch.float32)
    out = shard.shard_rois(rois, [3, 0])
    assert out[:, 0].tolist() == [1.0, 0.0, 0.0] and out[:, 1].tolist() == [1.0, 5.0, 7.0]
    assert shard.shard_rois(rois, [2]).shape == (0, 5)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # each rank owns its images and RoIs; nothing on the data path is exchanged
        images = shard.epoch_shard(8, rank, world, epoch=0)
        rng = np.random.default_rng(0)
        rois = torch.from_numpy(np.concatenate([rng.integers(0, 8, (40, 1)), rng.uniform(0, 100, (40, 4))], 1)).float()
        mine = shard.shard_rois(rois, images)
        counts = shard.sum_over_ranks([len(images), mine.shape[0]])
        # bench.py: the step time is the max over ranks, the work the sum
        t = shard.max_over_ranks([1.0 + rank, 5.0 - rank])
        out.put((rank, images, mine.shape[0], counts, t))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_ranks_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res[0][1] + res[1][1]) == list(range(8))          # the shards partition the images
    assert res[0][2] + res[1][2] == 40                               # ... and the RoIs
    assert res[0][3] == res[1][3] == [8.0, 40.0]                     # every rank sees the same totals
    assert res[0][4] == res[1][4] == [2.0, 5.0]                      # max over ranks
