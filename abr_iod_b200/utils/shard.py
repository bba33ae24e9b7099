"""Data-parallel sharding of the RoI hot path: by image, one process per GPU, no data-path collective.

Mirrors what the reference's launcher does for the whole training step: ``DistributedSampler`` hands rank r a contiguous
slice of the epoch's (seeded) image permutation (data/samplers/distributed.py:42-60) and ``images_per_gpu =
IMS_PER_BATCH // world`` (data/build.py:111-113).  Everything on the hot path is per image -- a RoI names its image in
column 0 (csrc/cuda/ROIAlign_cuda.cu:79), NMS and the paste run per image, and the ARD loss is a mean over the local
RoIs whose gradient DDP averages across ranks -- so a rank only ever needs its own images, RoIs and boxes.
"""
import torch
import torch.distributed as dist


def images_per_rank(images_per_batch: int, world_size: int) -> int:
    """data/build.py:111-113"""
    if images_per_batch % world_size != 0:
        raise ValueError("IMS_PER_BATCH (%d) must be divisible by the number of GPUs (%d)" % (images_per_batch, world_size))
    return images_per_batch // world_size


def epoch_shard(num_images: int, rank: int, world_size: int, epoch: int = 0, shuffle: bool = True):
    """Indices of the images rank ``rank`` owns in ``epoch`` -- data/samplers/distributed.py:42-60: a generator seeded
    with the epoch permutes the data set, it is padded to a multiple of the world size by wrapping around, and every
    rank takes a contiguous slice."""
    if shuffle:
        g = torch.Generator()
        g.manual_seed(epoch)
        indices = torch.randperm(num_images, generator=g).tolist()
    else:
        indices = list(range(num_images))
    per_rank = -(-num_images // world_size)
    total = per_rank * world_size
    indices += indices[: total - len(indices)]
    return indices[rank * per_rank: (rank + 1) * per_rank]


def shard_rois(rois: torch.Tensor, image_ids, rebase: bool = True) -> torch.Tensor:
    """Rows of a ``[R,5]`` RoI tensor that belong to ``image_ids`` (a rank's images), with column 0 renumbered to the
    position of the image inside the shard so that it indexes the rank's own feature-map batch."""
    ids = torch.as_tensor(list(image_ids), device=rois.device, dtype=torch.int64)
    col = rois[:, 0].to(torch.int64)
    lut = torch.full((int(max(int(col.max().item()) if rois.numel() else 0, int(ids.max().item()))) + 1,), -1,
                     dtype=torch.int64, device=rois.device)
    lut[ids] = torch.arange(len(ids), device=rois.device)
    local = lut[col]
    keep = local >= 0
    out = rois[keep].clone()
    if rebase:
        out[:, 0] = local[keep].to(rois.dtype)
    return out


def max_over_ranks(values, device=None) -> list:
    """Element-wise max of a few host floats over all ranks (bench.py's device-timed durations); identity when the
    default process group is not initialised."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [float(v) for v in values]
    t = torch.tensor([float(v) for v in values], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()


def sum_over_ranks(values, device=None) -> list:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [float(v) for v in values]
    t = torch.tensor([float(v) for v in values], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.tolist()
