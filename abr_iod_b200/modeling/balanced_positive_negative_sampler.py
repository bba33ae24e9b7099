"""``BalancedPositiveNegativeSampler`` with the reference's interface (modeling/balanced_positive_negative_sampler.py:5-74).

Two modes:
* default -- the reference's RANDOM STREAM: the same two ``torch.randperm`` calls per image, in the same order and with the
  same sizes, so that a seeded run picks the same RoIs as the reference.  The sizes are data-dependent
  (``positive.numel()``), so this form needs the reference's per-image ``nonzero`` synchronisations;
* ``device_sampling=True`` -- ``abr_sample_fg_bg``: ONE launch for the whole batch and no synchronisation.  Every image gets
  exactly the reference's counts (``min(#pos, int(B*f))`` positives, ``min(#neg, B - num_pos)`` negatives), drawn as a
  uniformly random subset from ONE ``torch.rand`` call (its own stream: a randperm prefix cannot be reproduced without
  knowing the counts on the host)."""
import torch

from .. import _lib


class BalancedPositiveNegativeSampler(object):
    def __init__(self, batch_size_per_image, positive_fraction, device_sampling=False):
        self.batch_size_per_image = batch_size_per_image
        self.positive_fraction = positive_fraction
        self.device_sampling = device_sampling

    def sample_on_device(self, matched_idxs, keys=None):
        """The whole batch in one launch.  ``keys``: optional uniform floats, one per RoI of the concatenated batch (drawn
        with ``torch.rand`` when absent).  Returns (pos masks, neg masks) as per-image views and counts [n_images, 2]."""
        sizes = [int(m.shape[0]) for m in matched_idxs]
        flat = torch.cat([m.reshape(-1) for m in matched_idxs]).to(torch.int64).contiguous()
        _lib.require_cuda(flat, "matched_idxs")
        dev = flat.device
        if keys is None:
            keys = torch.rand((flat.shape[0],), device=dev)
        keys = keys.to(torch.float32).contiguous()
        offs = [0]
        for n in sizes:
            offs.append(offs[-1] + n)
        offsets = torch.tensor(offs, dtype=torch.int32).to(dev, non_blocking=True)
        pos = torch.empty((flat.shape[0],), dtype=torch.uint8, device=dev)
        neg = torch.empty_like(pos)
        counts = torch.empty((len(sizes), 2), dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().abr_sample_fg_bg(flat.data_ptr(), keys.data_ptr(), offsets.data_ptr(), len(sizes),
                                                   int(self.batch_size_per_image), int(self.batch_size_per_image * self.positive_fraction),
                                                   pos.data_ptr(), neg.data_ptr(), counts.data_ptr(), _lib.stream_ptr(dev)))
        return list(pos.split(sizes)), list(neg.split(sizes)), counts

    def __call__(self, matched_idxs, objectness=None):
        """matched_idxs: list of per-image label tensors (-1 ignored, 0 negative, > 0 positive).
        Returns two lists of per-image uint8 masks: the sampled positives and the sampled negatives."""
        if self.device_sampling:
            pos, neg, _ = self.sample_on_device(matched_idxs)
            return pos, neg
        pos_idx, neg_idx = [], []
        for per_image in matched_idxs:
            positive = torch.nonzero(per_image >= 1).squeeze(1)
            negative = torch.nonzero(per_image == 0).squeeze(1)
            num_pos = min(positive.numel(), int(self.batch_size_per_image * self.positive_fraction))
            num_neg = min(negative.numel(), self.batch_size_per_image - num_pos)
            perm_pos = torch.randperm(positive.numel(), device=positive.device)[:num_pos]
            perm_neg = torch.randperm(negative.numel(), device=negative.device)[:num_neg]
            pos_mask = torch.zeros_like(per_image, dtype=torch.uint8)
            neg_mask = torch.zeros_like(per_image, dtype=torch.uint8)
            pos_mask[positive[perm_pos]] = 1
            neg_mask[negative[perm_neg]] = 1
            pos_idx.append(pos_mask)
            neg_idx.append(neg_mask)
        return pos_idx, neg_idx
