"""``BalancedPositiveNegativeSampler`` with the reference's interface and RANDOM STREAM
(modeling/balanced_positive_negative_sampler.py:5-74): the same two ``torch.randperm`` calls per image, in the same order
and with the same sizes, so that a seeded run picks the same RoIs as the reference.  This is host-side glue around the
device matching (a few index tensors per image), not a kernel."""
import torch


class BalancedPositiveNegativeSampler(object):
    def __init__(self, batch_size_per_image, positive_fraction):
        self.batch_size_per_image = batch_size_per_image
        self.positive_fraction = positive_fraction

    def __call__(self, matched_idxs, objectness=None):
        """matched_idxs: list of per-image label tensors (-1 ignored, 0 negative, > 0 positive).
        Returns two lists of per-image uint8 masks: the sampled positives and the sampled negatives."""
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
