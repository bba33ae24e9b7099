"""numpy restatement of the Prototype Box Selection scoring -- TEST INFRASTRUCTURE ONLY.

  * descriptors            tools/prototype_box_selection.py:96-101  (torch.mean over channels on the CPU, then .tolist())
  * mean_feature_ranking   tools/extract_memory.py:111-147 (Mem.mean_feature_sampling: top-up, class mean, normalisation by
                           the Frobenius norm of all descriptors, distance, argsort) in float64 like the reference, which
                           feeds Python floats to numpy
Pinned by tests/golden/prototype.npz, produced by running the reference's Mem.mean_feature_sampling here."""
import numpy as np
import torch


def descriptors(roi_align_features):
    return torch.mean(torch.from_numpy(np.ascontiguousarray(roi_align_features, np.float32)), dim=1).numpy()


def mean_feature_ranking(features, num_bbox_per_cls):
    feats = [np.asarray(f, np.float64) for f in features]
    source = list(range(len(feats)))
    if len(feats) < num_bbox_per_cls:
        deficit = num_bbox_per_cls - len(feats)
        feats.extend(feats[:deficit])
        source.extend(source[:deficit])
    boxes_fea = np.array(feats)
    cls_mean = np.mean(boxes_fea, axis=0)
    cls_mean /= np.linalg.norm(cls_mean)
    phi = boxes_fea / np.linalg.norm(boxes_fea)
    dist = np.sqrt(np.sum((cls_mean - phi) ** 2, axis=tuple(range(1, phi.ndim))))
    order = np.argsort(dist, kind="stable")[:num_bbox_per_cls]
    return order, dist, np.asarray(source)


def herding_ranking(features, num_bbox_per_cls):
    """tools/extract_memory.py:163-197 (Mem.herding_feature_sampling), the selection loop as written there, float64.
    NB the reference's function cannot run as shipped (``_ind_bbox_per_cls`` is read before assignment at :203), so this
    restatement follows the source and is NOT pinned by a golden run: parity unpinned for this rule."""
    feats = [np.asarray(f, np.float64) for f in features]
    source = list(range(len(feats)))
    if len(feats) < num_bbox_per_cls:
        deficit = num_bbox_per_cls - len(feats)
        feats.extend(feats[:deficit])
        source.extend(source[:deficit])
    boxes_fea = np.array(feats)
    boxes_fea = np.reshape(boxes_fea, (boxes_fea.shape[0], -1))
    cls_mean = np.mean(boxes_fea, axis=0)
    cls_mean /= np.linalg.norm(cls_mean)
    current_center = cls_mean * 0
    selected = []
    for f in range(len(boxes_fea)):
        candidate_centers = current_center * f / (f + 1) + boxes_fea / (f + 1)
        distances = pow(candidate_centers - cls_mean, 2).sum(axis=1)
        distances[selected] = np.inf
        new_index = distances.argmin().tolist()
        selected.append(new_index)
        current_center = candidate_centers[new_index]
    return np.asarray(selected[:num_bbox_per_cls]), np.asarray(source)


def sample_fg_bg(matched_idxs, keys, batch_size_per_image, positive_fraction):
    """The counts of modeling/balanced_positive_negative_sampler.py:19-68 with the subset drawn by smallest key (ties by
    index): numpy restatement of abr_sample_fg_bg for one image.  Returns (pos mask, neg mask)."""
    m, keys = np.asarray(matched_idxs), np.asarray(keys, np.float32)
    positive, negative = np.nonzero(m >= 1)[0], np.nonzero(m == 0)[0]
    num_pos = min(len(positive), int(batch_size_per_image * positive_fraction))
    num_neg = min(len(negative), batch_size_per_image - num_pos)
    pos_mask, neg_mask = np.zeros(len(m), np.uint8), np.zeros(len(m), np.uint8)
    pos_mask[positive[np.argsort(keys[positive], kind="stable")[:num_pos]]] = 1
    neg_mask[negative[np.argsort(keys[negative], kind="stable")[:num_neg]]] = 1
    return pos_mask, neg_mask
