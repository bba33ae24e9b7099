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
