"""``boxlist_nms`` with the reference's signature (structures/boxlist_ops.py:9-31) and its batched form."""
import torch

from ..layers.nms import nms as _box_nms
from ..layers.nms import nms_batched as _box_nms_batched


def boxlist_nms(boxlist, nms_thresh, max_proposals=-1, score_field="scores"):
    """Non-maximum suppression on a BoxList; scores come from ``score_field``.  If ``max_proposals > 0`` only the
    first ``max_proposals`` kept boxes (in ascending index order, like the reference) survive."""
    if nms_thresh <= 0:
        return boxlist
    mode = boxlist.mode
    boxlist = boxlist.convert("xyxy")
    keep = _box_nms(boxlist.bbox, boxlist.get_field(score_field), nms_thresh)
    if max_proposals > 0:
        keep = keep[:max_proposals]
    boxlist = boxlist[keep.to(boxlist.bbox.device)]
    return boxlist.convert(mode)


def boxlist_nms_batched(boxlists, nms_thresh, max_proposals=-1, score_field="scores"):
    """``[boxlist_nms(b, ...) for b in boxlists]`` with ONE device pass and one host sync for the whole batch
    (replaces the per-image loop of modeling/rpn/inference.py:111-117)."""
    if nms_thresh <= 0 or len(boxlists) == 0:
        return list(boxlists)
    modes = [b.mode for b in boxlists]
    xyxy = [b.convert("xyxy") for b in boxlists]
    nonempty = [i for i, b in enumerate(xyxy) if len(b) > 0]
    out = list(xyxy)
    if nonempty:
        keep, n_keep = _box_nms_batched([xyxy[i].bbox for i in nonempty],
                                        [xyxy[i].get_field(score_field) for i in nonempty], nms_thresh, max_proposals)
        counts = n_keep.tolist()  # the only synchronisation of the batch
        for j, i in enumerate(nonempty):
            out[i] = xyxy[i][keep[j, : counts[j]]]
    return [b.convert(m) for b, m in zip(out, modes)]


def cat_boxlist(bboxes):
    """One BoxList out of several of the same image (same size, mode and field names): structures/boxlist_ops.py:102-128.
    A single-element list is returned without a copy, like the reference's ``_cat``."""
    bboxes = list(bboxes)
    if not bboxes:
        raise ValueError("cat_boxlist needs at least one BoxList")
    first = bboxes[0]
    names = set(first.fields())
    for b in bboxes[1:]:
        if b.size != first.size or b.mode != first.mode or set(b.fields()) != names:
            raise ValueError("cat_boxlist: image size, mode and fields must agree")
    join = (lambda ts: ts[0]) if len(bboxes) == 1 else (lambda ts: torch.cat(ts, dim=0))
    from .bounding_box import make_boxlist

    out = make_boxlist(join([b.bbox for b in bboxes]), first.size, first.mode)
    for name in names:
        out.add_field(name, join([b.get_field(name) for b in bboxes]))
    return out


def boxlist_iou(boxlist1, boxlist2):
    """IoU of two BoxLists of the same image, [N,M] (structures/boxlist_ops.py:53-88; +1 pixel convention), one kernel."""
    from .. import _lib

    if boxlist1.size != boxlist2.size:
        raise RuntimeError("boxlists should have same image size, got {}, {}".format(boxlist1, boxlist2))
    b1 = boxlist1.convert("xyxy").bbox.detach().to(torch.float32).contiguous()
    b2 = boxlist2.convert("xyxy").bbox.detach().to(torch.float32).contiguous()
    _lib.require_cuda(b1, "boxlist1")
    out = torch.empty((b1.shape[0], b2.shape[0]), dtype=torch.float32, device=b1.device)
    with torch.cuda.device(b1.device):
        _lib.check(_lib.lib().abr_box_iou(b1.data_ptr(), b1.shape[0], b2.data_ptr(), b2.shape[0], out.data_ptr(),
                                          _lib.stream_ptr(b1.device)))
    return out
