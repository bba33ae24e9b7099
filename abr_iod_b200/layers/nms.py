"""``nms`` with the reference's signature (maskrcnn_benchmark/layers/nms.py:8 -> csrc/nms.h:10-28) plus the
batched form the RPN / post-processing loops need, both backed by ``abr_nms_batched`` of libabr_b200."""
import ctypes

import torch

from .. import _lib


def nms_batched(boxes_list, scores_list, nms_thresh, max_proposals=-1, cpu_tie_rule=False):
    """NMS of a whole batch of images in three launches and NO host synchronisation.

    Arguments:
        boxes_list (list[Tensor[n_i,4]]): xyxy boxes per image (CUDA)
        scores_list (list[Tensor[n_i]])
        nms_thresh (float), max_proposals (int): as in ``boxlist_nms``
    Returns:
        keep (LongTensor[n_images, stride]): per image, kept indices ascending, padded with -1
        n_keep (IntTensor[n_images]): valid entries per image (device tensor; reading it is the only sync)
    """
    n_images = len(boxes_list)
    assert n_images == len(scores_list) and n_images > 0
    device = boxes_list[0].device
    _lib.require_cuda(boxes_list[0], "boxes")
    sizes = [int(b.shape[0]) for b in boxes_list]
    boxes = torch.cat([b.detach().reshape(-1, 4) for b in boxes_list], 0).to(torch.float32).contiguous()
    scores = torch.cat([s.detach().reshape(-1) for s in scores_list], 0).to(torch.float32).contiguous()
    offsets = (ctypes.c_int * (n_images + 1))()
    for i, n in enumerate(sizes):
        offsets[i + 1] = offsets[i] + n
    nmax = max(sizes)
    stride = min(nmax, max_proposals) if max_proposals > 0 else nmax
    keep = torch.empty((n_images, stride), dtype=torch.int64, device=device)
    n_keep = torch.empty((n_images,), dtype=torch.int32, device=device)
    L = _lib.lib()
    ws_bytes = int(L.abr_nms_workspace_bytes(offsets, n_images))
    ws = torch.empty((max(ws_bytes, 1),), dtype=torch.uint8, device=device)
    with torch.cuda.device(device):
        _lib.check(L.abr_nms_batched(boxes.data_ptr(), scores.data_ptr(), offsets, n_images, float(nms_thresh),
                                     int(bool(cpu_tie_rule)), int(max_proposals), keep.data_ptr(), stride,
                                     n_keep.data_ptr(), ws.data_ptr(), ws_bytes, _lib.stream_ptr(device)))
    return keep, n_keep


def nms(boxes, scores, threshold):
    """``_C.nms(dets[N,4], scores[N], thr) -> LongTensor`` of kept original indices, ascending.
    Like the reference (csrc/nms.h:17-18) an empty CUDA input returns an empty CPU LongTensor."""
    _lib.require_cuda(boxes, "boxes")
    if boxes.numel() == 0:
        return torch.empty((0,), dtype=torch.int64, device="cpu")
    keep, n_keep = nms_batched([boxes], [scores], threshold)
    return keep[0, : int(n_keep.item())]
