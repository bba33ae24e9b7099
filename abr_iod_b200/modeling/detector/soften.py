"""The teacher's proposal pick of ``GeneralizedRCNN.generate_soften_proposal`` (modeling/detector/generalized_rcnn.py:
121-163 of the reference): per image, sort the RPN proposals by objectness, take the 128 best and draw 64 of them with
``random.sample`` -- the proposals whose pooled features feed the ARD loss.

The reference assembles the selection with one ``torch.cat`` per box (128 tiny kernels and allocations per image); here
it is one index gather per image.  The ``random.sample`` calls are the same, in the same order and with the same
arguments, so a seeded run picks the same proposals.  Host-side glue: no kernel of its own."""
import random

import torch

from ...structures.bounding_box import BoxList, make_boxlist


def select_soften_proposals(all_proposals, top_n=128, keep_n=64):
    """all_proposals: list[BoxList] with an ``objectness`` field.  Returns list[BoxList] (boxes + objectness) in the
    order ``random.sample`` drew them.  Equal objectness values keep their proposal order (the reference's unstable
    ``sort(descending=True)`` leaves it unspecified)."""
    selected = []
    for proposals in all_proposals:
        scores = proposals.get_field("objectness")
        order = torch.sort(scores, descending=True, stable=True)[1]
        num = len(proposals)
        if num < keep_n:  # generalized_rcnn.py:139-141
            picks = random.sample(range(0, num, 1), num)
        elif num < top_n:  # :142-144
            picks = random.sample(range(0, num, 1), keep_n)
        else:  # :145-147
            picks = random.sample(range(0, top_n, 1), keep_n)
        idx = order[torch.as_tensor(picks, dtype=torch.long, device=order.device)]
        out = make_boxlist(proposals.bbox[idx].reshape(-1, 4), proposals.size, proposals.mode)
        out.add_field("objectness", scores[idx].reshape(-1))
        selected.append(out)
    return selected
