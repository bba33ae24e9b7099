"""``BoxCoder`` parameters (modeling/box_coder.py:8-21 of the reference).  The decode itself is fused into the device
kernels that consume it (``abr_rpn_proposals``); any object with ``weights`` and ``bbox_xform_clip`` -- including the
reference's own BoxCoder -- can be passed where a box coder is expected."""
import math


class BoxCoder(object):
    def __init__(self, weights, bbox_xform_clip=math.log(1000.0 / 16)):
        self.weights = tuple(float(w) for w in weights)
        self.bbox_xform_clip = bbox_xform_clip
