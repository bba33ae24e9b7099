"""``BoxList`` container with the interface of the reference's structures/bounding_box.py:9-257.  As INPUT every op
of this package only needs ``bbox``, ``size``, ``mode``, ``convert``, ``fields/get_field/add_field``, ``__getitem__``,
``__len__``, ``area`` -- the reference's own class works.  The fused ops also CONSTRUCT BoxLists (proposals,
detections); those go through ``make_boxlist`` below, which uses the reference's class once ``compat.patch_loaded()``
has seen it, so that downstream reference code (``prediction.resize`` in voc_eval.py:20 / engine/inference.py:110,
``copy_with_fields`` in the mask head, ...) gets the type it expects.  The class here carries the same geometric
methods for stand-alone use."""
import torch

FLIP_LEFT_RIGHT, FLIP_TOP_BOTTOM = 0, 1  # PIL.Image constants, as in the reference (bounding_box.py:5-6)

OUTPUT_CLASS = None  # set by compat.patch_loaded() to maskrcnn_benchmark.structures.bounding_box.BoxList


def make_boxlist(bbox, image_size, mode="xyxy"):
    """A new BoxList of the class callers downstream expect (see the module docstring)."""
    cls = OUTPUT_CLASS if OUTPUT_CLASS is not None else BoxList
    return cls(bbox, image_size, mode)


class BoxList(object):
    def __init__(self, bbox, image_size, mode="xyxy"):
        device = bbox.device if isinstance(bbox, torch.Tensor) else torch.device("cpu")
        bbox = torch.as_tensor(bbox, dtype=torch.float32, device=device)
        if bbox.ndimension() != 2 or bbox.size(-1) != 4:
            raise ValueError("bbox should be [N,4], got %s" % (tuple(bbox.shape),))
        if mode not in ("xyxy", "xywh"):
            raise ValueError("mode should be 'xyxy' or 'xywh'")
        self.bbox = bbox
        self.size = image_size  # (image_width, image_height)
        self.mode = mode
        self.extra_fields = {}

    def add_field(self, field, field_data):
        self.extra_fields[field] = field_data

    def get_field(self, field):
        return self.extra_fields[field]

    def has_field(self, field):
        return field in self.extra_fields

    def fields(self):
        return list(self.extra_fields.keys())

    def convert(self, mode):
        if mode not in ("xyxy", "xywh"):
            raise ValueError("mode should be 'xyxy' or 'xywh'")
        if mode == self.mode:
            return self
        b = self.bbox
        if mode == "xywh":  # from xyxy, +1 convention
            new = torch.stack((b[:, 0], b[:, 1], b[:, 2] - b[:, 0] + 1, b[:, 3] - b[:, 1] + 1), dim=1)
        else:               # from xywh
            new = torch.stack((b[:, 0], b[:, 1], b[:, 0] + (b[:, 2] - 1).clamp(min=0),
                               b[:, 1] + (b[:, 3] - 1).clamp(min=0)), dim=1)
        out = BoxList(new, self.size, mode=mode)
        out.extra_fields.update(self.extra_fields)
        return out

    def to(self, device):
        out = BoxList(self.bbox.to(device), self.size, self.mode)
        for k, v in self.extra_fields.items():
            out.add_field(k, v.to(device) if hasattr(v, "to") else v)
        return out

    def __getitem__(self, item):
        out = BoxList(self.bbox[item], self.size, self.mode)
        for k, v in self.extra_fields.items():
            out.add_field(k, v[item])
        return out

    def __len__(self):
        return self.bbox.shape[0]

    def area(self):
        b = self.bbox
        if self.mode == "xyxy":
            return (b[:, 2] - b[:, 0] + 1) * (b[:, 3] - b[:, 1] + 1)
        return b[:, 2] * b[:, 3]

    def _xyxy_columns(self):
        b = self.convert("xyxy").bbox
        return b[:, 0:1], b[:, 1:2], b[:, 2:3], b[:, 3:4]

    def _with_fields_from(self, other, method, *args, **kwargs):
        """Copies ``other``'s extra fields; non-tensor fields (masks, keypoints) are transformed by their own ``method``."""
        for k, v in other.extra_fields.items():
            if not isinstance(v, torch.Tensor):
                v = getattr(v, method)(*args, **kwargs)
            self.add_field(k, v)
        return self

    def resize(self, size, *args, **kwargs):
        """bounding_box.py:90-127: boxes scaled to an image of ``size`` = (width, height); one multiply when both ratios
        agree, per-axis otherwise (the result keeps this BoxList's mode)."""
        rw, rh = (float(s) / float(o) for s, o in zip(size, self.size))
        if rw == rh:
            return BoxList(self.bbox * rw, size, self.mode)._with_fields_from(self, "resize", size, *args, **kwargs)
        x0, y0, x1, y1 = self._xyxy_columns()
        out = BoxList(torch.cat((x0 * rw, y0 * rh, x1 * rw, y1 * rh), dim=-1), size, "xyxy")
        return out._with_fields_from(self, "resize", size, *args, **kwargs).convert(self.mode)

    def transpose(self, method):
        """bounding_box.py:129-165: horizontal flip (with the 1-pixel convention) or vertical flip (without, as there)."""
        if method not in (FLIP_LEFT_RIGHT, FLIP_TOP_BOTTOM):
            raise NotImplementedError("Only FLIP_LEFT_RIGHT and FLIP_TOP_BOTTOM implemented")
        w, h = self.size
        x0, y0, x1, y1 = self._xyxy_columns()
        if method == FLIP_LEFT_RIGHT:
            flipped = torch.cat((w - x1 - 1, y0, w - x0 - 1, y1), dim=-1)
        else:
            flipped = torch.cat((x0, h - y1, x1, h - y0), dim=-1)
        return BoxList(flipped, self.size, "xyxy")._with_fields_from(self, "transpose", method).convert(self.mode)

    def crop(self, box):
        """bounding_box.py:167-193: coordinates relative to the (left, upper, right, lower) window, clamped to it."""
        w, h = box[2] - box[0], box[3] - box[1]
        x0, y0, x1, y1 = self._xyxy_columns()
        cropped = torch.cat(((x0 - box[0]).clamp(min=0, max=w), (y0 - box[1]).clamp(min=0, max=h),
                             (x1 - box[0]).clamp(min=0, max=w), (y1 - box[1]).clamp(min=0, max=h)), dim=-1)
        return BoxList(cropped, (w, h), "xyxy")._with_fields_from(self, "crop", box).convert(self.mode)

    def clip_to_image(self, remove_empty=True):
        """bounding_box.py:214-225: in-place clamp of xyxy coordinates to [0, size-1]; optionally drops empty boxes."""
        w, h = self.size
        self.bbox[:, 0::2].clamp_(min=0, max=w - 1)
        self.bbox[:, 1::2].clamp_(min=0, max=h - 1)
        if remove_empty:
            b = self.bbox
            return self[(b[:, 3] > b[:, 1]) & (b[:, 2] > b[:, 0])]
        return self

    def copy_with_fields(self, fields, skip_missing=False):
        """bounding_box.py:239-249: same boxes, only the named extra fields."""
        out = BoxList(self.bbox, self.size, self.mode)
        for f in fields if isinstance(fields, (list, tuple)) else [fields]:
            if self.has_field(f):
                out.add_field(f, self.get_field(f))
            elif not skip_missing:
                raise KeyError("Field '%s' not found in %s" % (f, self))
        return out

    def __repr__(self):
        return "BoxList(num_boxes=%d, image_width=%s, image_height=%s, mode=%s)" % (
            len(self), self.size[0], self.size[1], self.mode)
