"""Packed Box-Rehearsal prototype store: ONE file (and one device upload) for the whole memory instead of the reference's
one JPEG per box (``Mem.creat_and_save_box_image``, tools/extract_memory.py:213-236, read back file by file in
``PascalVOCDataset_ABR.load_boxes_from_old_mem`` / ``_sample_per_bbox_from_boxrehearsal``, data/datasets/voc_abr.py:395-399,
:530).

Layout: ``names`` (the reference's ``"{class}_{index:05d}"`` stems, so the paste keeps parsing the class from the name),
``shapes`` [n,2] (h, w), ``offsets`` [n+1] (bytes into ``pixels``) and ``pixels`` -- the crops as raw HWC uint8 RGB, back
to back: exactly the resident pool ``BoxRehearsalPaster`` keeps on the GPU, so loading a store is one ``np.load`` and one
host-to-device copy.  The crops are stored losslessly (the reference's JPEG round trip alters pixel values; a store built
from its JPEG files reproduces them bit for bit, one built from the source images keeps the original pixels)."""
import os

import numpy as np


class PrototypeStore(object):
    def __init__(self, names, shapes, offsets, pixels):
        self.names = [str(n) for n in names]
        self.shapes = np.asarray(shapes, np.int32).reshape(-1, 2)
        self.offsets = np.asarray(offsets, np.int64)
        self.pixels = np.asarray(pixels, np.uint8)
        if len(self.names) != len(self.shapes) or len(self.offsets) != len(self.names) + 1 or self.offsets[-1] != self.pixels.size:
            raise ValueError("inconsistent prototype store")

    def __len__(self):
        return len(self.names)

    def crop(self, i):
        h, w = self.shapes[i]
        return self.pixels[self.offsets[i]: self.offsets[i + 1]].reshape(h, w, 3)

    def prototypes(self):
        """The ``(file_name, HWC uint8 array)`` list ``BoxRehearsalPaster`` takes."""
        return [(n if os.path.splitext(n)[1] else n + ".jpg", self.crop(i)) for i, n in enumerate(self.names)]

    def save(self, path):
        np.savez(path, names=np.asarray(self.names), shapes=self.shapes, offsets=self.offsets, pixels=self.pixels)

    @classmethod
    def load(cls, path):
        with np.load(path, allow_pickle=False) as z:
            return cls(z["names"], z["shapes"], z["offsets"], z["pixels"])

    @classmethod
    def pack(cls, prototypes):
        """``prototypes``: iterable of ``(name, image)`` with ``image`` a PIL image or an HWC uint8 array."""
        names, shapes, chunks = [], [], []
        for name, im in prototypes:
            a = np.ascontiguousarray(np.asarray(im.convert("RGB") if hasattr(im, "convert") else im), np.uint8)
            if a.ndim != 3 or a.shape[2] != 3:
                raise ValueError("prototype %s: expected an RGB image, got shape %s" % (name, a.shape))
            names.append(name)
            shapes.append(a.shape[:2])
            chunks.append(a.reshape(-1))
        offsets = np.zeros(len(chunks) + 1, np.int64)
        np.cumsum([c.size for c in chunks], out=offsets[1:])
        pixels = np.concatenate(chunks) if chunks else np.zeros(0, np.uint8)
        return cls(names, shapes, offsets, pixels)

    @classmethod
    def from_boxes(cls, picks):
        """What ``creat_and_save_box_image`` does per selected box (extract_memory.py:213-236), packed: ``picks`` is an
        iterable of ``(class_id, index_in_class, image HWC uint8 array or PIL image, box [x1,y1,x2,y2])``; the crop is
        ``image[int(y1):int(y2), int(x1):int(x2)]`` (PIL ``crop`` with the truncated coordinates) and is named
        ``"{class}_{index:05d}.jpg"``."""
        protos = []
        for cls_id, ind, image, box in picks:
            a = np.asarray(image.convert("RGB") if hasattr(image, "convert") else image)
            x1, y1, x2, y2 = (int(v) for v in box)
            protos.append(("%s_%05d.jpg" % (cls_id, ind), a[y1:y2, x1:x2]))
        return cls.pack(protos)
