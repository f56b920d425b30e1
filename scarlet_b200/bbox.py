"""Integer bounding boxes in model-frame coordinates (host geometry only).

Mirrors the interface of the reference's ``scarlet/bbox.py`` (``Box`` 4-276, ``overlapped_slices`` 279-301);
the device consumes boxes as (shape, origin) integers in the source descriptors.
"""
import numpy as np


class Box:
    """A box of ``shape`` whose minimum corner sits at ``origin``.  2-D boxes are (H, W), 3-D (C, H, W)."""

    def __init__(self, shape, origin=None):
        self.shape = tuple(int(s) for s in shape)
        origin = (0,) * len(self.shape) if origin is None else tuple(int(o) for o in origin)
        if len(origin) != len(self.shape):
            raise ValueError("origin and shape need the same number of dimensions")
        self.origin = origin

    # -- constructors ---------------------------------------------------------------------------
    @staticmethod
    def from_bounds(*bounds):
        return Box([max(0, int(hi) - int(lo)) for lo, hi in bounds], origin=[int(lo) for lo, _ in bounds])

    @staticmethod
    def from_data(X, min_value=0):
        mask = np.asarray(X) > min_value
        if not mask.any():
            return Box.from_bounds(*([(0, 0)] * mask.ndim))
        idx = np.nonzero(mask)
        return Box.from_bounds(*[(int(i.min()), int(i.max()) + 1) for i in idx])

    # -- properties -------------------------------------------------------------------------------
    @property
    def D(self):
        return len(self.shape)

    @property
    def start(self):
        return self.origin

    @property
    def stop(self):
        return tuple(o + s for o, s in zip(self.origin, self.shape))

    @property
    def center(self):
        return tuple(o + s / 2 for o, s in zip(self.origin, self.shape))

    @property
    def bounds(self):
        return tuple((o, o + s) for o, s in zip(self.origin, self.shape))

    @property
    def slices(self):
        return tuple(slice(o, o + s) for o, s in zip(self.origin, self.shape))

    # -- queries / array helpers ---------------------------------------------------------------------
    def contains(self, p):
        if len(p) != self.D:
            raise ValueError("Dimension mismatch in %r and %d" % (p, self.D))
        return all(lo <= x < hi for x, (lo, hi) in zip(p, self.bounds))

    def extract_from(self, image, sub=None):
        if sub is None:
            sub = np.zeros(self.shape, dtype=image.dtype)
        im_sl, sub_sl = overlapped_slices(Box(image.shape), self)
        sub[sub_sl] = image[im_sl]
        return sub

    def insert_into(self, image, sub):
        im_sl, sub_sl = overlapped_slices(Box(image.shape), self)
        image[im_sl] = sub[sub_sl]
        return image

    def grow(self, radius):
        r = list(radius) if hasattr(radius, "__iter__") else [radius] * self.D
        return Box([s + 2 * q for s, q in zip(self.shape, r)], origin=[o - q for o, q in zip(self.origin, r)])

    # -- algebra ----------------------------------------------------------------------------------------
    def _same_dim(self, other):
        if other.D != self.D:
            raise ValueError("Dimension mismatch in the boxes %r and %r" % (other, self))

    def __or__(self, other):
        self._same_dim(other)
        return Box.from_bounds(*[(min(a0, b0), max(a1, b1)) for (a0, a1), (b0, b1) in zip(self.bounds, other.bounds)])

    def __and__(self, other):
        self._same_dim(other)
        return Box.from_bounds(*[(max(a0, b0), min(a1, b1)) for (a0, a1), (b0, b1) in zip(self.bounds, other.bounds)])

    def __getitem__(self, i):
        s, o = self.shape[i], self.origin[i]
        if not hasattr(s, "__iter__"):
            s, o = (s,), (o,)
        return Box(s, origin=o)

    def _offset(self, offset, sign):
        off = tuple(offset) if hasattr(offset, "__iter__") else (offset,) * self.D
        return tuple(int(a + sign * b) for a, b in zip(self.origin, off))

    def __iadd__(self, offset):
        self.origin = self._offset(offset, +1)
        return self

    def __add__(self, offset):
        return Box(self.shape, origin=self._offset(offset, +1))

    def __isub__(self, offset):
        self.origin = self._offset(offset, -1)
        return self

    def __sub__(self, offset):
        return Box(self.shape, origin=self._offset(offset, -1))

    def __matmul__(self, other):
        return Box.from_bounds(*(self.bounds + other.bounds))

    __imatmul__ = __matmul__

    def copy(self):
        return Box(self.shape, origin=self.origin)

    __copy__ = copy

    def __eq__(self, other):
        return isinstance(other, Box) and self.shape == other.shape and self.origin == other.origin

    def __hash__(self):
        return hash((self.shape, self.origin))

    def __repr__(self):
        return "<Box shape={0}, origin={1}>".format(self.shape, self.origin)


def overlapped_slices(bbox1, bbox2):
    """Slices into arrays bounded by ``bbox1`` and ``bbox2`` that address their common region."""
    ov = bbox1 & bbox2
    return (ov - bbox1.origin).slices, (ov - bbox2.origin).slices
