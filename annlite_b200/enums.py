"""The two small vocabularies the path's public calls accept, value-compatible with the reference
(annlite/enums.py: ``Metric`` 1..3, ``ExpandMode`` 1..3) so that pickled configs and integer metric ids mean the
same thing on both sides.  ``Metric.native`` is the C-ABI metric id of ``include/annb.h``."""
import enum


class _Named(enum.IntEnum):
    """IntEnum whose text form is the bare member name and that can be looked up case-insensitively."""

    __str__ = lambda self: self.name  # noqa: E731

    @classmethod
    def from_string(cls, text):
        member = cls.__members__.get(str(text).strip().upper())
        if member is None:
            raise ValueError('%s is not a valid enum for %r, must be one of %s' % (str(text).upper(), cls, list(cls)))
        return member

    @classmethod
    def coerce(cls, value):
        """Accept a member, its name or its integer value."""
        if isinstance(value, cls):
            return value
        if isinstance(value, str):
            return cls.from_string(value)
        return cls(int(value))


Metric = _Named('Metric', ['EUCLIDEAN', 'INNER_PRODUCT', 'COSINE'], start=1, module=__name__)
ExpandMode = _Named('ExpandMode', ['STEP', 'DOUBLE', 'ADAPTIVE'], start=1, module=__name__)

# ANNB_METRIC_L2 / ANNB_METRIC_IP / ANNB_METRIC_COSINE of include/annb.h
Metric.native = property(lambda self: {'EUCLIDEAN': 0, 'INNER_PRODUCT': 1, 'COSINE': 2}[self.name])
