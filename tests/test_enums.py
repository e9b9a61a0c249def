"""CPU: Metric / ExpandMode keep the reference's names and integer values (annlite/enums.py:27-35)."""
import pickle

import pytest

from annlite_b200.enums import ExpandMode, Metric


def test_values_and_text_form():
    assert [(m.name, int(m)) for m in Metric] == [('EUCLIDEAN', 1), ('INNER_PRODUCT', 2), ('COSINE', 3)]
    assert [(m.name, int(m)) for m in ExpandMode] == [('STEP', 1), ('DOUBLE', 2), ('ADAPTIVE', 3)]
    assert str(Metric.COSINE) == 'COSINE' and Metric.from_string('inner_product') is Metric.INNER_PRODUCT
    assert Metric.coerce(3) is Metric.COSINE and Metric.coerce('Euclidean') is Metric.EUCLIDEAN
    assert [m.native for m in Metric] == [0, 1, 2]          # ANNB_METRIC_* of include/annb.h
    assert pickle.loads(pickle.dumps(ExpandMode.ADAPTIVE)) is ExpandMode.ADAPTIVE


def test_unknown_name_is_a_value_error():
    with pytest.raises(ValueError, match='not a valid enum'):
        Metric.from_string('manhattan')
