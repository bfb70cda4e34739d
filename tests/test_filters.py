"""Filter construction and host evaluation (CPU).  `test_build_filter` is the reference's own
behavioural pin (tests/test_retrieval.py:122-152), asserted on the product's `_build_filter`."""
import random

import numpy as np
import pytest

import automative_rag_b200 as rag
from automative_rag_b200.filters import INT_MISSING, UnsupportedFilter, compile_filter
from automative_rag_b200.vectorstore import payload_passes
from oracle import filters as ofilters


def test_build_filter():
    # Test single value filter
    single_filter = rag.build_filter({"manufacturer": "Toyota"})
    assert len(single_filter.must) == 1
    assert single_filter.must[0].key == "metadata.manufacturer"
    assert single_filter.must[0].match.value == "Toyota"
    # Test list value filter
    list_filter = rag.build_filter({"manufacturer": ["Toyota", "Honda"]})
    assert len(list_filter.must) == 1
    assert len(list_filter.must[0].should) == 2
    assert list_filter.must[0].should[0].key == "metadata.manufacturer"
    assert list_filter.must[0].should[0].match.value == "Toyota"
    assert list_filter.must[0].should[1].key == "metadata.manufacturer"
    assert list_filter.must[0].should[1].match.value == "Honda"
    # Test year range filter
    year_filter = rag.build_filter({"year": 2023})
    assert len(year_filter.must) == 1
    assert year_filter.must[0].key == "metadata.year"
    assert year_filter.must[0].range.gte == 2023
    assert year_filter.must[0].range.lte == 2023
    # Test multiple field filter
    multi_filter = rag.build_filter({"manufacturer": "Toyota", "category": "sedan", "year": 2023})
    assert len(multi_filter.must) == 3


def test_build_filter_skips_none_and_empty():
    f = rag.build_filter({"manufacturer": None, "model": [], "category": [None], "year": [2020, None, 2021]})
    assert len(f.must) == 1 and len(f.must[0].should) == 2  # vectorstore.py:235-249
    assert rag.build_filter({}).must == []
    # year given as a string is an equality match, not a range (vectorstore.py:256, :267)
    f = rag.build_filter({"year": "2023"})
    assert f.must[0].match.value == "2023" and f.must[0].range is None


def test_oracle_and_product_filter_structures_agree():
    flt = {"manufacturer": ["Toyota", None, "BMW"], "year": 2021, "category": "suv", "model": None, "source": []}
    o = ofilters.build_filter(flt)
    p = rag.build_filter(flt)
    assert len(o) == len(p.must) == 3
    assert o[0][0] == "should" and [c[2] for c in o[0][1]] == [c.match.value for c in p.must[0].should]
    assert o[1] == ("range", "metadata.year", 2021, 2021) and p.must[1].range.gte == 2021
    assert o[2] == ("match", "metadata.category", "suv") and p.must[2].match.value == "suv"


def _random_payload(rng):
    md = {}
    if rng.random() < 0.9:
        md["manufacturer"] = rng.choice(["Toyota", "Honda", "BMW", "Tesla"])
    if rng.random() < 0.8:
        md["year"] = rng.choice([2019, 2020, 2021, 2022, 2023, "2021", 2021.5])
    if rng.random() < 0.7:
        md["category"] = rng.choice(["sedan", "suv", None, ["suv", "truck"]])
    if rng.random() < 0.5:
        md["custom"] = rng.choice(["a", "b", 3])
    return {"page_content": "x", "metadata": md}


FILTERS = [
    {"manufacturer": "Toyota"}, {"manufacturer": ["Toyota", "BMW"]}, {"year": 2021}, {"year": 2021.0},
    {"year": "2021"}, {"year": [2020, 2023]}, {"manufacturer": "Tesla", "year": 2022, "category": "suv"},
    {"category": ["suv", "sedan"]}, {"custom": "a"}, {"custom": 3}, {"manufacturer": "Nobody"}, {"year": 2021.5},
    {"manufacturer": [None]}, {},
]


@pytest.mark.parametrize("flt", FILTERS)
def test_host_predicate_matches_oracle(flt):
    rng = random.Random(5)
    payloads = [_random_payload(rng) for _ in range(400)]
    want = ofilters.filter_mask(payloads, flt)
    f = rag.build_filter(flt)
    got = np.array([payload_passes(p, f) for p in payloads])
    assert (got == want).all()


def test_compile_filter_to_columns():
    dicts = {"manufacturer": {"Toyota": 0, "Honda": 1}, "category": {"suv": 0}}
    ints = ("year",)
    c = compile_filter(rag.build_filter({"manufacturer": ["Honda", "Toyota", "Kia"], "year": 2021}), dicts, ints)
    assert c == [("manufacturer", [0, 1]), ("year", [2021])]
    # values that can never match give an empty set (the clause rejects every row)
    assert compile_filter(rag.build_filter({"manufacturer": "Kia"}), dicts, ints) == [("manufacturer", [])]
    assert compile_filter(rag.build_filter({"year": 2021.5}), dicts, ints) == [("year", [])]
    assert compile_filter(rag.build_filter({"year": "2021"}), dicts, ints) == [("year", [])]
    assert compile_filter(rag.build_filter({"manufacturer": 7}), dicts, ints) == [("manufacturer", [])]
    with pytest.raises(UnsupportedFilter):
        compile_filter(rag.build_filter({"custom": "a"}), dicts, ints)
    assert INT_MISSING == -(2 ** 31)
