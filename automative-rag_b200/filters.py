"""Metadata filters: the reference's `_build_filter` (vectorstore.py:216-276) and their
compilation to the columnar clause form `rs_filter_mask` evaluates on the device.

`Filter`, `FieldCondition`, `MatchValue`, `Range` have the attribute layout of the
qdrant_client.http.models classes the reference builds (vectorstore.py:11), so code and tests
that inspect `filter.must[0].key`, `.match.value`, `.range.gte`, `.should` read the same.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Dict, List, Optional, Sequence, Tuple, Union


@dataclass
class MatchValue:
    value: Any


@dataclass
class Range:
    gte: Optional[float] = None
    lte: Optional[float] = None
    gt: Optional[float] = None
    lt: Optional[float] = None


@dataclass
class FieldCondition:
    key: str
    match: Optional[MatchValue] = None
    range: Optional[Range] = None


@dataclass
class Filter:
    must: Optional[List[Union[FieldCondition, "Filter"]]] = None
    should: Optional[List[Union[FieldCondition, "Filter"]]] = None


def build_filter(metadata_filter: Dict[str, Union[str, List[str], int, List[int]]]) -> Filter:
    """Same rules, same order, as QdrantStore._build_filter (vectorstore.py:216-276)."""
    must_conditions: List[Union[FieldCondition, Filter]] = []
    for key, value in metadata_filter.items():
        field_path = f"metadata.{key}"
        if isinstance(value, list):
            if not value:  # empty list: no constraint (:235-236)
                continue
            should_conditions = [
                FieldCondition(key=field_path, match=MatchValue(value=v)) for v in value if v is not None
            ]
            if should_conditions:
                must_conditions.append(Filter(should=should_conditions))
        elif isinstance(value, (int, float)) and key == "year":
            must_conditions.append(FieldCondition(key=field_path, range=Range(gte=value, lte=value)))
        elif value is not None:
            must_conditions.append(FieldCondition(key=field_path, match=MatchValue(value=value)))
    return Filter(must=must_conditions)


# ---------------------------------------------------------------------------------------------
# Compilation to columnar clauses.
#
# A collection stores each indexed payload field as one int32 column on the device:
#   keyword fields  -> dictionary code of the string (code -1 = missing / not a string)
#   integer fields  -> the value itself (INT_MISSING = missing / not an integer)
# A clause is (field name, [int32 values]); a row passes a clause when its column value is in
# the set, and passes the filter when it passes every clause.  A value that can never match
# (unknown string, non-integral year, type mismatch) contributes nothing to its set; an empty
# set makes the clause — and the filter — reject every row, as Qdrant would.

INT_MISSING = -(2**31)


class UnsupportedFilter(ValueError):
    """The filter touches a field the collection keeps no column for."""


def _int_value(v: Any) -> Optional[int]:
    if isinstance(v, bool):
        return None
    if isinstance(v, int):
        return v if -(2**31) < v < 2**31 else None
    if isinstance(v, float) and v.is_integer():
        return int(v)
    return None


def compile_filter(flt: Filter, keyword_dicts: Dict[str, Dict[str, int]], int_fields: Sequence[str]
                   ) -> List[Tuple[str, List[int]]]:
    """Filter -> [(field, value set)] for rs_filter_mask.  `keyword_dicts[field]` maps string -> code."""
    clauses: List[Tuple[str, List[int]]] = []

    def field_of(cond: FieldCondition) -> str:
        if not cond.key.startswith("metadata."):
            raise UnsupportedFilter(f"unindexed key {cond.key!r}")
        name = cond.key[len("metadata."):]
        if name not in keyword_dicts and name not in int_fields:
            raise UnsupportedFilter(f"no column for field {name!r}")
        return name

    def values_of(cond: FieldCondition, name: str) -> List[int]:
        if cond.match is not None:
            v = cond.match.value
            if name in int_fields:
                iv = None if isinstance(v, float) else _int_value(v)  # MatchValue is int/str/bool only
                return [] if iv is None else [iv]
            if not isinstance(v, str):
                return []
            code = keyword_dicts[name].get(v)
            return [] if code is None else [code]
        if cond.range is not None:
            r = cond.range
            if name not in int_fields:
                return []
            if r.gt is not None or r.lt is not None or r.gte is None or r.lte is None or r.gte != r.lte:
                raise UnsupportedFilter("only the reference's Range(gte=v, lte=v) point ranges are compiled")
            iv = _int_value(r.gte)
            return [] if iv is None else [iv]
        raise UnsupportedFilter("condition without match or range")

    for cond in flt.must or []:
        if isinstance(cond, Filter):
            if cond.must:
                raise UnsupportedFilter("nested must")
            name = None
            vals: List[int] = []
            for sub in cond.should or []:
                if not isinstance(sub, FieldCondition):
                    raise UnsupportedFilter("nested filter inside should")
                n = field_of(sub)
                if name is None:
                    name = n
                elif n != name:
                    raise UnsupportedFilter("should over several fields")
                vals.extend(values_of(sub, n))
            if name is not None:
                clauses.append((name, sorted(set(vals))))
        else:
            name = field_of(cond)
            clauses.append((name, values_of(cond, name)))
    if flt.should:
        raise UnsupportedFilter("top-level should")
    return clauses


def pack_bits(bits) -> "np.ndarray":
    """bool [n] -> int32 words [ceil(n/32)], LSB-first: the mask layout rs_dense_topk reads."""
    import numpy as np

    bits = np.asarray(bits, dtype=np.uint8)
    n = bits.shape[0]
    padded = np.zeros((n + 31) // 32 * 32, dtype=np.uint8)
    padded[:n] = bits
    return np.packbits(padded.reshape(-1, 32), axis=1, bitorder="little").view("<u4").reshape(-1).view(np.int32).copy()
