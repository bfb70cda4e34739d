"""CPU restatement of the metadata filter — TEST INFRASTRUCTURE ONLY.

`build_filter` follows QdrantStore._build_filter
(/root/reference/src/core/query/retrieval/vectorstore.py:216-276) clause by clause:
  * AND across keys (`Filter(must=[...])`, :276); field path `metadata.<key>` (:231)
  * list value -> nested OR (`Filter(should=[MatchValue...])`), `None` items skipped (:239-240),
    empty list skipped entirely (:235-236), list of only-None skipped (:249)
  * key == "year" with an int/float -> `Range(gte=v, lte=v)` (:256-266)
  * any other non-None scalar -> `MatchValue` equality (:267-274); `None` skipped.
`row_passes` evaluates the predicate on one payload dict with the semantics of qdrant-client
local mode's `payload_filters.py` (third-party, absent here — restated): a dotted key walks
nested dicts; `MatchValue` is `==` (any element when the payload value is a list); `Range`
applies to numeric payload values only; a missing key never matches.

The structure is pinned by the reference's own test (tests/test_retrieval.py:122-152), mirrored
in tests/test_filters.py.  Never imported by the product package.
"""
from __future__ import annotations

from typing import Any, Dict, List, Tuple

import numpy as np

# a clause is ("match", path, value) | ("range", path, gte, lte) | ("should", [match clauses])
Clause = Tuple


def build_filter(metadata_filter: Dict[str, Any]) -> List[Clause]:
    must: List[Clause] = []
    for key, value in metadata_filter.items():
        path = f"metadata.{key}"
        if isinstance(value, list):
            if not value:
                continue
            should = [("match", path, v) for v in value if v is not None]
            if should:
                must.append(("should", should))
        elif isinstance(value, (int, float)) and key == "year":
            must.append(("range", path, value, value))
        elif value is not None:
            must.append(("match", path, value))
    return must


def _lookup(payload: Dict[str, Any], path: str):
    cur: Any = payload
    for part in path.split("."):
        if not isinstance(cur, dict) or part not in cur:
            return None
        cur = cur[part]
    return cur


def _match(value: Any, want: Any) -> bool:
    if value is None:
        return False
    if isinstance(value, list):
        return any(_match(v, want) for v in value)
    # qdrant keeps keyword / integer / bool matches apart: no cross-type equality
    if isinstance(want, bool) or isinstance(value, bool):
        return isinstance(want, bool) and isinstance(value, bool) and value == want
    if isinstance(want, str) != isinstance(value, str):
        return False
    return value == want


def _in_range(value: Any, gte, lte) -> bool:
    if isinstance(value, list):
        return any(_in_range(v, gte, lte) for v in value)
    if isinstance(value, bool) or not isinstance(value, (int, float)):
        return False
    return gte <= value <= lte


def clause_passes(payload: Dict[str, Any], clause: Clause) -> bool:
    kind = clause[0]
    if kind == "match":
        return _match(_lookup(payload, clause[1]), clause[2])
    if kind == "range":
        return _in_range(_lookup(payload, clause[1]), clause[2], clause[3])
    if kind == "should":
        return any(clause_passes(payload, c) for c in clause[1])
    raise ValueError(kind)


def row_passes(payload: Dict[str, Any], must: List[Clause]) -> bool:
    return all(clause_passes(payload, c) for c in must)


def filter_mask(payloads: List[Dict[str, Any]], metadata_filter: Dict[str, Any],
                deleted: np.ndarray | None = None) -> np.ndarray:
    """bool [n]: row i passes the filter and is not deleted."""
    must = build_filter(metadata_filter)
    out = np.array([row_passes(p, must) for p in payloads], dtype=bool)
    if deleted is not None:
        out &= ~np.asarray(deleted, dtype=bool)
    return out
