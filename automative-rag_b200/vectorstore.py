"""B200VectorStore — drop-in for the reference's QdrantStore on the dense-retrieval path.

Mirrors src/core/query/retrieval/vectorstore.py of the reference: same constructor shape
(`client, collection_name, embedding_function`, :25-30; embedding function mandatory, :39-40),
same `similarity_search_with_score(query, k=5, metadata_filter=None)` (:166-214) including the
"filtered search failed -> log and retry unfiltered" behaviour (:199-207), same `_build_filter`
(:216-276), `add_documents` (:124-164), `search_by_metadata` (:278-316), `delete_by_ids`
(:318-353), `get_stats` (:355-388), `get_embedding` (:390-411) and `repair_indices` (:412-470).

What changes is what sits underneath.  The reference forwards to a Qdrant server; here the
collection lives in one GPU's HBM:
  * vectors   [capacity, d] fp16 row-major, unit-normalised in fp32 on insert (Qdrant does the same
              for Distance.COSINE, :52-57,:75-81), plus fp32 1/|row| of the ROUNDED row so the
              kernel returns the exact cosine of what is stored;
  * payload   one int32 column per indexed field (the nine fields of `_create_payload_indexes`,
              :89-122): keyword fields dictionary-encoded, integer fields raw;
  * deletes   a tombstone bitmask — the same bit-packed form as the filter mask.
A search is: embed the query (kept: `embedding_function.embed_query`) -> rs_filter_mask builds
the row bitmask on the device -> rs_dense_topk_host scans the corpus (TMA-staged exact scan with
the mask fused, top-k in-kernel) -> k (score, id) pairs come back.  No CPU scoring path exists.
"""
from __future__ import annotations

import json
import logging
import os
import time
import uuid
from typing import Any, Dict, List, Optional, Sequence, Tuple, Union

import torch

from . import _ffi
from .documents import Document
from .filters import (INT_MISSING, FieldCondition, Filter, UnsupportedFilter, build_filter, compile_filter,
                      pack_bits)

logger = logging.getLogger(__name__)

# Fields indexed by the reference (vectorstore.py:93-103) and their schema (:110-113).
KEYWORD_FIELDS = ("manufacturer", "model", "category", "engine_type", "transmission", "source", "source_id")
INTEGER_FIELDS = ("year", "ingestion_time")
MAX_K = 2048  # rs_dense_topk: 1 <= k <= 2048


class Collection:
    """One named collection resident on one GPU."""

    def __init__(self, name: str, dim: int, engine: _ffi.Engine, dtype: torch.dtype = torch.float16,
                 distance: str = "Cosine", capacity: int = 1024):
        if dim % 8 != 0:
            raise ValueError(f"embedding dimension must be a multiple of 8 (got {dim})")
        self.name, self.dim, self.engine, self.dtype, self.distance = name, dim, engine, dtype, distance
        self.device = engine.device
        self.n = 0
        self.capacity = 0
        self.vectors = torch.empty(0, dim, dtype=dtype, device=self.device)
        self.inv_norm = torch.empty(0, dtype=torch.float32, device=self.device)
        self.columns: Dict[str, torch.Tensor] = {f: torch.empty(0, dtype=torch.int32, device=self.device)
                                                 for f in KEYWORD_FIELDS + INTEGER_FIELDS}
        self.keyword_dicts: Dict[str, Dict[str, int]] = {f: {} for f in KEYWORD_FIELDS}
        self.int_fields: List[str] = list(INTEGER_FIELDS)   # + integer columns created lazily for un-indexed keys
        # fields that hold a value the int32 column cannot express (a list-valued keyword, which Qdrant matches on any
        # element; a numeric string ...): filters on them are evaluated on the host payloads instead
        self.unencodable: Dict[str, bool] = {}
        self.tombstone = torch.empty(0, dtype=torch.int32, device=self.device)
        self.ids: List[str] = []
        self.id_to_row: Dict[str, int] = {}
        self.payloads: List[Dict[str, Any]] = []
        self.deleted = 0
        self._reserve(capacity)

    # -- storage ----------------------------------------------------------------------------
    def _reserve(self, want: int) -> None:
        if want <= self.capacity:
            return
        cap = max(want, 2 * self.capacity, 1024)
        cap = (cap + 31) // 32 * 32

        def grow(old: torch.Tensor, shape, fill=None):
            new = torch.empty(shape, dtype=old.dtype, device=self.device)
            if fill is not None:
                new.fill_(fill)
            if old.numel():
                new[: old.shape[0]] = old
            return new

        self.vectors = grow(self.vectors, (cap, self.dim))
        self.inv_norm = grow(self.inv_norm, (cap,))
        for f in self.columns:
            self.columns[f] = grow(self.columns[f], (cap,), fill=INT_MISSING)
        self.tombstone = grow(self.tombstone, (cap // 32,), fill=0)
        self.capacity = cap

    def _encode_value(self, f: str, v: Any) -> int:
        """int32 code of payload value `v` of field `f`.  A value the column cannot express marks the field, so that
        filters on it fall back to the host evaluator (the compiled path must never disagree with it)."""
        if f in self.keyword_dicts:
            if isinstance(v, str):
                d = self.keyword_dicts[f]
                return d.setdefault(v, len(d))
            if v is not None:
                self.unencodable[f] = True  # list-valued keyword, number in a keyword field ...
            return -1
        if isinstance(v, bool) or v is None:  # a bool never matches an integer condition (payload_passes)
            return INT_MISSING
        if isinstance(v, int) and -(2**31) < v < 2**31:
            return int(v)
        if isinstance(v, float) and v.is_integer() and abs(v) < 2**31:
            return int(v)  # 2020.0 matches MatchValue(2020) and Range(2020, 2020) on the host path too
        self.unencodable[f] = True
        return INT_MISSING

    def _encode_fields(self, metadata: Dict[str, Any]) -> Dict[str, int]:
        return {f: self._encode_value(f, metadata.get(f)) for f in self.columns}

    def ensure_column(self, field: str) -> bool:
        """Create (once) the device column of a payload key the reference does not index — `custom_filters` of the
        API (query_models.py:22-28) reach the store as ordinary metadata keys.  The type is taken from the stored
        values: all strings -> keyword (dictionary-encoded), all integers -> integer.  Returns False when the values
        are mixed or neither, in which case filters on the key stay on the host evaluator."""
        if field in self.columns:
            return True
        vals = [(p.get("metadata", {}) or {}).get(field) for p in self.payloads]
        present = [v for v in vals if v is not None]
        if present and all(isinstance(v, str) for v in present):
            self.keyword_dicts[field] = {}
        elif present and all(isinstance(v, int) and not isinstance(v, bool) for v in present):
            self.int_fields.append(field)
        else:
            return False
        col = torch.full((self.capacity,), INT_MISSING, dtype=torch.int32, device=self.device)
        self.columns[field] = col
        if vals:
            col[: self.n] = torch.tensor([self._encode_value(field, v) for v in vals], dtype=torch.int32, device=self.device)
        return True

    def upsert(self, ids: Sequence[str], vectors: torch.Tensor, payloads: Sequence[Dict[str, Any]]) -> None:
        """vectors: float32 [m, d] on any device."""
        m = len(ids)
        if m == 0:
            return
        for pid in ids:  # Qdrant upsert replaces a point with the same id
            if pid in self.id_to_row:
                self.delete([pid])
        self._reserve(self.n + m)
        v = vectors.to(self.device, torch.float32)
        if self.distance == "Cosine":
            v = v / v.norm(dim=1, keepdim=True).clamp_min(1e-30)
        stored = v.to(self.dtype)
        self.vectors[self.n: self.n + m] = stored
        self.inv_norm[self.n: self.n + m] = 1.0 / stored.float().norm(dim=1).clamp_min(1e-30)
        cols = {f: [] for f in self.columns}
        for p in payloads:
            enc = self._encode_fields(p.get("metadata", {}) or {})
            for f, val in enc.items():
                cols[f].append(val)
        for f, vals in cols.items():
            self.columns[f][self.n: self.n + m] = torch.tensor(vals, dtype=torch.int32, device=self.device)
        for i, pid in enumerate(ids):
            self.id_to_row[pid] = self.n + i
        self.ids.extend(ids)
        self.payloads.extend(payloads)
        self.n += m

    def rebuild_columns(self) -> List[str]:
        """Re-encode every stored payload into fresh payload columns and keyword dictionaries — the counterpart of
        dropping and recreating Qdrant's payload indexes (vectorstore.py:412-470).  Returns the rebuilt fields."""
        self.keyword_dicts = {f: {} for f in self.keyword_dicts}
        self.unencodable = {}
        cols: Dict[str, List[int]] = {f: [] for f in self.columns}
        for p in self.payloads:
            for f, val in self._encode_fields(p.get("metadata", {}) or {}).items():
                cols[f].append(val)
        for f, vals in cols.items():
            col = torch.full((self.capacity,), INT_MISSING, dtype=torch.int32, device=self.device)
            if vals:
                col[: self.n] = torch.tensor(vals, dtype=torch.int32, device=self.device)
            self.columns[f] = col
        return list(self.columns)

    def delete(self, ids: Sequence[str]) -> int:
        """Tombstone the rows of `ids` (unknown ids are ignored, as Qdrant does)."""
        rows = [self.id_to_row.pop(i) for i in ids if i in self.id_to_row]
        if not rows:
            return 0
        words: Dict[int, int] = {}
        for x in rows:
            words[x // 32] = words.get(x // 32, 0) | (1 << (x % 32))
        idx = torch.tensor(list(words.keys()), dtype=torch.int64, device=self.device)
        # int32 view of the uint32 bit patterns (bit 31 set -> negative)
        vals = torch.tensor([v - (1 << 32) if v >= (1 << 31) else v for v in words.values()],
                            dtype=torch.int32, device=self.device)
        self.tombstone[idx] = self.tombstone[idx] | vals
        self.deleted += len(rows)
        return len(rows)

    # -- masks ------------------------------------------------------------------------------
    def device_mask(self, flt: Optional[Filter]) -> Optional[torch.Tensor]:
        """Bit-packed pass mask over rows [0, n) for `flt` AND not-deleted; None = all rows pass."""
        if flt is None and self.deleted == 0:
            return None
        tomb = self.tombstone[: (self.n + 31) // 32] if self.deleted else None
        if flt is None:
            return self.engine.filter_mask([], [], self.n, tombstone=tomb)
        for key in _filter_keys(flt):  # un-indexed payload keys get their column on first use
            if key.startswith("metadata."):
                self.ensure_column(key[len("metadata."):])
        try:
            clauses = compile_filter(flt, self.keyword_dicts, self.int_fields)
            if any(self.unencodable.get(name) for name, _ in clauses):
                raise UnsupportedFilter("a filtered field holds values the device column cannot express")
        except UnsupportedFilter as e:
            logger.warning(f"filter evaluated on the host payloads ({e})")
            return self._host_mask(flt, tomb)
        cols = [self.columns[name][: self.n] for name, _ in clauses]
        sets = [vals for _, vals in clauses]
        return self.engine.filter_mask(cols, sets, self.n, tombstone=tomb)

    def matching_rows(self, flt: Optional[Filter], limit: int) -> List[int]:
        """Row indices (insertion order) of the first `limit` live rows that pass `flt`: the mask is built on the
        device, its set bits are listed on the device, `limit` integers come back."""
        if self.n == 0 or limit <= 0:
            return []
        mask = self.device_mask(flt)
        if mask is None:
            return list(range(min(self.n, limit)))
        bits = (mask.view(-1, 1) >> torch.arange(32, device=mask.device, dtype=torch.int32)) & 1
        rows = torch.nonzero(bits.view(-1)[: self.n], as_tuple=False).view(-1)[:limit]
        return rows.tolist()

    # -- persistence (the reference keeps its vectors in Qdrant's storage volume, docker-compose.yml:229-230) -----
    def save(self, path: str) -> None:
        """Write the collection to directory `path`: vectors (row-major, as stored), 1/|row|, the int32 payload
        columns, tombstone words — one device-to-host copy per array (in 1 GiB pieces) — plus ids, payloads and the
        keyword dictionaries as JSON."""
        os.makedirs(path, exist_ok=True)

        def dump(t: torch.Tensor, name: str) -> None:
            flat = t.contiguous().view(-1)
            step = max(1, (1 << 30) // max(flat.element_size(), 1))
            with open(os.path.join(path, name), "wb") as f:
                for a in range(0, flat.numel(), step):
                    f.write(flat[a: a + step].cpu().view(torch.uint8).numpy().tobytes())

        n = self.n
        dump(self.vectors[:n], "vectors.bin")
        dump(self.inv_norm[:n], "inv_norm.bin")
        dump(self.tombstone[: (n + 31) // 32], "tombstone.bin")
        for f, col in self.columns.items():
            dump(col[:n], f"column.{f}.bin")
        meta = {"format": 1, "name": self.name, "dim": self.dim, "dtype": str(self.dtype).replace("torch.", ""),
                "distance": self.distance, "n": n, "deleted": self.deleted, "fields": list(self.columns),
                "int_fields": self.int_fields, "keyword_dicts": self.keyword_dicts, "unencodable": self.unencodable,
                "ids": self.ids}
        with open(os.path.join(path, "payloads.json"), "w") as f:
            json.dump(self.payloads, f)
        with open(os.path.join(path, "meta.json"), "w") as f:  # written last: a directory without it is incomplete
            json.dump(meta, f)

    @classmethod
    def load(cls, path: str, engine: _ffi.Engine) -> "Collection":
        """Rebuild a collection saved by `save` on `engine`'s GPU: one host-to-device copy per array."""
        import numpy as np

        with open(os.path.join(path, "meta.json")) as f:
            meta = json.load(f)
        if meta.get("format") != 1:
            raise ValueError(f"unknown collection format {meta.get('format')!r} in {path}")
        dtype = getattr(torch, meta["dtype"])
        n = meta["n"]
        col = cls(meta["name"], meta["dim"], engine, dtype=dtype, distance=meta["distance"], capacity=max(n, 1))
        col.int_fields = list(meta["int_fields"])
        col.keyword_dicts = {f: dict(d) for f, d in meta["keyword_dicts"].items()}
        col.unencodable = dict(meta.get("unencodable", {}))

        def read(name: str, dt: torch.dtype, shape) -> torch.Tensor:
            raw = np.fromfile(os.path.join(path, name), dtype=np.uint8)
            return torch.from_numpy(raw).view(dt).view(shape).to(col.device)

        col.vectors[:n] = read("vectors.bin", dtype, (n, meta["dim"]))
        col.inv_norm[:n] = read("inv_norm.bin", torch.float32, (n,))
        col.tombstone[: (n + 31) // 32] = read("tombstone.bin", torch.int32, ((n + 31) // 32,))
        for f in meta["fields"]:
            if f not in col.columns:
                col.columns[f] = torch.full((col.capacity,), INT_MISSING, dtype=torch.int32, device=col.device)
            col.columns[f][:n] = read(f"column.{f}.bin", torch.int32, (n,))
        with open(os.path.join(path, "payloads.json")) as f:
            col.payloads = json.load(f)
        col.ids = list(meta["ids"])
        col.n, col.deleted = n, meta["deleted"]
        tomb = col.tombstone[: (n + 31) // 32].cpu().numpy().view(np.uint32) if col.deleted else None
        for row, pid in enumerate(col.ids):  # a re-upserted id maps to its newest, live row
            if tomb is None or not (int(tomb[row >> 5]) >> (row & 31)) & 1:
                col.id_to_row[pid] = row
        return col

    def _host_mask(self, flt: Filter, tomb: Optional[torch.Tensor]) -> torch.Tensor:
        """Filters on un-indexed payload keys: evaluate on the host payloads, upload the bits."""
        import numpy as np

        bits = np.fromiter((payload_passes(p, flt) for p in self.payloads[: self.n]), dtype=np.uint8, count=self.n)
        mask = torch.from_numpy(pack_bits(bits)).to(self.device)
        if tomb is not None:
            mask &= ~tomb
        return mask


def _lookup(payload: Dict[str, Any], dotted: str) -> Any:
    cur: Any = payload
    for part in dotted.split("."):
        if not isinstance(cur, dict) or part not in cur:
            return None
        cur = cur[part]
    return cur


def _filter_keys(flt: Union[Filter, FieldCondition]) -> List[str]:
    if isinstance(flt, FieldCondition):
        return [flt.key]
    out: List[str] = []
    for c in (flt.must or []) + (flt.should or []):
        out.extend(_filter_keys(c))
    return out


def payload_passes(payload: Dict[str, Any], flt: Union[Filter, FieldCondition]) -> bool:
    """Host evaluation of a filter on one payload (used for un-indexed keys and search_by_metadata)."""
    if isinstance(flt, FieldCondition):
        value = _lookup(payload, flt.key)
        values = value if isinstance(value, list) else [value]
        if flt.match is not None:
            want = flt.match.value
            for v in values:
                if v is None or isinstance(v, bool) != isinstance(want, bool) or isinstance(v, str) != isinstance(want, str):
                    continue
                if v == want:
                    return True
            return False
        if flt.range is not None:
            r = flt.range
            for v in values:
                if isinstance(v, bool) or not isinstance(v, (int, float)):
                    continue
                if ((r.gte is None or v >= r.gte) and (r.lte is None or v <= r.lte)
                        and (r.gt is None or v > r.gt) and (r.lt is None or v < r.lt)):
                    return True
            return False
        return False
    ok = all(payload_passes(payload, c) for c in (flt.must or []))
    if flt.should:
        ok = ok and any(payload_passes(payload, c) for c in flt.should)
    return ok


class B200Client:
    """Holds the named collections of one GPU (the role QdrantClient plays for QdrantStore)."""

    def __init__(self, device: Union[int, str, torch.device] = 0):
        self.engine = _ffi.get_engine(device)
        self.collections: Dict[str, Collection] = {}

    def collection_exists(self, name: str) -> bool:
        return name in self.collections

    def create_collection(self, collection_name: str, size: int, distance: str = "Cosine",
                          dtype: torch.dtype = torch.float16) -> Collection:
        if collection_name in self.collections:
            raise ValueError(f"collection {collection_name!r} already exists")
        col = Collection(collection_name, size, self.engine, dtype=dtype, distance=distance)
        self.collections[collection_name] = col
        return col

    def get_collection(self, name: str) -> Collection:
        return self.collections[name]

    def delete_collection(self, name: str) -> None:
        self.collections.pop(name, None)

    def save(self, path: str) -> None:
        """Persist every collection under directory `path` (one sub-directory each)."""
        os.makedirs(path, exist_ok=True)
        for i, (name, col) in enumerate(self.collections.items()):
            col.save(os.path.join(path, f"collection-{i}"))

    def load(self, path: str) -> List[str]:
        """Load every collection found under `path`; returns their names."""
        names = []
        for sub in sorted(os.listdir(path)):
            d = os.path.join(path, sub)
            if os.path.exists(os.path.join(d, "meta.json")):
                col = Collection.load(d, self.engine)
                self.collections[col.name] = col
                names.append(col.name)
        return names


class B200VectorStore:
    """Same public surface as the reference's QdrantStore (vectorstore.py:17)."""

    def __init__(self, client: B200Client, collection_name: str, embedding_function: Any):
        if embedding_function is None:  # vectorstore.py:39-40
            raise ValueError("Embedding function is required. No more metadata-only mode.")
        self.client = client
        self.collection_name = collection_name
        self.embedding_function = embedding_function
        logger.info(f"Initializing B200VectorStore with collection: {collection_name}")
        self._ensure_collection()

    # vectorstore.py:60-87
    def _ensure_collection(self) -> None:
        if not self.client.collection_exists(self.collection_name):
            sample_embedding = self.embedding_function.embed_query("sample text")
            self.client.create_collection(self.collection_name, size=len(sample_embedding), distance="Cosine")
            logger.info(f"Collection '{self.collection_name}' created with {len(sample_embedding)} dimensions")
        else:
            logger.info(f"Collection '{self.collection_name}' already exists")

    @property
    def collection(self) -> Collection:
        return self.client.get_collection(self.collection_name)

    # vectorstore.py:124-164
    def add_documents(self, documents: List[Document]) -> List[str]:
        if not documents:
            logger.warning("No documents provided to add_documents")
            return []
        current_time = time.time()
        for doc in documents:
            if "ingestion_time" not in doc.metadata:
                doc.metadata["ingestion_time"] = current_time
        doc_ids = []
        for doc in documents:
            if "id" not in doc.metadata or not doc.metadata["id"]:
                doc.metadata["id"] = f"doc-{str(time.time())}-{len(doc_ids)}"
            doc_ids.append(doc.metadata["id"])
        try:
            texts = [d.page_content for d in documents]
            vectors = torch.as_tensor(self.embedding_function.embed_documents(texts), dtype=torch.float32)
            point_ids = [uuid.uuid4().hex for _ in documents]  # langchain_qdrant generates uuid4 point ids
            payloads = [{"page_content": d.page_content, "metadata": d.metadata} for d in documents]
            self.collection.upsert(point_ids, vectors, payloads)
            logger.info(f"Collection now has {self.collection.n - self.collection.deleted} vectors")
            return point_ids
        except Exception as e:
            logger.error(f"Error adding documents to vector store: {str(e)}")
            raise

    # vectorstore.py:166-214
    def similarity_search_with_score(
            self,
            query: str,
            k: int = 5,
            metadata_filter: Optional[Dict[str, Union[str, List[str], int, List[int]]]] = None,
    ) -> List[Tuple[Document, float]]:
        logger.info(f"Performing similarity search for query: '{query}' with k={k}")
        if metadata_filter:
            filter_obj = self._build_filter(metadata_filter)
            try:
                results = self._search(query, k, filter_obj)
                logger.info(f"Search returned {len(results)} results with filter")
                return results
            except Exception as e:
                logger.error(f"Error in filtered search: {str(e)}, falling back to unfiltered search")
                results = self._search(query, k, None)
                logger.info(f"Fallback search returned {len(results)} results")
                return results
        results = self._search(query, k, None)
        logger.info(f"Search returned {len(results)} results without filter")
        return results

    def _search(self, query: str, k: int, flt: Optional[Filter]) -> List[Tuple[Document, float]]:
        col = self.collection
        if col.n == 0 or k <= 0:
            return []
        qvec = torch.as_tensor(self.embedding_function.embed_query(query), dtype=torch.float32)
        if qvec.numel() != col.dim:
            raise ValueError(f"query embedding has {qvec.numel()} dims, collection has {col.dim}")
        mask = col.device_mask(flt)
        if k > MAX_K:  # never truncate silently (the reference has no such limit; deployed k is 20-40)
            logger.warning(f"similarity search asked for k={k}; the scan kernel returns at most {MAX_K} results")
        scores, ids = col.engine.dense_topk_host(
            col.vectors[: col.n], qvec.to(col.dtype).contiguous(), min(k, MAX_K), mask_dev=mask,
            inv_norm=col.inv_norm[: col.n], metric=_ffi.RS_METRIC_COSINE)
        out: List[Tuple[Document, float]] = []
        for s, i in zip(scores[0].tolist(), ids[0].tolist()):
            if i < 0:
                break
            p = col.payloads[i]
            out.append((Document(page_content=p.get("page_content", ""), metadata=p.get("metadata", {})), float(s)))
        return out

    # vectorstore.py:216-276
    def _build_filter(self, metadata_filter: Dict[str, Union[str, List[str], int, List[int]]]) -> Filter:
        return build_filter(metadata_filter)

    # vectorstore.py:278-316
    def search_by_metadata(self, metadata_filter: Dict[str, Any], limit: int = 100) -> List[Document]:
        filter_obj = self._build_filter(metadata_filter)
        try:
            col = self.collection
            documents: List[Document] = []
            for row in col.matching_rows(filter_obj, limit):  # mask and row list built on the device; insertion order
                p = col.payloads[row]
                documents.append(Document(page_content=p.get("page_content", ""), metadata=p.get("metadata", {})))
            return documents
        except Exception as e:
            logger.error(f"Error in metadata search: {str(e)}")
            return []

    # vectorstore.py:318-353
    def delete_by_ids(self, ids: List[str]) -> None:
        if not ids:
            logger.warning("No document IDs provided for deletion")
            return
        if not self.client.collection_exists(self.collection_name):
            logger.error(f"Collection '{self.collection_name}' does not exist")
            return
        n = self.collection.delete(ids)
        logger.info(f"Successfully deleted {n} documents")

    # vectorstore.py:355-388
    def get_stats(self) -> Dict[str, Any]:
        try:
            col = self.collection
            live = col.n - col.deleted
            return {
                "name": self.collection_name,
                "vectors_count": live,
                "points_count": live,
                "status": "green",
                "config": {"params": {"vectors": {"size": col.dim, "distance": col.distance}}},
                "payload_indices": {f"metadata.{f}": "keyword" for f in KEYWORD_FIELDS}
                                   | {f"metadata.{f}": "integer" for f in INTEGER_FIELDS},
                "device": str(col.device),
                "hbm_bytes": int(col.n * col.dim * col.vectors.element_size()),
            }
        except Exception as e:
            logger.error(f"Error getting collection stats: {str(e)}")
            return {"name": self.collection_name, "error": str(e)}

    # vectorstore.py:412-470
    def repair_indices(self) -> Dict[str, Any]:
        """Rebuild the payload columns from the stored payloads; same result shape as the reference."""
        logger.info(f"Attempting to repair indices for collection '{self.collection_name}'")
        results: Dict[str, Any] = {"recreated_indices": [], "errors": [], "success": False}
        try:
            results["recreated_indices"] = [f"metadata.{f}" for f in self.collection.rebuild_columns()]
            results["success"] = True
            logger.info("Repair completed successfully")
        except Exception as e:
            error_msg = f"Error during repair: {str(e)}"
            logger.error(error_msg)
            results["errors"].append(error_msg)
        return results

    def save(self, path: str) -> None:
        """Persist this store's collection (vectors, payload columns, tombstones, payloads) to directory `path`."""
        self.collection.save(path)

    def load(self, path: str) -> None:
        """Replace this store's collection by the one saved at `path`."""
        col = Collection.load(path, self.client.engine)
        col.name = self.collection_name
        self.client.collections[self.collection_name] = col

    # vectorstore.py:390-411
    def get_embedding(self, id: str) -> Optional[List[float]]:
        col = self.collection
        row = col.id_to_row.get(id)
        if row is None:
            return None
        return col.vectors[row].float().tolist()
