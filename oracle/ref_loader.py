"""Load the reference's own ColBERTReranker._compute_maxsim_scores for pinning the oracle.

TEST INFRASTRUCTURE ONLY.  This module is used by tests/golden/make_golden.py (run in the
build container, where /root/reference exists) to produce the committed golden vectors.
It is never imported by the product package, by `-m gpu` tests, by smoke() or by bench.py:
/root/reference does not exist on the GPU box.

The reference file (src/core/query/llm/rerankers.py) imports three modules that are not
installed here (langchain_core.documents :7, sentence_transformers :10, src.config.settings :12).
`_compute_maxsim_scores` (rerankers.py:215-265) touches none of them: it uses only
`self.batch_size`, `self.amp_enabled` and torch.  We therefore stub the three modules in
sys.modules, load the file by path, and call the function unbound.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types
import warnings
from types import SimpleNamespace

REFERENCE_ROOT = os.environ.get("RAG_REFERENCE_ROOT", "/root/reference")
_RERANKERS = os.path.join(REFERENCE_ROOT, "src/core/query/llm/rerankers.py")


def reference_available() -> bool:
    return os.path.isfile(_RERANKERS)


def _stub(name: str, **attrs) -> None:
    if name in sys.modules:
        return
    mod = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(mod, k, v)
    sys.modules[name] = mod


def load_reference_rerankers():
    """Return the reference `rerankers` module object (unmodified source, stubbed imports)."""
    if not reference_available():
        raise FileNotFoundError(f"reference not present at {_RERANKERS}")

    class _Document:  # shape of langchain_core.documents.Document as the reference uses it
        def __init__(self, page_content: str = "", metadata=None):
            self.page_content = page_content
            self.metadata = metadata or {}

    _stub("langchain_core")
    _stub("langchain_core.documents", Document=_Document)
    _stub("sentence_transformers", CrossEncoder=object)
    _stub("src")
    _stub("src.config")
    _stub("src.config.settings", settings=SimpleNamespace())
    spec = importlib.util.spec_from_file_location("_reference_rerankers", _RERANKERS)
    mod = importlib.util.module_from_spec(spec)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        spec.loader.exec_module(mod)
    return mod


def reference_maxsim_scores(query_embeddings, doc_embeddings_list, batch_size: int = 16):
    """Call the reference's ColBERTReranker._compute_maxsim_scores (rerankers.py:215-265) on CPU."""
    mod = load_reference_rerankers()
    self_ = SimpleNamespace(batch_size=batch_size, amp_enabled=False)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")  # torch.cuda.amp.autocast FutureWarning
        return mod.ColBERTReranker._compute_maxsim_scores(self_, query_embeddings, doc_embeddings_list)
