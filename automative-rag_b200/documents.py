"""`Document` as the reference uses it (langchain_core.documents.Document: `page_content` +
`metadata`; imported at vectorstore.py:7 and rerankers.py:7 of the reference).  langchain_core
is used when installed so callers can pass their own objects; otherwise a same-shaped class."""
from __future__ import annotations

from typing import Any, Dict, Optional

try:  # pragma: no cover - optional dependency
    from langchain_core.documents import Document  # type: ignore
except Exception:  # noqa: BLE001

    class Document:  # type: ignore[no-redef]
        __slots__ = ("page_content", "metadata")

        def __init__(self, page_content: str = "", metadata: Optional[Dict[str, Any]] = None, **_: Any):
            self.page_content = page_content
            self.metadata = metadata if metadata is not None else {}

        def __repr__(self) -> str:
            return f"Document(page_content={self.page_content!r}, metadata={self.metadata!r})"

        def __eq__(self, other: object) -> bool:
            return (
                isinstance(other, Document)
                and self.page_content == other.page_content
                and self.metadata == other.metadata
            )
