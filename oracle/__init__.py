"""CPU oracle for the retrieval-scoring hot path — TEST INFRASTRUCTURE ONLY.

Restates, in numpy / torch-CPU, the reference's algorithm for the two scoring stages so the
CUDA kernels can be checked against it.  Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s cpu_baseline / `--impl reference` legs may import this package; the product
package (`automative-rag_b200/`) never does and fails loudly without its CUDA library.

Pinning status:
  * MaxSim  — PINNED: `oracle.maxsim` is asserted equal to the reference's own
    `ColBERTReranker._compute_maxsim_scores` (rerankers.py:215-265) executed in the build
    container; the inputs' seeds and the reference's outputs are committed under
    tests/golden/ by tests/golden/make_golden.py.
  * Rerank tail (sort / min-max / blend) — PINNED the same way against code lifted verbatim in
    behaviour from rerankers.py:298-343 run by make_golden.py through the reference module.
  * Filter (`_build_filter`) — structure pinned by the reference's own test
    (tests/test_retrieval.py:122-152), restated in tests/test_filters.py.
  * Dense top-k scores/ids — PARITY UNPINNED: the arithmetic lives in qdrant-client 1.13.3 /
    the Qdrant server (poetry.lock:5310-5311, docker-compose.yml:225), neither of which is
    present in /root/reference nor installable here; the reference's tests hold no golden
    vector for it.  `oracle.dense` restates qdrant-client local mode's published algorithm.
"""
