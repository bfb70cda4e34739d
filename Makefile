# Convenience targets; everything is plain python underneath (see README.md).
PY ?= python

.PHONY: build test gpu-test smoke bench bench-ref c-client clean

build:            ## nvcc -gencode arch=compute_100a,code=sm_100a -> automative-rag_b200/lib/librag_b200.so
	$(PY) -c "import __graft_entry__ as g; g.build()"

test: build       ## CPU suite: oracle vs golden vectors, host logic, ABI surface, SASS content, gloo world-size 2
	$(PY) -m pytest tests -x -q -m "not gpu"

gpu-test: build   ## parity tests through the C ABI (needs a B200)
	$(PY) -m pytest tests -x -q -m gpu

smoke: build      ## one small pass of every stage on cuda:0, checked against the oracle
	$(PY) -c "import __graft_entry__ as g; g.smoke()"

bench: build      ## one JSON line: queries/s, e2e, roofline, cpu_baseline (N GPUs: torchrun ... bench.py --gpus N)
	$(PY) bench.py

bench-ref:        ## the reference arm: the CPU path on this box's host cores
	$(PY) bench.py --impl reference

c-client: build   ## the C ABI from plain C99
	gcc -std=c99 -Wall -Iinclude examples/c_client.c -o /tmp/c_client -Lautomative-rag_b200/lib -lrag_b200 \
	    -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,$(CURDIR)/automative-rag_b200/lib -Wl,-rpath,/usr/local/cuda/lib64
	/tmp/c_client

clean:
	rm -rf automative-rag_b200/build automative-rag_b200/lib
