#!/bin/bash
# Run every GPU test file in its own process under a timeout (a hung kernel must not eat the box),
# then a short bench.  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
FILES=${@:-"tests/test_dense_gpu.py tests/test_merge_filter_gpu.py tests/test_maxsim_gpu.py tests/test_dropin_gpu.py tests/test_dense_batch_gpu.py"}
for f in $FILES; do
  name=$(basename $f .py)
  timeout ${TEST_TIMEOUT:-240} python -m pytest $f -q -m gpu --timeout 120 -x -q > gpurun_out/$name.log 2>&1
  echo "== $f exit $?"; tail -n 15 gpurun_out/$name.log
done
