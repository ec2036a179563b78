#!/bin/bash
# GPU check of the training backward: parity tests, timing, memcheck of one small backward
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_training.py -q -s > gpurun_out/r02_bwd_tests.log 2>&1; echo "pytest exit $?"
grep -E "worst|max-norm|passed|failed|^E  " gpurun_out/r02_bwd_tests.log | tail -30
timeout 300 python tools/bench_training.py --out gpurun_out/r02_training_step.json > gpurun_out/r02_training_step.log 2>&1; echo "bench_training exit $?"
tail -3 gpurun_out/r02_training_step.log
if [ "$1" = "memcheck" ]; then
  timeout 420 compute-sanitizer --tool memcheck --log-file gpurun_out/r02_memcheck_backward.log python -m pytest tests/test_gpu_training.py -x -q -k "deterministic" > gpurun_out/r02_memcheck_backward.out 2>&1; echo "memcheck exit $?"
  tail -5 gpurun_out/r02_memcheck_backward.log
fi
