#!/bin/bash
# One gpurun call: GPU parity tests, the bench line, the ncu launch list and two `--set full` captures.
# Usage (from the repo root on the GPU box): bash tools/gpu_check.sh <tag> [full]
TAG=${1:-r01}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_gpu_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_gpu_tests.log
tail -5 gpurun_out/${TAG}_gpu_tests.log
python bench.py --profile-out gpurun_out/${TAG}_per_kernel.json > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_bench.json | cut -c1-1500
if [ "$2" = "full" ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:conv3x3_pair -s 14 -c 2 -f -o gpurun_out/${TAG}_prof_pair \
      python tools/profile_head.py 1 > gpurun_out/${TAG}_ncu_pair.log 2>&1
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:conv_gemm_f16_kernel<256, 2, 2" -c 4 -f \
      -o gpurun_out/${TAG}_prof_conv3 python tools/profile_head.py 1 > gpurun_out/${TAG}_ncu_conv3.log 2>&1
  ls -la gpurun_out/
fi
