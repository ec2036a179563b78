#!/bin/bash
# Round-2 measurement batch on one GPU: GPU test-suite, bench (both precision modes + variants), the ncu launch list of an
# exact-mode episode and `ncu --set full` captures of the two roofline kernels in exact mode.
#   gpurun --timeout 1700 -- 'bash tools/gpu_round2b.sh [tests] [bench] [ncu]'
set -u
mkdir -p gpurun_out
what="${*:-tests bench ncu}"
if [[ "$what" == *tests* ]]; then
  timeout 1300 python -m pytest tests -m gpu -q 2>&1 | tail -25
fi
if [[ "$what" == *bench* ]]; then
  timeout 600 python bench.py --steps 20 --warmup 5 --profile-out gpurun_out/r02_per_kernel_exact.json > gpurun_out/r02_bench_b.json 2> gpurun_out/r02_bench_b.err
  tail -c 1500 gpurun_out/r02_bench_b.err
  python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_b.json").read().strip().splitlines()[-1])
for k in ("value", "ms_per_step", "e2e", "clocks", "cpu_baseline", "variants"):
    print(k, d.get(k))
print("fast", {k: d["fast_mode"][k] for k in ("value", "ms_per_step")}, d["fast_mode"]["e2e"]["value"])
print("roofline", {k: d["roofline"][k] for k in ("achieved", "frac", "share_of_step")}, "tensor", {k: d["roofline_tensor"][k] for k in ("achieved", "frac", "share_of_step")})
for k, v in list(d["per_kernel"].items())[:24]:
    print("   ", k, v)
PY
fi
if [[ "$what" == *ncu* ]]; then
  # launch list of exact-mode episodes (skip the first, cold one)
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 130 -c 270 --csv --log-file gpurun_out/r02_launches_exact.csv \
      python bench.py --steps 2 --warmup 1 --precision exact --no-variants --no-cpu-baseline > gpurun_out/r02_launches_exact.log 2>&1
  python tools/summarize_launches.py gpurun_out/r02_launches_exact.csv > gpurun_out/r02_launch_shares_exact.md 2>/dev/null; head -30 gpurun_out/r02_launch_shares_exact.md
  # the HBM-bound roofline kernel: staged split 1x1 convolution (res2 shortcut + 3 x conv3, 8 images)
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
      -k "regex:.*conv_gemm_f16_kernel<.int.128, .int.3, .int.2.*" -c 4 -f -o gpurun_out/r02_prof_conv3_split \
      python tools/profile_head.py 1 > gpurun_out/r02_ncu_conv3_split.log 2>&1
  tail -n 3 gpurun_out/r02_ncu_conv3_split.log
  # the tensor-bound roofline kernel: CTA-pair 3x3 convolution of the FCOS towers in split mode (after the 14 backbone launches)
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
      -k "regex:.*conv3x3_pair_kernel<.int.3, .int.8, .int.256.*" -s 14 -c 2 -f -o gpurun_out/r02_prof_tower_split \
      python tools/profile_head.py 1 > gpurun_out/r02_ncu_tower_split.log 2>&1
  tail -n 3 gpurun_out/r02_ncu_tower_split.log
  # separable ROIAlign at the class-sweep size
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
      -k "regex:.*roi_align_separable.*" -c 1 -f -o gpurun_out/r02_prof_roi_align \
      python tools/bench_sweep_sharded.py --steps 1 --warmup 0 > gpurun_out/r02_ncu_roi_align.log 2>&1
  tail -n 3 gpurun_out/r02_ncu_roi_align.log
  ls -la gpurun_out/*.ncu-rep
fi
