#!/bin/bash
# quick iteration on the GPU box: harness (correctness [+ bench]), parity tests, bench line (+ A/B env toggles)
TAG=${1:-it}
mkdir -p gpurun_out
timeout 300 ./build/test_conv_gemm ${HARNESS_ARG} > gpurun_out/${TAG}_harness.log 2>&1; echo "harness exit $?"
grep -E "FAIL|correctness|bench" gpurun_out/${TAG}_harness.log | tail -60
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_gpu_tests.log 2>&1; echo "pytest exit $?"
tail -4 gpurun_out/${TAG}_gpu_tests.log
timeout 600 python bench.py --no-cpu-baseline --profile-out gpurun_out/${TAG}_per_kernel.json > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_bench.json"))
    print("BENCH value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "roof", d["roofline"]["frac"] if d.get("roofline") else None)
    for k,v in list(d["per_kernel"].items())[:14]: print("  ", k, v["launches"], v["ms"])
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/${TAG}_bench.err").read()[-2000:])
PY
for kv in $AB_ENVS; do
  env $kv timeout 600 python bench.py --no-cpu-baseline > gpurun_out/${TAG}_bench_${kv}.json 2>/dev/null
  python -c "import json;d=json.load(open('gpurun_out/${TAG}_bench_${kv}.json'));print('$kv', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])"
done
