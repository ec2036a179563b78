#!/bin/bash
# First GPU call(s) of the next round: everything that was built after round 1's GPU budget ended.
#   gpurun --timeout 900 -- 'bash tools/gpu_round2.sh one'          (1 GPU)
#   gpurun --gpus 2 --timeout 600 -- 'bash tools/gpu_round2.sh two' (2 GPUs)
set -u
mkdir -p gpurun_out
case "${1:-one}" in
one)
  # 1. opt-in experiment parity (chunked trunk)
  SYLPH_RUN_UNVERIFIED=1 timeout 300 python -m pytest tests/test_gpu_zz_experiments.py tests/test_gpu_zexchange.py -x -q 2>&1 | tail -15
  # 1b. the resident-weight pair kernel in the kernel harness: bit-check against the CPU reference, then the 33-image shapes
  timeout 300 ./build/test_conv_gemm case PAIR_BRES 2>&1 | tail -8
  timeout 300 ./build/test_conv_gemm pairnarrow 2>&1 | tee gpurun_out/r02_pairnarrow_bres.log | tail -12
  # 2. A/B of the L2-chunked trunk on the headline config (value only; 20 steps each)
  for cfg in "|0|0" "1,2,4,0|0|0" "2,4,8,0|0|0" "1,1,2,4|0|0" "1,2,4,0|1|0" "1,2,4,0|2|0" "2,2,4,0|2|0" "3,6,0,0|2|0" "1,2,4,0|0|64" "2,4,8,0|0|96"; do
    ch="${cfg%%|*}"; rest="${cfg#*|}"; il="${rest%%|*}"; pm="${rest##*|}"
    echo "== SYLPH_TRUNK_CHUNK='$ch' SYLPH_TRUNK_INTERLEAVE=$il SYLPH_L2_PERSIST_MB=$pm" | tee -a gpurun_out/r02_trunk_chunk_ab.log
    SYLPH_TRUNK_CHUNK="$ch" SYLPH_TRUNK_INTERLEAVE="$il" SYLPH_L2_PERSIST_MB="$pm" timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 \
      | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print({k: d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['clocks'], {k: v['ms'] for k, v in list(d['per_kernel'].items())[:8]})" \
      | tee -a gpurun_out/r02_trunk_chunk_ab.log
  done
  # 2b. res2 / res3 conv2 on the CTA-pair kernel with resident weights
  for pb in 0 1; do
    echo "== SYLPH_PAIR_BRES=$pb" | tee -a gpurun_out/r02_trunk_chunk_ab.log
    SYLPH_PAIR_BRES=$pb timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 \
      | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print({k: d[k] for k in ('value','ms_per_step')}, d['clocks'], d['per_kernel'].get('res.conv2_3x3'))" \
      | tee -a gpurun_out/r02_trunk_chunk_ab.log
  done
  # 2c. FPN laterals on the staged kernels (1 = CTA pair, 2 = single CTA), against the fused default
  for lm in 0 1 2; do
    echo "== SYLPH_LATERAL=$lm" | tee -a gpurun_out/r02_trunk_chunk_ab.log
    SYLPH_LATERAL=$lm timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 \
      | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print({k: d[k] for k in ('value','ms_per_step')}, d['clocks'], d['per_kernel'].get('fpn.lateral1x1'), d['per_kernel'].get('fpn.upsample_add'))" \
      | tee -a gpurun_out/r02_trunk_chunk_ab.log
  done
  # 3. configs[4] (LVIS 1203-class sweep) on one GPU with both exchange forms (world 1: the peer form is the two kernels alone)
  timeout 300 python tools/bench_sweep_sharded.py --out gpurun_out/r02_cfg5_CodeGenerator_n1.json 2>&1 | tail -2
  ;;
two)
  timeout 300 python -m pytest tests/test_gpu_dist.py tests/test_gpu_zexchange.py -x -q 2>&1 | tail -15
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    tools/bench_sharded.py --steps 10 --out gpurun_out/r02_cfg4_sharded_n2.json 2>&1 | tail -3
  for gen in CodeGenerator ROIEncoder; do
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
      tools/bench_sweep_sharded.py --generator $gen --out gpurun_out/r02_cfg5_${gen}_n2.json 2>&1 | tail -2
  done
  ;;
ncu)
  # DRAM bytes of the trunk's GEMM launches with and without the L2-resident schedule.  One-pass metrics and
  # --cache-control none: ncu must neither flush L2 between kernels nor replay them, or the residency under test is gone.
  for cfg in "|0" "1,2,4,0|0" "1,2,4,0|2"; do
    ch="${cfg%%|*}"; il="${cfg##*|}"; tag="chunk_${ch//,/_}_il${il}"
    SYLPH_TRUNK_CHUNK="$ch" SYLPH_TRUNK_INTERLEAVE="$il" timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
      --cache-control none --clock-control none --print-units base --kernel-name-base demangled -k "regex:conv_gemm|conv1x1_pair|conv3x3_pair" -c 900 --csv \
      --log-file gpurun_out/r02_trunk_dram_${tag}.csv python tools/profile_head.py 1 > gpurun_out/r02_trunk_dram_${tag}.log 2>&1
    python - "$tag" <<'PY'
import csv, sys
tag = sys.argv[1]
rows = [r for r in csv.reader(open(f"gpurun_out/r02_trunk_dram_{tag}.csv", errors="ignore")) if len(r) > 10]
hdr = rows[0]
i_name, i_metric, i_val, i_id = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
agg, launches = {}, set()
for r in rows[1:]:
    agg[r[i_metric]] = agg.get(r[i_metric], 0.0) + float(r[i_val].replace(",", ""))
    launches.add(r[i_id])
print(tag, "GEMM launches", len(launches), {k: round(v / 1e9, 3) for k, v in agg.items()}, "(bytes -> GB, ns -> s)")
PY
  done
  ;;
esac
