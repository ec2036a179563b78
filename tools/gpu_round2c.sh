#!/bin/bash
# Round-2 final evidence batch on one GPU: ncu launch list of exact-mode episodes and `ncu --set full` captures of the two 1x1
# kernels of the trunk at the 33-image shapes.   gpurun --timeout 1500 -- 'bash tools/gpu_round2c.sh'
set -u
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 130 -c 270 --csv --log-file gpurun_out/r02_launches_final.csv \
    python bench.py --steps 2 --warmup 1 --precision exact --no-variants --no-cpu-baseline > gpurun_out/r02_launches_final.log 2>&1
python tools/summarize_launches.py gpurun_out/r02_launches_final.csv > gpurun_out/r02_launch_shares_final.md 2>/dev/null; head -40 gpurun_out/r02_launch_shares_final.md
# single-CTA staged split kernel: second pass, res2 shortcut + the three res2 conv3 launches (33 images)
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k "regex:.*conv_gemm_f16_kernel<.int.128, .int.3, .int.2.*" -s 8 -c 4 -f -o gpurun_out/r02_prof_conv3_split_33 \
    python tools/profile_trunk.py 2 > gpurun_out/r02_ncu_conv3_split_33.log 2>&1
tail -n 2 gpurun_out/r02_ncu_conv3_split_33.log
# CTA-pair split kernel: second pass, res3.sc, res4.sc, res4 block 0 / 1 conv1 + conv3
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k "regex:.*conv1x1_pair_split_kernel.*" -s 21 -c 6 -f -o gpurun_out/r02_prof_pair_split_33 \
    python tools/profile_trunk.py 2 > gpurun_out/r02_ncu_pair_split_33.log 2>&1
tail -n 2 gpurun_out/r02_ncu_pair_split_33.log
ls -la gpurun_out/*_33.ncu-rep
