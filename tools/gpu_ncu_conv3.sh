#!/bin/bash
# `ncu --set full` capture of the staged bottleneck 1x1 kernel (shortcut + 3 x conv3 of res2, 8 images 800x1333)
TAG=${1:-r01}
ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k "regex:.*conv_gemm_f16_kernel<.int.256, .int.2, .int.2.*" -c 4 -f -o gpurun_out/${TAG}_prof_conv3 \
    python tools/profile_head.py 1 > gpurun_out/${TAG}_ncu_conv3.log 2>&1
tail -n 6 gpurun_out/${TAG}_ncu_conv3.log
ls -la gpurun_out/${TAG}_prof_conv3.ncu-rep
