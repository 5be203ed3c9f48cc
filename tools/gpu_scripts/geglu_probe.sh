#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/bench_geglu.py probe 2>&1 | tail -14
timeout 300 python tools/profile_unet.py --config sd15 --iters 3 --profile > gpurun_out/unet_prof.txt 2>&1; tail -8 gpurun_out/unet_prof.txt
SDXE_PROFILE_DUMP=gpurun_out/ops_sd15.csv timeout 300 python tools/profile_unet.py --config sd15 --iters 1 --profile > /dev/null 2>&1
python tools/analyze_ops.py gpurun_out/ops_sd15.csv 2>/dev/null | head -24
