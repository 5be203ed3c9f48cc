#!/bin/bash
mkdir -p gpurun_out
SDXE_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:attentionx -s 0 -c 1 -o gpurun_out/r27_attnx python tools/profile_unet.py --config sd15 --iters 1 > gpurun_out/r27_ncu.log 2>&1
echo "ncu rc=$?"
