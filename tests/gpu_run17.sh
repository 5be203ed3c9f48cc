#!/bin/bash
mkdir -p gpurun_out
SDXE_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 0 -c 7 -o gpurun_out/r17_gemm python tools/profile_unet.py --config sd15 --iters 1 > gpurun_out/r17_ncu0.log 2>&1
echo "ncu0 rc=$?"
SDXE_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention -s 0 -c 2 -o gpurun_out/r17_attn python tools/profile_unet.py --config sd15 --iters 1 > gpurun_out/r17_ncu1.log 2>&1
echo "ncu1 rc=$?"
