#!/bin/bash
mkdir -p gpurun_out
SDXE_CLUSTER=0 SDXE_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 1 -c 2 -o gpurun_out/r15_conv_nocl python tools/profile_unet.py --config sd15 --iters 1 > gpurun_out/r15_ncu0.log 2>&1
echo "ncu0 rc=$?"
SDXE_CLUSTER=1 SDXE_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 1 -c 2 -o gpurun_out/r15_conv_cl python tools/profile_unet.py --config sd15 --iters 1 > gpurun_out/r15_ncu1.log 2>&1
echo "ncu1 rc=$?"
SDXE_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention2_kernel -s 0 -c 1 -o gpurun_out/r15_attn2 python tools/profile_unet.py --config sd15 --iters 1 > gpurun_out/r15_ncu2.log 2>&1
echo "ncu2 rc=$?"
