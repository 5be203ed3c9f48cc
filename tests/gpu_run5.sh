#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/ops_*.csv
SDXE_GEMM_STAGED=0 SDXE_PROFILE_DUMP=gpurun_out/ops_direct.csv python tools/profile_unet.py --config sd15 --iters 1 --profile > gpurun_out/r5_direct.log 2>&1
SDXE_GEMM_STAGED=1 SDXE_PROFILE_DUMP=gpurun_out/ops_staged.csv python tools/profile_unet.py --config sd15 --iters 1 --profile > gpurun_out/r5_staged.log 2>&1
tail -7 gpurun_out/r5_direct.log; tail -7 gpurun_out/r5_staged.log
SDXE_NO_GRAPH=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 30 -c 12 -o gpurun_out/r5_gemm python tools/profile_unet.py --config sd15 --iters 1 > gpurun_out/r5_ncu_gemm.log 2>&1
echo "ncu gemm rc=$?"
SDXE_NO_GRAPH=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_kernel -s 0 -c 3 -o gpurun_out/r5_attn python tools/profile_unet.py --config sd15 --iters 1 > gpurun_out/r5_ncu_attn.log 2>&1
echo "ncu attn rc=$?"
SDXE_NO_GRAPH=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gemm_kernel|attention_kernel|gn_|layer_norm|skinny|im2col|upsample|nhwc|timestep|cast_" -c 600 --csv --log-file gpurun_out/r5_launches.csv python tools/profile_unet.py --config sd15 --iters 1 > gpurun_out/r5_ncu_list.log 2>&1
echo "ncu list rc=$?"
ls -la gpurun_out | tail -12
