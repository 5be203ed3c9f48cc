#!/bin/bash
mkdir -p gpurun_out
K='regex:^(gemm_kernel|attention_kernel|attention2_kernel|attentionx_kernel|gn_.*|layer_norm.*|skinny_linear.*|im2col.*|upsample2x.*|nhwc_to_nchw.*|cast_.*|timestep_emb.*|denoiser_in.*|cfg_combine.*|euler_a.*|dpmpp.*|post_quant.*)$'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s 1200 -c 1500 --csv --log-file gpurun_out/r32_bench_launches.csv python bench.py --steps 1 --warmup 1 --no-extras > gpurun_out/r32_ncu_bench.log 2>&1
echo "ncu bench rc=$? lines=$(wc -l < gpurun_out/r32_bench_launches.csv)"
SDXE_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention2_kernel -s 0 -c 1 -o gpurun_out/r32_attn2 python tools/profile_unet.py --config sd15 --iters 1 > gpurun_out/r32_ncu1.log 2>&1
echo "ncu attn2 rc=$?"
SDXE_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 1 -c 3 -o gpurun_out/r32_gemm python tools/profile_unet.py --config sd15 --iters 1 > gpurun_out/r32_ncu2.log 2>&1
echo "ncu gemm rc=$?"
