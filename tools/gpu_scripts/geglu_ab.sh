#!/bin/bash
# A/B of the GEGLU epilogue arithmetic (folded-constant GELU vs the textbook arrangement), an ncu source capture of the
# GEGLU GEMM, and the resident-vs-host-buffer split of one txt2img batch
mkdir -p gpurun_out
echo "== new gelu"; timeout 300 python tools/bench_geglu.py 2>&1 | tail -6
timeout 300 ncu --set full --import-source on --clock-control none -k regex:gemm_kernel -c 1 -o gpurun_out/geglu_full -f python tools/bench_geglu.py once > gpurun_out/geglu_ncu.log 2>&1
echo "ncu rc=$?"
timeout 300 python tools/profile_pipeline.py --config sd15 --iters 4 2>&1 | tail -3
cd stable-diffusion-webui_b200/csrc && touch gemm.cu && make EXTRA=-DSDXE_GELU_V1 > /dev/null 2>&1; cd ../..
echo "== old gelu"; timeout 300 python tools/bench_geglu.py 2>&1 | tail -6
