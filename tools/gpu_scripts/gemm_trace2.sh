#!/bin/bash
mkdir -p gpurun_out
cd stable-diffusion-webui_b200/csrc && touch gemm.cu && make GEMM_TRACE=1 > /dev/null 2>&1; cd ../..
for deep in 0 1; do
  for cfg in "65536 320 320 res" "65536 320 320 plain" "65536 960 320 plain" "16384 640 640 res"; do
    set -- $cfg
    SDXE_EPI_DEEP=$deep SDXE_GEMM_TRACE_DUMP=6 timeout 120 python tools/trace_gemm.py $1 $2 $3 $4
    echo "== deep=$deep $cfg"; python tools/analyze_gemm_trace.py gpurun_out/gemm_trace.txt 2>/dev/null | sed -n '1p;7,10p' | cut -c1-40,150-
  done
done
