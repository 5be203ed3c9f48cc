#!/bin/bash
mkdir -p gpurun_out
cd stable-diffusion-webui_b200/csrc && touch gemm.cu && make GEMM_TRACE=1 > /dev/null 2>&1; cd ../..
for cfg in "65536 320 320 res" "65536 320 320 plain"; do
  set -- $cfg
  SDXE_GEMM_TRACE_DUMP=6 timeout 120 python tools/trace_gemm.py $1 $2 $3 $4
  echo "== $cfg"; python tools/analyze_gemm_trace.py gpurun_out/gemm_trace.txt 2>/dev/null | tail -11
done
