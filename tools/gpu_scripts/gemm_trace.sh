#!/bin/bash
mkdir -p gpurun_out
cd stable-diffusion-webui_b200/csrc && touch gemm.cu && make GEMM_TRACE=1 > /dev/null 2>&1; cd ../..
SDXE_GEMM_TRACE_DUMP=6 timeout 120 python tools/trace_gemm.py 65536 2560 320 geglu && cp gpurun_out/gemm_trace.txt gpurun_out/gemm_trace_geglu.txt
SDXE_GEMM_TRACE_DUMP=6 timeout 120 python tools/trace_gemm.py 65536 320 320 res && cp gpurun_out/gemm_trace.txt gpurun_out/gemm_trace_res.txt
SDXE_GEMM_TRACE_DUMP=6 timeout 120 python tools/trace_gemm.py 8192 1280 1280 res && cp gpurun_out/gemm_trace.txt gpurun_out/gemm_trace_mid.txt
python tools/analyze_gemm_trace.py gpurun_out/gemm_trace_geglu.txt | head -14
