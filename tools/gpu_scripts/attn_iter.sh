#!/bin/bash
# attention iteration: correctness (prims tests), timing of the stand-alone shapes, then a timeline trace build
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_prims_gpu.py -m gpu -q -x --no-header -p no:cacheprovider -k attention 2>&1 | tail -3
timeout 300 python tools/bench_attn.py --check --iters 20 2>&1 | tee gpurun_out/r2_attn_bench.txt
if [ "$1" == "trace" ]; then
  touch stable-diffusion-webui_b200/csrc/attention2.cu
  make -C stable-diffusion-webui_b200/csrc ATT_TRACE=1 -j8 > /dev/null 2>&1
  SDXE_ATT_TRACE_DUMP=1 timeout 120 python tools/bench_attn.py --shapes sd15_l0 --iters 2 > /dev/null 2>&1
  ls -la gpurun_out/attn_trace.txt
fi
