#!/bin/bash
# Bring-up driver: each group in its own process (a trapped kernel poisons the CUDA context), bounded by `timeout`.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
nproc >> gpurun_out/smi.txt
run() {
  name=$1; shift
  timeout 300 python -m pytest tests/test_prims_gpu.py -m gpu -q -x --no-header -p no:cacheprovider "$@" > gpurun_out/bringup_$name.log 2>&1
  echo "== $name rc=$? : $(tail -n 1 gpurun_out/bringup_$name.log)"
}
run gemm_small -k "test_gemm_plain and 128-64-64"
run gemm_plain -k "test_gemm_plain"
run gemm_tiles -k "test_gemm_tile_widths"
run geglu -k "test_gemm_geglu"
run conv -k "test_conv3x3"
run attn_small -k "test_attention and 1-1-128-128-64"
run attn -k "test_attention"
run gn -k "test_group_norm"
run ln -k "test_layer_norm"
grep -h -E "^(FAILED|ERROR)|assert|Error|watchdog" gpurun_out/bringup_*.log | head -60
