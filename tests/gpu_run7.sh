#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/ops_*.csv
run() {
  name=$1; shift
  timeout 1200 "$@" > gpurun_out/r7_$name.log 2>&1
  echo "== $name rc=$? : $(tail -n 1 gpurun_out/r7_$name.log | cut -c1-300)"
}
run pytest python -m pytest tests/test_prims_gpu.py tests/test_engine_gpu.py -m gpu -q -x --no-header -p no:cacheprovider
SDXE_PROFILE_DUMP=gpurun_out/ops_new.csv python tools/profile_unet.py --config sd15 --iters 1 --profile > gpurun_out/r7_prof.log 2>&1
tail -7 gpurun_out/r7_prof.log
SDXE_NO_GRAPH=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 3 -c 9 -o gpurun_out/r7_gemm python tools/profile_unet.py --config sd15 --iters 1 > gpurun_out/r7_ncu_gemm.log 2>&1
echo "ncu gemm rc=$?"
SDXE_NO_GRAPH=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_kernel -s 0 -c 2 -o gpurun_out/r7_attn python tools/profile_unet.py --config sd15 --iters 1 > gpurun_out/r7_ncu_attn.log 2>&1
echo "ncu attn rc=$?"
