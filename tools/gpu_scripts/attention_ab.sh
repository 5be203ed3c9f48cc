#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/ops_*.csv
run() {
  name=$1; shift
  timeout 900 "$@" > gpurun_out/r33_$name.log 2>&1
  echo "== $name rc=$? : $(tail -n 1 gpurun_out/r33_$name.log | cut -c1-300)"
}
run pytest_attn python -m pytest tests/test_prims_gpu.py -m gpu -q -x --no-header -p no:cacheprovider -k attention
grep -h -E "watchdog|Error|error|assert" gpurun_out/r33_pytest_attn.log | head -8
run pytest_eng python -m pytest tests/test_engine_gpu.py tests/test_pipeline_gpu.py -m gpu -q -x --no-header -p no:cacheprovider
SDXE_PROFILE_DUMP=gpurun_out/ops_a4.csv timeout 300 python tools/profile_unet.py --config sd15 --iters 1 --profile > gpurun_out/r33_prof_a4.log 2>&1
grep attention gpurun_out/r33_prof_a4.log
for m in 4 2 4 2; do
  SDXE_ATTN=$m timeout 300 python tools/profile_unet.py --config sd15 --iters 30 > gpurun_out/r33_graph_a$m.log 2>&1
  echo "ATTN=$m $(grep 'unet forward' gpurun_out/r33_graph_a$m.log)"
done
for m in 4 2; do
  SDXE_ATTN=$m timeout 300 python tools/profile_unet.py --config sdxl --iters 10 > gpurun_out/r33_graph_sdxl_a$m.log 2>&1
  echo "SDXL ATTN=$m $(grep 'unet forward' gpurun_out/r33_graph_sdxl_a$m.log)"
done
