#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/ops_*.csv
run() {
  name=$1; shift
  timeout 900 "$@" > gpurun_out/r26_$name.log 2>&1
  echo "== $name rc=$? : $(tail -n 1 gpurun_out/r26_$name.log | cut -c1-300)"
}
run pytest_attn python -m pytest tests/test_prims_gpu.py -m gpu -q -x --no-header -p no:cacheprovider -k attention
grep -h -E "watchdog|Error|error|assert" gpurun_out/r26_pytest_attn.log | head -8
SDXE_PROFILE_DUMP=gpurun_out/ops_sd15.csv timeout 300 python tools/profile_unet.py --config sd15 --iters 1 --profile > gpurun_out/r26_prof.log 2>&1
grep attention gpurun_out/r26_prof.log
SDXE_PROFILE_DUMP=gpurun_out/ops_sdxl.csv timeout 300 python tools/profile_unet.py --config sdxl --iters 1 --profile > gpurun_out/r26_prof_sdxl.log 2>&1
tail -8 gpurun_out/r26_prof_sdxl.log
