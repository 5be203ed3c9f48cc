#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/ops_*.csv
run() {
  name=$1; shift
  timeout 1200 "$@" > gpurun_out/r11_$name.log 2>&1
  echo "== $name rc=$? : $(tail -n 1 gpurun_out/r11_$name.log | cut -c1-300)"
}
run pytest python -m pytest tests/test_prims_gpu.py tests/test_engine_gpu.py -m gpu -q -x --no-header -p no:cacheprovider
SDXE_CLUSTER=0 SDXE_PROFILE_DUMP=gpurun_out/ops_nocluster.csv python tools/profile_unet.py --config sd15 --iters 1 --profile > gpurun_out/r11_prof0.log 2>&1
tail -7 gpurun_out/r11_prof0.log
SDXE_PROFILE_DUMP=gpurun_out/ops_new.csv python tools/profile_unet.py --config sd15 --iters 1 --profile > gpurun_out/r11_prof.log 2>&1
tail -7 gpurun_out/r11_prof.log
run bench python bench.py --steps 3 --warmup 3 --no-extras
