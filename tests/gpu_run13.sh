#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/ops_*.csv
run() {
  name=$1; shift
  timeout 1500 "$@" > gpurun_out/r13_$name.log 2>&1
  echo "== $name rc=$? : $(tail -n 1 gpurun_out/r13_$name.log | cut -c1-300)"
}
run pytest python -m pytest tests/test_prims_gpu.py tests/test_engine_gpu.py -m gpu -q -x --no-header -p no:cacheprovider
SDXE_PROFILE_DUMP=gpurun_out/ops_new.csv python tools/profile_unet.py --config sd15 --iters 1 --profile > gpurun_out/r13_prof.log 2>&1
tail -7 gpurun_out/r13_prof.log
SDXE_PROFILE_DUMP=gpurun_out/ops_sdxl.csv python tools/profile_unet.py --config sdxl --iters 1 --profile > gpurun_out/r13_prof_sdxl.log 2>&1
tail -7 gpurun_out/r13_prof_sdxl.log
run bench python bench.py --steps 3 --warmup 3 --no-extras
