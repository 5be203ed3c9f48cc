#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/ops_*.csv
run() {
  name=$1; shift
  timeout 900 "$@" > gpurun_out/r16_$name.log 2>&1
  echo "== $name rc=$? : $(tail -n 1 gpurun_out/r16_$name.log | cut -c1-300)"
}
run pytest python -m pytest tests/test_prims_gpu.py tests/test_engine_gpu.py -m gpu -q -x --no-header -p no:cacheprovider
grep -h -E "watchdog|Error|error" gpurun_out/r16_pytest.log | head -5
SDXE_PROFILE_DUMP=gpurun_out/ops_nocl.csv timeout 300 python tools/profile_unet.py --config sd15 --iters 1 --profile > gpurun_out/r16_prof_nocl.log 2>&1
tail -7 gpurun_out/r16_prof_nocl.log
SDXE_CLUSTER=1 SDXE_PROFILE_DUMP=gpurun_out/ops_cl.csv timeout 300 python tools/profile_unet.py --config sd15 --iters 1 --profile > gpurun_out/r16_prof_cl.log 2>&1
tail -7 gpurun_out/r16_prof_cl.log
run bench python bench.py --steps 3 --warmup 3 --no-extras
