#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/ops_*.csv
run() {
  name=$1; shift
  timeout 900 "$@" > gpurun_out/r18_$name.log 2>&1
  echo "== $name rc=$? : $(tail -n 1 gpurun_out/r18_$name.log | cut -c1-300)"
}
run pytest python -m pytest tests/test_prims_gpu.py tests/test_engine_gpu.py -m gpu -q -x --no-header -p no:cacheprovider
grep -h -E "watchdog|Error|error" gpurun_out/r18_pytest.log | head -5
SDXE_PROFILE_DUMP=gpurun_out/ops_a3.csv timeout 300 python tools/profile_unet.py --config sd15 --iters 1 --profile > gpurun_out/r18_prof_a3.log 2>&1
tail -7 gpurun_out/r18_prof_a3.log
SDXE_ATTN=2 SDXE_PROFILE_DUMP=gpurun_out/ops_a2.csv timeout 300 python tools/profile_unet.py --config sd15 --iters 1 --profile > gpurun_out/r18_prof_a2.log 2>&1
tail -7 gpurun_out/r18_prof_a2.log | grep attention
run bench python bench.py --steps 3 --warmup 3 --no-extras
