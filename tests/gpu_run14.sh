#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/ops_*.csv
run() {
  name=$1; shift
  timeout 900 "$@" > gpurun_out/r14_$name.log 2>&1
  echo "== $name rc=$? : $(tail -n 1 gpurun_out/r14_$name.log | cut -c1-300)"
}
SDXE_CLUSTER=0 run pytest_nocl python -m pytest tests/test_prims_gpu.py tests/test_engine_gpu.py -m gpu -q -x --no-header -p no:cacheprovider
SDXE_CLUSTER=0 SDXE_PROFILE_DUMP=gpurun_out/ops_nocl.csv python tools/profile_unet.py --config sd15 --iters 1 --profile > gpurun_out/r14_prof_nocl.log 2>&1
tail -7 gpurun_out/r14_prof_nocl.log
run pytest_gemm python -m pytest tests/test_prims_gpu.py -m gpu -q -x --no-header -p no:cacheprovider -k "gemm or conv"
grep -h -E "watchdog|Error|error" gpurun_out/r14_pytest_gemm.log | head -5
run pytest_cl python -m pytest tests/test_engine_gpu.py -m gpu -q -x --no-header -p no:cacheprovider
SDXE_PROFILE_DUMP=gpurun_out/ops_cl.csv timeout 300 python tools/profile_unet.py --config sd15 --iters 1 --profile > gpurun_out/r14_prof_cl.log 2>&1
tail -7 gpurun_out/r14_prof_cl.log
run bench python bench.py --steps 3 --warmup 3 --no-extras
