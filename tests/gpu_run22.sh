#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/ops_*.csv
run() {
  name=$1; shift
  timeout 900 "$@" > gpurun_out/r22_$name.log 2>&1
  echo "== $name rc=$? : $(tail -n 1 gpurun_out/r22_$name.log | cut -c1-300)"
}
run pytest python -m pytest tests/test_prims_gpu.py tests/test_engine_gpu.py -m gpu -q -x --no-header -p no:cacheprovider
grep -h -E "watchdog|Error|error" gpurun_out/r22_pytest.log | head -5
SDXE_PROFILE_DUMP=gpurun_out/ops_sd15.csv timeout 300 python tools/profile_unet.py --config sd15 --iters 1 --profile > gpurun_out/r22_prof.log 2>&1
tail -8 gpurun_out/r22_prof.log
timeout 300 python tools/profile_unet.py --config sd15 --iters 20 > gpurun_out/r22_graph_sd15.log 2>&1
tail -2 gpurun_out/r22_graph_sd15.log
SDXE_GN_ONEPASS=0 timeout 300 python tools/profile_unet.py --config sd15 --iters 20 > gpurun_out/r22_graph_sd15_gn3.log 2>&1
tail -2 gpurun_out/r22_graph_sd15_gn3.log
run bench python bench.py --steps 3 --warmup 3 --no-extras
