#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/ops_*.csv
run() {
  name=$1; shift
  timeout 1200 "$@" > gpurun_out/r29_$name.log 2>&1
  echo "== $name rc=$? : $(tail -n 1 gpurun_out/r29_$name.log | cut -c1-400)"
}
run pytest python -m pytest tests/test_engine_gpu.py tests/test_pipeline_gpu.py -m gpu -q -x --no-header -p no:cacheprovider
grep -h -E "watchdog|Error|error|assert" gpurun_out/r29_pytest.log | head -8
SDXE_PROFILE_DUMP=gpurun_out/ops_sd15.csv timeout 300 python tools/profile_unet.py --config sd15 --iters 1 --profile > gpurun_out/r29_prof.log 2>&1
tail -8 gpurun_out/r29_prof.log
timeout 300 python tools/profile_unet.py --config sd15 --iters 30 > gpurun_out/r29_graph.log 2>&1
grep "unet forward" gpurun_out/r29_graph.log
timeout 300 python tools/profile_unet.py --config sdxl --iters 10 > gpurun_out/r29_graph_sdxl.log 2>&1
grep "unet forward" gpurun_out/r29_graph_sdxl.log
