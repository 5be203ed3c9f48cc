#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/ops_*.csv
run() {
  name=$1; shift
  timeout 900 "$@" > gpurun_out/r25_$name.log 2>&1
  echo "== $name rc=$? : $(tail -n 1 gpurun_out/r25_$name.log | cut -c1-300)"
}
run pytest_attn python -m pytest tests/test_prims_gpu.py -m gpu -q -x --no-header -p no:cacheprovider -k attention
grep -h -E "watchdog|Error|error|assert" gpurun_out/r25_pytest_attn.log | head -8
run pytest_eng python -m pytest tests/test_engine_gpu.py -m gpu -q -x --no-header -p no:cacheprovider
SDXE_PROFILE_DUMP=gpurun_out/ops_sd15.csv timeout 300 python tools/profile_unet.py --config sd15 --iters 1 --profile > gpurun_out/r25_prof.log 2>&1
tail -8 gpurun_out/r25_prof.log
for x in 1 0; do
  SDXE_ATTNX=$x timeout 300 python tools/profile_unet.py --config sd15 --iters 30 > gpurun_out/r25_graph_x$x.log 2>&1
  echo "ATTNX=$x $(grep 'unet forward' gpurun_out/r25_graph_x$x.log)"
done
