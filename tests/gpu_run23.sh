#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/ops_*.csv
run() {
  name=$1; shift
  timeout 900 "$@" > gpurun_out/r23_$name.log 2>&1
  echo "== $name rc=$? : $(tail -n 1 gpurun_out/r23_$name.log | cut -c1-300)"
}
run pytest python -m pytest tests/test_prims_gpu.py tests/test_engine_gpu.py -m gpu -q -x --no-header -p no:cacheprovider
grep -h -E "watchdog|Error|error" gpurun_out/r23_pytest.log | head -5
for pdl in 1 0 1 0; do
  SDXE_PDL=$pdl timeout 300 python tools/profile_unet.py --config sd15 --iters 30 > gpurun_out/r23_graph_pdl$pdl.log 2>&1
  echo "PDL=$pdl $(grep 'unet forward' gpurun_out/r23_graph_pdl$pdl.log)"
done
SDXE_PROFILE_DUMP=gpurun_out/ops_sd15.csv timeout 300 python tools/profile_unet.py --config sd15 --iters 1 --profile > gpurun_out/r23_prof.log 2>&1
tail -8 gpurun_out/r23_prof.log
run bench python bench.py --steps 3 --warmup 3 --no-extras
