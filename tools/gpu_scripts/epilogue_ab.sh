#!/bin/bash
# deeper epilogue staging (one more chunk buffer per warp, cp.async.bulk.wait_group.read 1): validation + same-box A/B
mkdir -p gpurun_out
rm -f gpurun_out/ops_*.csv
run() {
  name=$1; shift
  timeout 900 "$@" > gpurun_out/r35_$name.log 2>&1
  echo "== $name rc=$? : $(tail -n 1 gpurun_out/r35_$name.log | cut -c1-300)"
}
run pytest python -m pytest tests/test_prims_gpu.py tests/test_engine_gpu.py -m gpu -q -x --no-header -p no:cacheprovider
grep -h -E "watchdog|Error|error|assert" gpurun_out/r35_pytest.log | head -8
SDXE_PROFILE_DUMP=gpurun_out/ops_deep.csv timeout 300 python tools/profile_unet.py --config sd15 --iters 1 --profile > gpurun_out/r35_prof.log 2>&1
grep -E "gemm|conv" gpurun_out/r35_prof.log
for d in 1 0 1 0; do
  SDXE_EPI_DEEP=$d timeout 300 python tools/profile_unet.py --config sd15 --iters 30 > gpurun_out/r35_graph_d$d.log 2>&1
  echo "DEEP=$d $(grep 'unet forward' gpurun_out/r35_graph_d$d.log)"
done
for d in 1 0; do
  SDXE_EPI_DEEP=$d timeout 300 python tools/profile_unet.py --config sdxl --iters 10 > gpurun_out/r35_graph_sdxl_d$d.log 2>&1
  echo "SDXL DEEP=$d $(grep 'unet forward' gpurun_out/r35_graph_sdxl_d$d.log)"
done
