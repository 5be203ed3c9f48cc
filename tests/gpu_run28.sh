#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/ops_*.csv
run() {
  name=$1; shift
  timeout 1200 "$@" > gpurun_out/r28_$name.log 2>&1
  echo "== $name rc=$? : $(tail -n 1 gpurun_out/r28_$name.log | cut -c1-400)"
}
run pytest python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider
grep -h -E "watchdog|Error|error" gpurun_out/r28_pytest.log | head -5
SDXE_PROFILE_DUMP=gpurun_out/ops_sd15.csv timeout 300 python tools/profile_unet.py --config sd15 --iters 1 --profile > gpurun_out/r28_prof.log 2>&1
tail -8 gpurun_out/r28_prof.log
timeout 300 python tools/profile_unet.py --config sd15 --iters 30 > gpurun_out/r28_graph.log 2>&1
grep "unet forward" gpurun_out/r28_graph.log
run smoke python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
run bench python bench.py --steps 3 --warmup 3 --no-extras
