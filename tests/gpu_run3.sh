#!/bin/bash
mkdir -p gpurun_out
run() {
  name=$1; shift
  timeout 1200 "$@" > gpurun_out/r3_$name.log 2>&1
  echo "== $name rc=$? : $(tail -n 1 gpurun_out/r3_$name.log | cut -c1-300)"
}
run pytest python -m pytest tests -m gpu -q -s --no-header -p no:cacheprovider
run smoke python __graft_entry__.py smoke
run bench python bench.py --steps 3 --warmup 3
grep -h -E "rel err|engine [0-9]|^(FAILED|ERROR)|watchdog|smoke:" gpurun_out/r3_*.log | head -80
tail -c 6000 gpurun_out/r3_bench.log
