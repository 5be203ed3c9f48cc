#!/bin/bash
mkdir -p gpurun_out
run() {
  name=$1; shift
  timeout 1200 "$@" > gpurun_out/r31_$name.log 2>&1
  echo "== $name rc=$? : $(tail -n 1 gpurun_out/r31_$name.log | cut -c1-400)"
}
run pytest_new python -m pytest tests/test_img2img_gpu.py tests/test_sd_models_gpu.py -m gpu -q -x -s --no-header -p no:cacheprovider
grep -h -E "watchdog|Error|error|assert|engine " gpurun_out/r31_pytest_new.log | head -20
