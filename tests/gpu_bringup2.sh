#!/bin/bash
mkdir -p gpurun_out
run() {
  name=$1; shift
  timeout 600 "$@" > gpurun_out/b2_$name.log 2>&1
  echo "== $name rc=$? : $(tail -n 1 gpurun_out/b2_$name.log)"
}
SDXE_NO_GRAPH=1 run tiny_nograph python -m pytest tests/test_engine_gpu.py -m gpu -q -x -s --no-header -p no:cacheprovider -k "test_tiny_unet"
run tiny_graph python -m pytest tests/test_engine_gpu.py -m gpu -q -x -s --no-header -p no:cacheprovider -k "test_tiny_unet"
run tiny_vae python -m pytest tests/test_engine_gpu.py -m gpu -q -x -s --no-header -p no:cacheprovider -k "test_tiny_vae"
run sd15 python -m pytest tests/test_engine_gpu.py -m gpu -q -x -s --no-header -p no:cacheprovider -k "test_sd15_unet_forward"
run full_vae python -m pytest tests/test_engine_gpu.py -m gpu -q -x -s --no-header -p no:cacheprovider -k "test_full_vae"
grep -h -E "engine [0-9]|^(FAILED|ERROR)|Error|watchdog|assert" gpurun_out/b2_*.log | head -60
