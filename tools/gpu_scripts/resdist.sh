#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_prims_gpu.py tests/test_engine_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
for d in 0 1; do
  echo "== SDXE_RES_DIST2=$d"
  SDXE_RES_DIST2=$d SDXE_PROFILE_DUMP=gpurun_out/ops_sd15_$d.csv timeout 300 python tools/profile_unet.py --config sd15 --iters 1 --profile > /dev/null 2>&1
  python tools/analyze_ops.py gpurun_out/ops_sd15_$d.csv | grep "res=1" | grep "gemm:" | head -8 | cut -c1-60,118-
  SDXE_RES_DIST2=$d timeout 300 python tools/profile_unet.py --config sd15 --iters 10 2>&1 | tail -2 | head -1
  SDXE_RES_DIST2=$d timeout 300 python tools/profile_unet.py --config sdxl --iters 5 2>&1 | tail -2 | head -1
done
