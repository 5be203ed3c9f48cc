#!/bin/bash
# 16-warp GEGLU epilogue: correctness (GEMM + engine parity tests), then the shape table and the UNet forward time
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_prims_gpu.py tests/test_engine_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -4
timeout 300 python tools/bench_geglu.py 2>&1 | tail -6
timeout 300 python tools/profile_unet.py --config sd15 --iters 10 2>&1 | tail -4
timeout 300 python tools/profile_unet.py --config sdxl --iters 5 2>&1 | tail -4
