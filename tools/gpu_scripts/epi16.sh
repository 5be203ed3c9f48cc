#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_prims_gpu.py tests/test_engine_gpu.py tests/test_img2img_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
timeout 900 python -m pytest tests/test_fullsize_gpu.py -m gpu -q -x -p no:cacheprovider -k "unet" 2>&1 | tail -3
SDXE_PROFILE_DUMP=gpurun_out/ops_sd15.csv timeout 300 python tools/profile_unet.py --config sd15 --iters 1 --profile > /dev/null 2>&1
python tools/analyze_ops.py gpurun_out/ops_sd15.csv > gpurun_out/ops_sd15.txt; head -3 gpurun_out/ops_sd15.txt | tail -2; grep "gemm:" gpurun_out/ops_sd15.txt | head -12
SDXE_PROFILE_DUMP=gpurun_out/ops_sdxl.csv timeout 300 python tools/profile_unet.py --config sdxl --iters 1 --profile > /dev/null 2>&1
python tools/analyze_ops.py gpurun_out/ops_sdxl.csv > gpurun_out/ops_sdxl.txt; head -1 gpurun_out/ops_sdxl.txt
timeout 300 python tools/profile_unet.py --config sd15 --iters 10 2>&1 | tail -2 | head -1
timeout 300 python tools/profile_unet.py --config sdxl --iters 5 2>&1 | tail -2 | head -1
timeout 300 python tools/profile_unet.py --config sd15 --iters 3 --vae-only --profile 2>&1 | tail -6
