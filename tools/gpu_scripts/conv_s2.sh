#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_img2img_gpu.py tests/test_pipeline_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
timeout 900 python -m pytest tests/test_fullsize_gpu.py -m gpu -q -x -p no:cacheprovider -k "unet" 2>&1 | tail -3
SDXE_PROFILE_DUMP=gpurun_out/ops_sd15.csv timeout 300 python tools/profile_unet.py --config sd15 --iters 1 --profile > /dev/null 2>&1
grep -i "conv3s2\|,5,," gpurun_out/ops_sd15.csv | head -14
for v in 0 1; do echo "implicit=$v"; SDXE_CONV_S2_IMPLICIT=$v timeout 300 python tools/profile_unet.py --config sd15 --iters 10 2>&1 | tail -2 | head -1; SDXE_CONV_S2_IMPLICIT=$v timeout 300 python tools/profile_unet.py --config sdxl --iters 5 2>&1 | tail -2 | head -1; done
