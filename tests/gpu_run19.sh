#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/ops_*.csv
SDXE_PROFILE_DUMP=gpurun_out/ops_vae.csv timeout 300 python tools/profile_unet.py --config sd15 --iters 1 --profile --vae-only > gpurun_out/r19_prof_vae.log 2>&1
tail -8 gpurun_out/r19_prof_vae.log
SDXE_PROFILE_DUMP=gpurun_out/ops_sdxl.csv timeout 600 python tools/profile_unet.py --config sdxl --iters 1 --profile > gpurun_out/r19_prof_sdxl.log 2>&1
tail -8 gpurun_out/r19_prof_sdxl.log
