#!/bin/bash
# final-tree parity numbers (printed by the full-size tests) and one ncu --set full capture of the 16-warp GEGLU GEMM
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_fullsize_gpu.py -m gpu -q -x -s -p no:cacheprovider > gpurun_out/r2_fullsize_final.log 2>&1
echo "fullsize rc=$?"; grep -E "^C[234]|^  |passed|failed|rror|sd15 UNet|sdxl UNet|VAE decode" gpurun_out/r2_fullsize_final.log | head -40
timeout 300 ncu --set full --import-source on --clock-control none -k regex:gemm_kernel -c 1 -o gpurun_out/geglu16_full -f python tools/bench_geglu.py once > gpurun_out/geglu16_ncu.log 2>&1
echo "ncu rc=$?"
ncu -i gpurun_out/geglu16_full.ncu-rep --page raw --csv 2>/dev/null > gpurun_out/geglu16_raw.csv; wc -c gpurun_out/geglu16_raw.csv
