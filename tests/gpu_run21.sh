#!/bin/bash
mkdir -p gpurun_out
SDXE_NO_GRAPH=1 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:gemm_kernel --csv --log-file gpurun_out/r21_gemm_dram.csv python tools/profile_unet.py --config sd15 --iters 1 > gpurun_out/r21_ncu_dram.log 2>&1
echo "ncu dram rc=$? lines=$(wc -l < gpurun_out/r21_gemm_dram.csv)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file gpurun_out/r21_bench_launches.csv python bench.py --steps 1 --warmup 1 --no-extras > gpurun_out/r21_ncu_bench.log 2>&1
echo "ncu bench rc=$? lines=$(wc -l < gpurun_out/r21_bench_launches.csv)"
timeout 900 python bench.py --config sdxl --steps 2 --warmup 3 --no-extras > gpurun_out/r21_bench_sdxl.log 2>&1
echo "sdxl: $(tail -n 1 gpurun_out/r21_bench_sdxl.log | cut -c1-400)"
