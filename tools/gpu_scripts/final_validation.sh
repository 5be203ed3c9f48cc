#!/bin/bash
# round-end validation: everything the driver re-runs, plus the ncu evidence copied into profiles/
mkdir -p gpurun_out
run() {
  name=$1; shift
  timeout 1500 "$@" > gpurun_out/final_$name.log 2>&1
  echo "== $name rc=$? : $(tail -n 1 gpurun_out/final_$name.log | cut -c1-300)"
}
run pytest python -m pytest tests -m gpu -q -s --no-header -p no:cacheprovider
grep -h -E "engine |lora:|watchdog|FAILED|Error" gpurun_out/final_pytest.log | head -60 > gpurun_out/final_parity.txt
wc -l gpurun_out/final_parity.txt
run smoke python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
run bench python bench.py --steps 3 --warmup 3
run bench_sdxl python bench.py --config sdxl --steps 2 --warmup 3 --no-extras
SDXE_NO_GRAPH=1 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:gemm_kernel --csv --log-file gpurun_out/final_gemm_dram.csv python tools/profile_unet.py --config sd15 --iters 1 --profile > gpurun_out/final_ncu_dram.log 2>&1
echo "ncu dram rc=$? lines=$(wc -l < gpurun_out/final_gemm_dram.csv)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:sdxe -c 1500 --csv --log-file gpurun_out/final_bench_launches.csv python bench.py --steps 1 --warmup 1 --no-extras > gpurun_out/final_ncu_bench.log 2>&1
echo "ncu bench rc=$? lines=$(wc -l < gpurun_out/final_bench_launches.csv)"
