#!/bin/bash
# knockout sweep of the attention2 softmax / MMA work items (timing only; outputs are garbage for ko != 0)
mkdir -p gpurun_out
for ko in 0 32 33 34 36 40 48 35 96 160 224 63 255; do
  echo "KO=$ko $(SDXE_ATT_KO=$ko timeout 120 python tools/bench_attn.py --shapes sd15_l0,sdxl_l1 --iters 10 2>&1 | tr '\n' '|')"
done 2>&1 | tee gpurun_out/r2_attn_ko.txt
