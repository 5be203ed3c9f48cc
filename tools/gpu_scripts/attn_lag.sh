#!/bin/bash
# A/B of the exponential loop's software-pipelining depth in attention2 (rebuilds attention2.cu on the box per variant)
for l in 1 2 4; do
  touch stable-diffusion-webui_b200/csrc/attention2.cu
  make -C stable-diffusion-webui_b200/csrc EXTRA=-DATT2_LAG=$l -j8 > /dev/null 2>&1
  echo "ATT2_LAG=$l"; timeout 200 python tools/bench_attn.py --shapes sd15_l0,sdxl_l1 --iters 20 --check 2>&1 | grep us
done
