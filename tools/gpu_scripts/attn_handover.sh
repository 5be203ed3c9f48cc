#!/bin/bash
# A/B: where in the exponential loop the MUFU turn token is handed over (16 = after the loop)
for h in 16 12 8 4; do
  touch stable-diffusion-webui_b200/csrc/attention2.cu
  make -C stable-diffusion-webui_b200/csrc EXTRA=-DATT2_HANDOVER=$h -j8 > /dev/null 2>&1
  echo "ATT2_HANDOVER=$h"; timeout 200 python tools/bench_attn.py --shapes sd15_l0,sdxl_l1,sd15_l1 --iters 20 --check 2>&1 | grep us
done
