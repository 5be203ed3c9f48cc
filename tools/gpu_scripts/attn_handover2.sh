#!/bin/bash
for cfg in "-DATT2_HANDOVER=6" "-DATT2_HANDOVER=7" "-DATT2_HANDOVER=9" "-DATT2_HANDOVER=10" "-DATT2_HANDOVER=8 -DATT2_POLY_MASK=0x80" "-DATT2_HANDOVER=8 -DATT2_POLY_MASK=0x88"; do
  touch stable-diffusion-webui_b200/csrc/attention2.cu
  make -C stable-diffusion-webui_b200/csrc EXTRA="$cfg" -j8 > /dev/null 2>&1
  echo "CFG $cfg"; timeout 200 python tools/bench_attn.py --shapes sd15_l0,sdxl_l1 --iters 20 2>&1 | grep us
done
