#!/bin/bash
# A/B of the FMA-pipe exp2 fraction in attention2 (rebuilds attention2.cu on the box per variant)
for m in 0x00 0x88 0xA8 0xAA; do
  touch stable-diffusion-webui_b200/csrc/attention2.cu
  make -C stable-diffusion-webui_b200/csrc EXTRA=-DATT2_POLY_MASK=$m -j8 > /dev/null 2>&1
  echo "POLY_MASK=$m"; timeout 200 python tools/bench_attn.py --shapes sd15_l0,sdxl_l1 --iters 20 --check 2>&1 | grep us
done
