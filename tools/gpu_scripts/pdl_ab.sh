#!/bin/bash
mkdir -p gpurun_out
for pdl in 1 0 1 0; do
  SDXE_PDL=$pdl timeout 300 python tools/profile_unet.py --config sd15 --iters 30 > gpurun_out/r24_graph_pdl$pdl.log 2>&1
  echo "PDL=$pdl $(grep 'unet forward' gpurun_out/r24_graph_pdl$pdl.log)"
done
for pdl in 1 0; do
  SDXE_PDL=$pdl timeout 300 python tools/profile_unet.py --config sdxl --iters 10 > gpurun_out/r24_graph_sdxl_pdl$pdl.log 2>&1
  echo "SDXL PDL=$pdl $(grep 'unet forward' gpurun_out/r24_graph_sdxl_pdl$pdl.log)"
done
