#!/bin/bash
# final-tree evidence for profiles/: sdxe-kernel launch list of the bench under ncu, per-op tables of one UNet forward
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:gemm_kernel|attention|gn_|skinny|upsample|cfg_combine|lincomb|im2col|nhwc_to|cast_|timestep_emb|layer_norm|act_inplace' -s 1200 -c 1500 --csv --log-file gpurun_out/r2_final_launches.csv python bench.py --steps 1 --warmup 1 --no-extras --only-headline > gpurun_out/r2_final_ncu_bench.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/r2_final_launches.csv
for c in sd15 sdxl; do
  SDXE_PROFILE_DUMP=gpurun_out/ops_$c.csv timeout 300 python tools/profile_unet.py --config $c --iters 1 --profile > /dev/null 2>&1
  python tools/analyze_ops.py gpurun_out/ops_$c.csv > gpurun_out/r2_unet_ops_$c.txt; head -1 gpurun_out/r2_unet_ops_$c.txt
  timeout 300 python tools/profile_unet.py --config $c --iters 10 2>&1 | tail -2 | head -1
done
timeout 300 python tools/profile_unet.py --config sd15 --iters 3 --vae-only --profile 2>&1 | tail -7
