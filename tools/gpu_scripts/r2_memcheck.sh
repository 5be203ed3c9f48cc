#!/bin/bash
# compute-sanitizer memcheck over the tests that reach the kernels changed late in round 2 (16-warp GEGLU epilogue,
# residual stream, implicit stride-2 conv, mma.sync skinny linear)
mkdir -p gpurun_out
timeout 1100 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_engine_gpu.py tests/test_prims_gpu.py tests/test_img2img_gpu.py -m gpu -q -x -p no:cacheprovider -k "tiny_unet or tiny_vae or gemm_geglu or gemm_plain or tiny_img2img" > gpurun_out/r2_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "passed|failed|ERROR SUMMARY|Invalid|out of bounds" gpurun_out/r2_memcheck.log | tail -6
