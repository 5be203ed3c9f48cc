#!/bin/bash
# compute-sanitizer memcheck over the small GPU tests (tiny UNet / VAE / img2img / LoRA + a few primitives)
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 --launch-timeout 120 python -m pytest tests/test_engine_gpu.py tests/test_img2img_gpu.py tests/test_lora_gpu.py -m gpu -q -x --no-header -p no:cacheprovider -k "tiny or lora" > gpurun_out/san_memcheck.log 2>&1
echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|Invalid|out of bounds|misaligned|passed|failed" gpurun_out/san_memcheck.log | head -20
