#!/bin/bash
# round-2 validation: full-size parity tests (printed numbers), the whole GPU suite, the whole-metric bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_fullsize_gpu.py -m gpu -q -x -s -p no:cacheprovider > gpurun_out/r2_fullsize.log 2>&1
echo "fullsize rc=$?"; grep -E "^C[234]|^  |passed|failed|Error|error|sd15 UNet|sdxl UNet|VAE decode" gpurun_out/r2_fullsize.log | head -60
if [ "$1" != "quick" ]; then
  timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider --deselect tests/test_fullsize_gpu.py 2>&1 | tail -3
fi
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r2_bench.err; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2_bench.json').read().strip().splitlines()[-1])
    def show(tag,b):
        print(tag, 'value %.3f e2e %.3f img/s  frac %.3f' % (b['value'], b['e2e']['value'], b.get('whole_job_frac',0)), 'sdp', (b.get('torch_sdp_gpu') or {}).get('value'), 'cpu', (b.get('cpu_baseline') or {}).get('value'), 'parity', b.get('shard_parity'))
    show('sd15 bf16', d)
    for k in ('fp16','sdxl','c4'):
        if k in d: show(k, d[k])
    print('roofline', d['roofline']['frac'], {k:v['ms'] for k,v in d['roofline']['by_kernel_class_unet'].items()})
    print('clocks', d['clocks'])
except Exception as e:
    print('parse failed', e)
PY
