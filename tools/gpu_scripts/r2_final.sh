#!/bin/bash
# round-2 final validation on one GPU: whole GPU suite, smoke(), the default bench line, the ncu launch list of the bench
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r2_final_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2_final_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_final_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2_final_smoke.log
timeout 1200 python bench.py > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err
echo "bench rc=$?"; tail -2 gpurun_out/r2_final_bench.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2_final_bench.json').read().strip().splitlines()[-1])
    def show(tag,b):
        print(tag, 'value %.3f e2e %.3f img/s  frac %.3f' % (b['value'], b['e2e']['value'], b.get('whole_job_frac',0)), 'sdp', (b.get('torch_sdp_gpu') or {}).get('value'), 'cpu', (b.get('cpu_baseline') or {}).get('value'), 'parity', b.get('shard_parity'))
    show('sd15 bf16', d)
    for k in ('fp16','sdxl','c4'):
        if k in d: show(k, d[k])
    print('roofline', d['roofline']['frac'], {k:v['ms'] for k,v in d['roofline']['by_kernel_class_unet'].items()})
    print('clocks', d['clocks'], 'launches', d['gpu_launches'])
except Exception as e:
    print('parse failed', e)
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 1500 --csv --log-file gpurun_out/r2_final_launches.csv python bench.py --steps 1 --warmup 1 --no-extras --only-headline > gpurun_out/r2_final_ncu_bench.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/r2_final_launches.csv
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 2>/dev/null | tail -1 | cut -c1-400
