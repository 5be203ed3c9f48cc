#!/bin/bash
mkdir -p gpurun_out
run() {
  name=$1; shift
  timeout 1200 "$@" > gpurun_out/r4_$name.log 2>&1
  echo "== $name rc=$? : $(tail -n 1 gpurun_out/r4_$name.log | cut -c1-400)"
}
run pytest python -m pytest tests/test_prims_gpu.py tests/test_engine_gpu.py -m gpu -q -x --no-header -p no:cacheprovider
run prof python tools/profile_unet.py --config sd15 --iters 3 --profile
cat gpurun_out/r4_prof.log | tail -8
run bench python bench.py --steps 3 --warmup 3 --no-extras
SDXE_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 700 --csv --log-file gpurun_out/r4_launches.csv python tools/profile_unet.py --config sd15 --iters 2 > gpurun_out/r4_ncu.log 2>&1
echo "ncu rc=$?"
python - <<'PY'
import csv, collections
rows=[]
try:
    with open('gpurun_out/r4_launches.csv') as f:
        lines=[l for l in f if not l.startswith('==')]
    r=csv.DictReader(lines)
    agg=collections.defaultdict(lambda:[0,0.0])
    for row in r:
        k=row['Kernel Name'][:60]; v=float(row['Metric Value'].replace(',',''))
        if row['Metric Unit'] in ('nsecond','ns'): v/=1e3
        elif row['Metric Unit'] in ('msecond','ms'): v*=1e3
        agg[k][0]+=1; agg[k][1]+=v
    tot=sum(v[1] for v in agg.values())
    for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1]):
        print(f"{k:60s} n={v[0]:4d} total_us={v[1]:10.1f} share={v[1]/tot:6.3f} avg_us={v[1]/v[0]:8.1f}")
except Exception as e:
    print('parse failed', e)
PY
