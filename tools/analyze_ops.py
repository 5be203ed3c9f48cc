"""Summarise an SDXE_PROFILE_DUMP csv: idx,kind,desc,us,flops,bytes — grouped by op description."""
import collections
import sys

KIND = ["gemm", "conv3x3", "attention", "group_norm", "layer_norm", "other"]
path = sys.argv[1]
peak_tf, peak_gbs = 1384.2, 6473.0
agg = collections.OrderedDict()
tot = 0.0
for line in open(path):
    parts = line.rstrip("\n").split(",")
    if len(parts) < 6:
        continue
    idx, kind, us, fl, by = int(parts[0]), int(parts[1]), float(parts[-3]), float(parts[-2]), float(parts[-1])
    desc = ",".join(parts[2:-3]) or KIND[kind]
    a = agg.setdefault((kind, desc), [0, 0.0, 0.0, 0.0])
    a[0] += 1
    a[1] += us
    a[2] += fl
    a[3] += by
    tot += us
print(f"total {tot/1000:.3f} ms")
for (kind, desc), (n, us, fl, by) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    tf = fl / us / 1e6 if us > 0 else 0
    gbs = by / us / 1e3 if us > 0 else 0
    ideal = max(fl / (peak_tf * 1e6), by / (peak_gbs * 1e3))
    print(f"{us:9.1f} us {100*us/tot:5.1f}%  n={n:3d} avg={us/n:8.1f}  {tf:7.1f} TF/s {gbs:7.1f} GB/s  ideal={ideal:8.1f}us  x{us/max(ideal,1e-9):5.1f}  {KIND[kind]}: {desc}")
