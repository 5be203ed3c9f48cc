"""Prints the attention2 timeline trace (gpurun_out/attn_trace.txt; build with `make ATT_TRACE=1`, run with
SDXE_ATT_TRACE_DUMP=1). Roles 0/1: softmax warp (half 0, quarter 0) of tile A/B; roles 2/3: MMA issuer of tile A/B."""
import re
import sys

d = {}
for ln in open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/attn_trace.txt"):
    m = re.match(r"role (\d) it\s+(\d+):(.*)", ln)
    d[(int(m.group(1)), int(m.group(2)))] = [int(x) for x in m.group(3).split()]
lo, hi = 8, 16
print("softmax: wait_start | +s_full +pass_done +pv_done_seen +arrived | period, offset of tile B vs A")
for t in (0, 1):
    for it in range(lo, hi):
        r, prev = d[(t, it)], d[(t, it - 1)]
        extra = f" exp_start {r[5] - r[0]} exp_loop {r[3] - r[5]}" if r[5] > 0 else ""
        print(f" T{t} {it:2d}", [r[k] - r[0] for k in range(1, 5)], "period", r[0] - prev[0], "" if t == 0 else f"offset {r[0] - d[(0, it)][0]}", extra)
print("MMA issuer: loop_top | +pops +p_ready +S_issued +PV_issued +released | period")
for t in (0, 1):
    for it in range(lo, hi):
        r, prev = d[(2 + t, it)], d[(2 + t, it - 1)]
        print(f" M{t} {it:2d}", [r[k] - r[0] for k in range(1, 6)], "period", r[0] - prev[0])
