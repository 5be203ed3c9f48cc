"""Per-tile timeline of CTA 0 from gpurun_out/gemm_trace.txt (clock64, relative to the producer's first stamp)."""
import sys

rows = {}
hdr = ""
for line in open(sys.argv[1]):
    if line.startswith("#"):
        hdr = line.strip()
        continue
    r, t, *ev = (int(v) for v in line.split())
    rows[(r, t)] = ev
t0 = min(v for ev in rows.values() for v in ev if v)
print(hdr)
print("tile | producer: start  slot-free  issued | MMA: begin  acc-free  first-full  committed | epi first warp: ready-to-wait  acc-full  done | epi last warp: wait full done")
prev_commit = prev_done = None
for t in range(40):
    p, m, e0, e1 = rows.get((0, t)), rows.get((1, t)), rows.get((2, t)), rows.get((3, t))
    if not m or not m[3]:
        break
    f = lambda v: f"{v - t0:7d}" if v else "      -"
    extra = ""
    if prev_commit is not None:
        extra = f"  | tile period {m[3] - prev_commit:6d}  epilogue {e0[2] - e0[1]:6d}  mainloop {m[3] - m[2]:6d}"
    print(f"{t:3d} | {f(p[0])} {f(p[1])} {f(p[2])} | {f(m[0])} {f(m[1])} {f(m[2])} {f(m[3])} | {f(e0[0])} {f(e0[1])} {f(e0[2])} | {f(e1[0])} {f(e1[1])} {f(e1[2])}{extra}")
    prev_commit = m[3]

if (4, 0) in rows and any(rows[(4, 0)]):
    print("first epilogue warp, first chunk of each tile (cycles): tcgen05.ld+wait | arithmetic (+residual wait) | buffer-free wait | st.shared | proxy fence | TMA store issue")
    for t in range(2, 12):
        a, b = rows.get((4, t)), rows.get((5, t))
        if not a or not a[3] or not b[2]:
            break
        print(f"{t:3d} | {a[1] - a[0]:6d} {a[2] - a[1]:6d} {a[3] - a[2]:6d} {b[0] - a[3]:6d} {b[1] - b[0]:6d} {b[2] - b[1]:6d}   chunk total {b[2] - a[0]:6d}")
