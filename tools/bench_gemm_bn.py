"""Tile-width sweep of sdxe_gemm for a few problem shapes (input to the gemm_pick_bn cost model).
   gpurun -- 'python tools/bench_gemm_bn.py'"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import sdwebui_b200  # noqa: E402,F401
from sdwebui_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
dt = torch.bfloat16
shapes = [(8192, 1280, 1280, True), (8192, 1280, 5120, True), (8192, 3840, 1280, False), (16384, 640, 640, True),
          (4096, 1280, 1280, True), (65536, 320, 320, True), (32768, 640, 640, True)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for (M, N, K, res) in shapes:
    a = torch.randn(M, K, device=dev).to(dt)
    w = (torch.randn(N, K, device=dev) / K ** 0.5).to(dt)
    b = torch.randn(N, device=dev)
    r = torch.randn(M, N, device=dev).to(dt) if res else None
    row = []
    for bn in (0, 96, 112, 128, 144, 160, 192, 208, 224, 240, 256):
        if bn and N % 8:
            continue
        for _ in range(3):
            ops.gemm(a, w, b, r, force_bn=bn)
        ts = []
        for _ in range(10):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.gemm(a, w, b, r, force_bn=bn)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort()
        row.append((bn, ts[len(ts) // 2]))
    best = min(row[1:], key=lambda x: x[1])
    print(f"M={M} N={N} K={K} res={int(res)}: auto {row[0][1]:.1f} us | " + " ".join(f"{bn}:{t:.1f}" for bn, t in row[1:]) + f" | best {best[0]} ({best[1]:.1f} us)")
