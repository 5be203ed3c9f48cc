"""GEGLU-epilogue GEMM (the FeedForward up-projection of every transformer block) at the SD1.5 / SDXL shapes: median of 10
launches, L2 flushed between them, CUDA events.   gpurun -- 'python tools/bench_geglu.py'"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import sdwebui_b200  # noqa: E402,F401
from sdwebui_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
dt = torch.bfloat16
once = len(sys.argv) > 1 and sys.argv[1] == "once"   # one launch per shape (for an ncu capture)
shapes = [(65536, 2560, 320), (16384, 5120, 640), (4096, 10240, 1280), (32768, 5120, 640), (8192, 10240, 1280)]
probe = len(sys.argv) > 1 and sys.argv[1] == "probe"  # order / repeat sensitivity: every timing is printed
if probe:
    shapes = [(16384, 5120, 640), (65536, 2560, 320), (65536, 2560, 320), (65536, 2560, 640), (32768, 2560, 320), (65536, 1280, 320)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for (M, N, K) in shapes[:1] if once else shapes:
    a = torch.randn(M, K, device=dev).to(dt)
    w = (torch.randn(N, K, device=dev) / K ** 0.5).to(dt)
    b = torch.randn(N, device=dev)
    for _ in range(1 if once else 3):
        out = ops.gemm(a, w, b, None, geglu=True)
    if once:
        torch.cuda.synchronize()
        break
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.gemm(a, w, b, None, geglu=True)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    us = ts[len(ts) // 2]
    h = (a.float() @ w.float().t() + b)
    ref = h[:, : N // 2] * torch.nn.functional.gelu(h[:, N // 2:])
    err = (out.float() - ref).abs().max().item() / ref.abs().max().item()
    if probe:
        print("  all:", " ".join(f"{t:.0f}" for t in ts))
    print(f"geglu M={M} N={N} K={K}: {us:.1f} us  {2.0 * M * N * K / us / 1e6:.0f} TFLOP/s  max err / max |ref| {err:.2e}")
    del h, ref
