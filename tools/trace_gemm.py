"""Runs one sdxe_gemm shape a few times; with a `make GEMM_TRACE=1` build and SDXE_GEMM_TRACE_DUMP=<n> the n-th launch
writes gpurun_out/gemm_trace.txt (CTA 0's producer / MMA / epilogue timeline).  python tools/trace_gemm.py M N K mode"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import sdwebui_b200  # noqa: E402,F401
from sdwebui_b200 import ops  # noqa: E402

M, N, K = (int(v) for v in sys.argv[1:4])
mode = sys.argv[4] if len(sys.argv) > 4 else "plain"
dev, dt = torch.device("cuda:0"), torch.bfloat16
a = torch.randn(M, K, device=dev).to(dt)
w = (torch.randn(N, K, device=dev) / K ** 0.5).to(dt)
b = torch.randn(N, device=dev)
r = torch.randn(M, N, device=dev).to(dt) if mode == "res" else None
for _ in range(8):
    ops.gemm(a, w, b, r, geglu=(mode == "geglu"))
torch.cuda.synchronize()
