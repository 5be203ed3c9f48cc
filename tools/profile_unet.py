"""Profiling driver: N UNet forwards (SD1.5 or SDXL, CFG batch) and optionally one VAE decode, for ncu.
  ncu --metrics gpu__time_duration.sum --clock-control none -s <skip> -c <n> --csv --log-file gpurun_out/launches.csv \
      python tools/profile_unet.py --config sd15 --iters 2
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import sdwebui_b200  # noqa: E402,F401
from sdwebui_b200 import checkpoint as C  # noqa: E402
from sdwebui_b200.engine import UNetEngine, UNetSpec, VAEDecoderEngine, VAESpec  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="sd15")
ap.add_argument("--iters", type=int, default=2)
ap.add_argument("--batch", type=int, default=0)
ap.add_argument("--dtype", default="bf16")
ap.add_argument("--vae", action="store_true")
ap.add_argument("--vae-only", action="store_true", help="skip the UNet; profile one VAE decode of n/2 latents")
ap.add_argument("--profile", action="store_true", help="print the engine's own per-class event timing")
args = ap.parse_args()
dev = torch.device("cuda:0")
dt = torch.bfloat16 if args.dtype == "bf16" else torch.float16
spec = UNetSpec.sd15() if args.config == "sd15" else UNetSpec.sdxl()
n = args.batch or (16 if args.config == "sd15" else 8)
hw = 64 if args.config == "sd15" else 128
def report(e):
    for k, v in e.profile_read().items():
        if v["launches"]:
            print(f"{k:12s} launches {v['launches']:5d}  ms/iter {v['ms'] / args.iters:8.3f}  TFLOP/s {v['flops'] / max(v['ms'], 1e-9) / 1e9:8.1f}  GB/s {v['bytes'] / max(v['ms'], 1e-9) / 1e6:8.1f}")


if args.vae_only:
    vae = VAEDecoderEngine(VAESpec(), dtype=dt, device=dev)
    vae.load_state_dict(C.synthetic_state_dict(C.vae_decoder_param_shapes(VAESpec()), 1, device=dev, dtype=torch.float16))
    vae.finalize()
    z = torch.randn(n // 2, 4, hw, hw, device=dev, dtype=dt)
    vae.decode(z)
    if args.profile:
        vae.profile(True)
    for _ in range(args.iters):
        vae.decode(z)
    torch.cuda.synchronize()
    if args.profile:
        report(vae)
    print("done")
    sys.exit(0)
eng = UNetEngine(spec, dtype=dt, device=dev)
eng.load_state_dict(C.synthetic_state_dict(C.unet_param_shapes(spec), 0, device=dev, dtype=torch.float16))
eng.finalize()
x = torch.randn(n, 4, hw, hw, device=dev, dtype=dt)
t = torch.full((n,), 500.0, device=dev, dtype=dt)
ctx = torch.randn(n, 77, spec.context_dim, device=dev, dtype=dt)
y = torch.randn(n, spec.adm_in_channels, device=dev, dtype=dt) if spec.adm_in_channels else None
if args.profile:
    eng.forward(x, t, ctx, y)
    eng.profile(True)
if not args.profile:
    for _ in range(2):  # plan build + graph capture happen on the first call of a shape
        eng.forward(x, t, ctx, y)
    torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(args.iters):
    eng.forward(x, t, ctx, y)
ev1.record()
torch.cuda.synchronize()
print(f"unet forward: {ev0.elapsed_time(ev1) / args.iters:.3f} ms/iter ({'eager+events' if args.profile else 'graph replay'})")
if args.profile:
    report(eng)
if args.vae:
    vae = VAEDecoderEngine(VAESpec(), dtype=dt, device=dev)
    vae.load_state_dict(C.synthetic_state_dict(C.vae_decoder_param_shapes(VAESpec()), 1, device=dev, dtype=torch.float16))
    vae.finalize()
    z = torch.randn(n // 2, 4, hw, hw, device=dev, dtype=dt)
    for _ in range(args.iters):
        vae.decode(z)
    torch.cuda.synchronize()
print("done")
