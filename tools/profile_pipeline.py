"""Where does one txt2img batch go? full process_images vs its UNet calls alone vs the VAE decode alone (CUDA events)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from sdwebui_b200.processing import StableDiffusionProcessingTxt2Img, process_images  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="sd15")
ap.add_argument("--iters", type=int, default=3)
args = ap.parse_args()
dev = torch.device("cuda:0")
w = bench.WORKLOADS[args.config]
dt = torch.bfloat16
B = w["batch"]
model, _ = bench.build_model(w["arch"], dt, dev, 0, 1)
c, u = bench.make_conds(w, B, dev, 7)
c, u = bench.to_dev(c, dev, False), bench.to_dev(u, dev, False)
seeds = list(range(1000, 1000 + B))


def full():
    p = StableDiffusionProcessingTxt2Img(sd_model=model, c=c, uc=u, seeds=seeds, sampler_name=w["sampler"], steps=w["steps"],
                                         cfg_scale=7.0, width=w["width"], height=w["height"], randn_source="GPU")
    return process_images(p, to_host=False)


h, wd = w["height"] // 8, w["width"] // 8
x = torch.randn(2 * B, 4, h, wd, device=dev, dtype=dt)
t = torch.full((2 * B,), 500.0, device=dev, dtype=dt)
ctx = torch.randn(2 * B, 77, w["ctx_dim"], device=dev, dtype=dt)
y = torch.randn(2 * B, w["adm"], device=dev, dtype=dt) if w["adm"] else None
z = torch.randn(B, 4, h, wd, device=dev, dtype=dt)


def unet_only():
    for _ in range(w["steps"]):
        model.unet.engine.forward(x, t, ctx, y)


def vae_only():
    model.vae.decode(z)


def timeit(fn, k):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k


def full_e2e():
    ch, uh = bench.to_dev(c_host, dev), bench.to_dev(u_host, dev)
    p = StableDiffusionProcessingTxt2Img(sd_model=model, c=ch, uc=uh, seeds=seeds, sampler_name=w["sampler"], steps=w["steps"],
                                         cfg_scale=7.0, width=w["width"], height=w["height"], randn_source="GPU")
    return process_images(p, to_host=True)


c_host, u_host = bench.make_conds(w, B, dev, 7)
c_host, u_host = bench.pin(c_host), bench.pin(u_host)
full()
full()
full_e2e()
# resident vs host-buffer call, interleaved so that clock drift under the power cap hits both alike
import time  # noqa: E402
ab = {"resident": [], "e2e": []}
for _ in range(args.iters):
    for name, fn in (("resident", full), ("e2e", full_e2e)):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ab[name].append((time.perf_counter() - t0) * 1e3)
print(f"{args.config}: interleaved wall clock per batch: resident {sorted(ab['resident'])[len(ab['resident']) // 2]:.1f} ms, "
      f"e2e {sorted(ab['e2e'])[len(ab['e2e']) // 2]:.1f} ms  (all: {[round(v, 1) for v in ab['resident']]} vs {[round(v, 1) for v in ab['e2e']]})")
tf, tu, tv = timeit(full, args.iters), timeit(unet_only, args.iters), timeit(vae_only, args.iters)
print(f"{args.config}: process_images {tf:.1f} ms = {w['steps']} UNet calls {tu:.1f} ms + VAE decode {tv:.1f} ms + sampler/glue {tf - tu - tv:.1f} ms")
