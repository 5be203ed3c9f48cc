"""Pins sdwebui_b200/sd_schedulers.py and the Restart sampler to the REFERENCE's in-tree code
(modules/sd_schedulers.py, modules/sd_samplers_extra.py), executed unmodified from /root/reference with a stub `k_diffusion`
(only the three published one-line schedules + to_d) and stub `modules.shared`:

    python tests/golden/make_golden_sched.py   ->   tests/golden/sched_ref.npz
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def toy_model(x, sigma, **kw):
    """a smooth, nonlinear stand-in for the CFG denoiser D(x, sigma)."""
    s = sigma.view(-1, 1, 1, 1)
    return x / (1.0 + s * s) + 0.1 * torch.tanh(x * 0.5) * s / (1.0 + s)


class CountingNoise:
    """randn_like replacement: a fixed, seeded sequence of draws (the product test replays the same sequence)."""

    def __init__(self, shape, seed):
        self.g = torch.Generator().manual_seed(seed)
        self.shape = shape

    def randn_like(self, x):
        return torch.randn(self.shape, generator=self.g)


def main():
    import oracle.kdiffusion as OK
    import sdwebui_b200  # noqa: F401
    from sdwebui_b200 import samplers as S

    def pkg(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        return m

    noise_holder = types.SimpleNamespace(randn_like=None)
    kd_sampling = pkg("k_diffusion.sampling", get_sigmas_karras=OK.get_sigmas_karras, get_sigmas_exponential=OK.get_sigmas_exponential,
                      get_sigmas_polyexponential=OK.get_sigmas_polyexponential, to_d=OK.to_d, torch=noise_holder)
    kd = pkg("k_diffusion", sampling=kd_sampling)
    opts = types.SimpleNamespace(beta_dist_alpha=0.6, beta_dist_beta=0.6)
    sd_model = types.SimpleNamespace(is_sdxl=False)
    shared = pkg("modules.shared", opts=opts, sd_model=sd_model)
    stubs = {"k_diffusion": kd, "k_diffusion.sampling": kd_sampling, "modules": pkg("modules", shared=shared), "modules.shared": shared}
    saved = {k: sys.modules.get(k) for k in stubs}
    sys.modules.update(stubs)
    try:
        mods = {}
        for name, rel in (("sch", "modules/sd_schedulers.py"), ("extra", "modules/sd_samplers_extra.py")):
            spec = importlib.util.spec_from_file_location("ref_" + name, os.path.join(REF, rel))
            mods[name] = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mods[name])
        sch, extra = mods["sch"], mods["extra"]
        # inner model: the product's CompVisDenoiser on the CPU (sigmas table, sigma_to_t, t_to_sigma, get_sigmas)
        inner = S.CompVisDenoiser(types.SimpleNamespace(alphas_cumprod=S.make_alphas_cumprod(), device="cpu"))
        smin, smax = inner.sigmas[0].item(), inner.sigmas[-1].item()
        out = {}
        for n in (5, 11, 20, 31):
            for s in sch.schedulers:
                if s.function is None:
                    continue
                kw = {"sigma_min": smin, "sigma_max": smax}
                if s.need_inner_model:
                    kw["inner_model"] = inner
                for sdxl in ((False, True) if s.name == "align_your_steps" else (False,)):
                    sd_model.is_sdxl = sdxl
                    sig = s.function(n=n, **kw, device="cpu")
                    out[f"{s.name}{'_sdxl' if sdxl else ''}_{n}"] = np.asarray(sig, dtype=np.float64)
        # Restart sampler on a toy model: 3 step counts exercise no restart (< 20), one restart segment, two segments
        for steps in (10, 24, 40):
            sigmas = OK.get_sigmas_karras(steps, smin, smax, 7.0, "cpu")
            x0 = torch.randn(2, 4, 8, 8, generator=torch.Generator().manual_seed(steps)) * sigmas[0]
            noise_holder.randn_like = CountingNoise((2, 4, 8, 8), 100 + steps).randn_like
            calls = []
            res = extra.restart_sampler(toy_model, x0.clone(), sigmas, callback=lambda d: calls.append(float(d["sigma_hat"])), disable=True)
            out[f"restart_{steps}"] = res.numpy()
            out[f"restart_{steps}_sigmas"] = np.array(calls)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    path = os.path.join(HERE, "sched_ref.npz")
    np.savez_compressed(path, **out)
    print(path, len(out), "arrays", os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
