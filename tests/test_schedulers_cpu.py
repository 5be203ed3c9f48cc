"""sdwebui_b200/sd_schedulers.py == the reference's modules/sd_schedulers.py on the same inputs (tests/golden/sched_ref.npz,
written by tests/golden/make_golden_sched.py which executes the reference file), for every schedule and several step counts;
KDiffusionSampler.get_sigmas option handling (discard_next_to_last_sigma, sigma overrides, rho)."""
import os
import types

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "sched_ref.npz"))


def _inner():
    from sdwebui_b200 import samplers as S

    return S.CompVisDenoiser(types.SimpleNamespace(alphas_cumprod=S.make_alphas_cumprod(), device="cpu"))


def test_every_schedule_matches_reference():
    from sdwebui_b200 import sd_schedulers as SS

    inner = _inner()
    smin, smax = inner.sigmas[0].item(), inner.sigmas[-1].item()
    checked = 0
    for key in GOLD.files:
        if key.startswith("restart"):
            continue
        name, n = key.rsplit("_", 1)
        sdxl = name.endswith("_sdxl")
        name = name[:-5] if sdxl else name
        s = SS.schedulers_map[name]
        kw = {"sigma_min": smin, "sigma_max": smax}
        if s.need_inner_model:
            kw["inner_model"] = inner
        if name == "align_your_steps":
            kw["is_sdxl"] = sdxl
        sig = s.function(n=int(n), **kw, device="cpu")
        ref = GOLD[key]
        assert sig.shape[0] == ref.shape[0], key
        assert np.allclose(np.asarray(sig, dtype=np.float64), ref, rtol=2e-6, atol=1e-7), key
        checked += 1
    assert checked >= 40
    assert [s.name for s in SS.schedulers] == ["automatic", "uniform", "karras", "exponential", "polyexponential", "sgm_uniform",
                                               "kl_optimal", "align_your_steps", "simple", "normal", "ddim", "beta"]
    assert SS.schedulers_map["SGM Uniform"] is SS.schedulers_map["sgm_uniform"]


def test_get_sigmas_options():
    from sdwebui_b200 import samplers as S
    from sdwebui_b200 import sd_schedulers as SS

    sd_model = types.SimpleNamespace(alphas_cumprod=S.make_alphas_cumprod(), device="cpu", is_sdxl=False)
    inner = _inner()
    smin, smax = inner.sigmas[0].item(), inner.sigmas[-1].item()
    p = types.SimpleNamespace(scheduler="Automatic", is_hr_pass=False)
    # DPM2: karras by default, asks for one more step and drops the penultimate sigma (:80-84,130-131)
    k = S.KDiffusionSampler("DPM2", sd_model)
    sig = k.get_sigmas(p, 10)
    full = SS.get_sigmas_karras(11, smin, smax, 7.0)
    assert sig.shape[0] == 11 and torch.allclose(sig, torch.cat([full[:-2], full[-1:]]))
    # Euler a: no default scheduler -> the model's own uniform-in-t schedule
    e = S.KDiffusionSampler("Euler a", sd_model)
    assert torch.allclose(e.get_sigmas(p, 7), inner.get_sigmas(7))
    # explicit scheduler + option overrides
    p.scheduler = "Exponential"
    assert torch.allclose(e.get_sigmas(p, 9), SS.get_sigmas_exponential(9, smin, smax))
    p.scheduler = "Karras"
    e.sched_opts.rho, e.sched_opts.sigma_min, e.sched_opts.sigma_max = 5.0, 0.05, 12.0
    assert torch.allclose(e.get_sigmas(p, 9), SS.get_sigmas_karras(9, 0.05, 12.0, 5.0))
    e.sched_opts.always_discard_next_to_last_sigma = True
    assert e.get_sigmas(p, 9).shape[0] == 10
    with pytest.raises(Exception):
        p.scheduler = "nonsense"
        e.get_sigmas(p, 5)
    assert sorted(x[0] for x in S.samplers_k_diffusion) == sorted(["DPM++ 2M", "DPM++ 2S a", "Euler a", "Euler", "LMS", "Heun", "DPM2", "DPM2 a", "Restart"])
