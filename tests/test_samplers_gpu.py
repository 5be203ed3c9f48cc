"""The samplers added for row N4 (Euler, Heun, DPM2, DPM2 a, DPM++ 2S a, LMS, Restart): their step updates run as fused
sdxe_lincomb launches with host-side scalars. Checked on the GPU against (a) the oracle's torch restatement of the published
k-diffusion algorithms on a toy denoiser (fp32 both: only the association order differs), (b) for Restart, the REFERENCE's own
modules/sd_samplers_extra.py output (tests/golden/sched_ref.npz), and (c) end to end through process_images with the tiny UNet."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def toy_model(x, sigma, **kw):
    s = sigma.view(-1, 1, 1, 1)
    return x / (1.0 + s * s) + 0.1 * torch.tanh(x * 0.5) * s / (1.0 + s)


class Seq:
    def __init__(self, shape, seed, device):
        self.g, self.shape, self.device = torch.Generator().manual_seed(seed), shape, device

    def __call__(self, *a):
        return torch.randn(self.shape, generator=self.g).to(self.device)


@pytest.mark.parametrize("name", ["sample_euler", "sample_heun", "sample_dpm_2", "sample_dpm_2_ancestral", "sample_dpmpp_2s_ancestral", "sample_lms"])
def test_sampler_matches_oracle_on_toy_model(cuda, name):
    import oracle.kdiffusion as OK
    from sdwebui_b200 import samplers as S

    sigmas = OK.get_sigmas_karras(12, 0.0292, 14.6146, 7.0, "cpu")
    x0 = (torch.randn(3, 4, 16, 16, generator=torch.Generator().manual_seed(1)) * sigmas[0]).to(cuda)
    kw_o, kw_p = {}, {}
    if name != "sample_lms":
        kw_o["noise_sampler"] = Seq((3, 4, 16, 16), 9, cuda)
        kw_p["noise_sampler"] = Seq((3, 4, 16, 16), 9, cuda)
    if name in ("sample_euler", "sample_heun", "sample_dpm_2"):
        kw_o["s_churn"] = kw_p["s_churn"] = 3.0   # exercises the sigma_hat / churn-noise branch
    want = getattr(OK, name)(toy_model, x0.clone(), sigmas.to(cuda), **kw_o)
    got = getattr(S, name)(toy_model, x0.clone(), sigmas, **kw_p)
    err = ((got - want).norm() / want.norm()).item()
    print(f"{name}: rel err {err:.2e}")
    assert err < 2e-6


def test_restart_sampler_matches_reference(cuda):
    import oracle.kdiffusion as OK
    from sdwebui_b200 import samplers as S

    gold = np.load(os.path.join(HERE, "golden", "sched_ref.npz"))
    for steps in (10, 24, 40):
        sigmas = OK.get_sigmas_karras(steps, 0.029167532920837402, 14.614642143249512, 7.0, "cpu")
        x0 = torch.randn(2, 4, 8, 8, generator=torch.Generator().manual_seed(steps)) * sigmas[0]
        calls = []
        got = S.restart_sampler(toy_model, x0.to(cuda), sigmas, callback=lambda d: calls.append(float(d["sigma_hat"])),
                                noise_sampler=Seq((2, 4, 8, 8), 100 + steps, cuda))
        ref = torch.from_numpy(gold[f"restart_{steps}"])
        err = ((got.cpu() - ref).norm() / ref.norm()).item()
        print(f"restart {steps} steps: {len(calls)} denoiser steps, rel err vs reference {err:.2e}")
        assert np.allclose(np.array(calls), gold[f"restart_{steps}_sigmas"], rtol=1e-6)   # same step list incl. restart segments
        assert err < 5e-6


@pytest.mark.parametrize("sampler,steps", [("Heun", 5), ("DPM2 a", 6), ("DPM++ 2S a", 5), ("LMS", 6), ("Euler", 6)])
def test_tiny_txt2img_other_samplers(cuda, sampler, steps):
    from oracle.pipeline import OraclePipeline, SamplingParams
    from oracle.synth import init_module_, synthetic_context
    from oracle.unet import UNetModel, tiny_config
    from oracle.vae import AutoencoderKLDecode, tiny_vae_config
    from sdwebui_b200.engine import UNetSpec, VAEDecoderEngine, VAESpec
    from sdwebui_b200.processing import SdModel, StableDiffusionProcessingTxt2Img, process_images
    from sdwebui_b200.sd_unet import SdxeUnet

    ucfg, vcfg = tiny_config(), tiny_vae_config()
    unet = init_module_(UNetModel(ucfg), 1).eval().to(cuda)
    vae = init_module_(AutoencoderKLDecode(vcfg), 2).eval().to(cuda)
    su = SdxeUnet(unet.state_dict(), UNetSpec.from_any(ucfg), dtype=torch.float16, device=cuda)
    su.activate()
    ve = VAEDecoderEngine(VAESpec.from_any(vcfg), dtype=torch.float16, device=cuda)
    ve.load_state_dict(vae.state_dict())
    ve.finalize()
    model = SdModel(su, ve, is_sdxl=False, dtype_unet=torch.float16, device=cuda)
    B, seeds = 2, (1000, 1001)
    cond, uncond = synthetic_context(B, 77, ucfg.context_dim, 3, cuda), synthetic_context(B, 77, ucfg.context_dim, 4, cuda)
    sp = SamplingParams(sampler=sampler, steps=steps, width=128, height=128, seeds=seeds, randn_source="NV")
    lat32, _ = OraclePipeline(unet, vae, cuda, dtype_unet=torch.float32).txt2img(sp, cond, uncond)
    p = StableDiffusionProcessingTxt2Img(sd_model=model, c=cond, uc=uncond, seeds=list(seeds), sampler_name=sampler, steps=steps,
                                         width=128, height=128, randn_source="NV", do_not_decode=True)
    res = process_images(p)
    err = ((res.latents - lat32).norm() / lat32.norm()).item()
    print(f"tiny {sampler}: latent rel err vs fp32 oracle {err:.3e}")
    assert err < 1.5e-2
    su.deactivate()
    ve.close()
