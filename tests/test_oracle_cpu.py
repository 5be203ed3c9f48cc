"""CPU tests of the oracle: everything the reference can pin, is pinned here (see oracle/__init__.py)."""
import math
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_philox_known_answer_vector():
    """modules/rng_philox.py:5-15 — the only known-answer vector in the reference tree."""
    from oracle.rng import PhiloxGenerator

    expect = np.array([[-0.92466259, -0.42534415, -2.6438457, 0.14518388],
                       [-0.12086647, -0.57972564, -0.62285122, -0.32838709],
                       [-1.07454231, -0.36314407, -1.67105067, 2.26550497]])
    got = PhiloxGenerator(0).randn((3, 4))
    assert np.allclose(got, expect, atol=2e-6)


@pytest.mark.parametrize("impl", ["oracle", "product"])
def test_philox_matches_reference_module(impl):
    """bit-for-bit against outputs of the reference's own module (fixture made by tests/golden/make_golden.py)."""
    if impl == "oracle":
        from oracle.rng import PhiloxGenerator as G
    else:
        from sdwebui_b200.rng import PhiloxGenerator as G
    ref = np.load(os.path.join(GOLD, "philox_ref.npz"))
    for seed in (0, 1000, 123456789012, 2 ** 32 + 5):
        g = G(seed)
        for call in range(3):
            assert np.array_equal(g.randn((4, 8, 8)), ref[f"s{seed}_c{call}"]), (seed, call)
    assert np.array_equal(G(0).randn((3, 4)), ref["doc_3x4"])


def test_parameter_counts():
    """SURVEY §8(c) structural KATs."""
    from oracle.unet import UNetModel, sd15_config, sdxl_config
    from oracle.vae import AutoencoderKLDecode, VAEConfig

    with torch.device("meta"):
        assert sum(p.numel() for p in UNetModel(sd15_config()).parameters()) == 859_520_964
        assert sum(p.numel() for p in UNetModel(sdxl_config()).parameters()) == 2_567_463_684
        vae = AutoencoderKLDecode(VAEConfig())
        assert sum(p.numel() for p in vae.decoder.parameters()) == 49_490_179
        assert sum(p.numel() for p in vae.post_quant_conv.parameters()) == 20


def test_state_dict_key_layout():
    """ldm key names the reference itself relies on (extensions-builtin/Lora/networks.py:43-98, sd_hijack.py:191-203)."""
    from oracle.unet import UNetModel, sd15_config, sdxl_config

    with torch.device("meta"):
        k15 = set(UNetModel(sd15_config()).state_dict())
        kxl = set(UNetModel(sdxl_config()).state_dict())
    for k in ("input_blocks.3.0.op.weight", "input_blocks.6.0.op.weight", "input_blocks.9.0.op.weight",
              "output_blocks.2.1.conv.weight", "output_blocks.5.2.conv.weight", "output_blocks.8.2.conv.weight",
              "input_blocks.1.0.in_layers.2.weight", "input_blocks.1.0.emb_layers.1.weight", "input_blocks.1.0.out_layers.3.weight",
              "input_blocks.4.0.skip_connection.weight", "middle_block.1.transformer_blocks.0.attn2.to_k.weight",
              "input_blocks.1.1.transformer_blocks.0.ff.net.0.proj.weight", "out.2.weight", "time_embed.2.bias"):
        assert k in k15, k
    assert "label_emb.0.0.weight" in kxl and "label_emb.0.0.weight" not in k15
    assert "input_blocks.8.1.transformer_blocks.9.attn1.to_q.weight" in kxl      # depth-10 level
    assert "input_blocks.4.1.transformer_blocks.1.norm3.weight" in kxl           # depth-2 level
    assert not any(k.startswith("input_blocks.1.1") for k in kxl)               # no attention at level 0
    assert "output_blocks.2.2.conv.weight" in kxl and "output_blocks.5.2.conv.weight" in kxl


def test_sigma_table_endpoints():
    """modules/sd_schedulers.py:59-63 quotes sigma_max 14.615 / sigma_min 0.029 for these models."""
    from oracle import kdiffusion as K

    s = K.DiscreteSchedule(K.make_alphas_cumprod())
    assert abs(s.sigmas[-1].item() - 14.615) < 2e-3
    assert abs(s.sigmas[0].item() - 0.029) < 5e-4
    sig = s.get_sigmas(20)
    assert sig.shape == (21,) and sig[-1] == 0 and abs(sig[0].item() - s.sigmas[-1].item()) < 1e-6
    assert torch.all(sig[:-1][1:] < sig[:-1][:-1])
    kar = K.get_sigmas_karras(30, s.sigmas[0].item(), s.sigmas[-1].item())
    assert kar.shape == (31,) and abs(kar[0].item() - 14.6146) < 1e-3 and abs(kar[-2].item() - 0.0292) < 1e-4
    # sigma_to_t inverts t_to_sigma on the grid and interpolates in log space in between
    t = torch.tensor([0.0, 10.0, 500.5, 998.0])
    assert torch.allclose(s.sigma_to_t(s.t_to_sigma(t)), t, atol=2e-3)


def test_img2img_step_arithmetic():
    """modules/sd_samplers_common.py:22-31: hires 20 steps at denoise 0.75 -> steps 26, t_enc 19."""
    from oracle import kdiffusion as K

    assert K.setup_img2img_steps(20, 0.75) == (26, 19)
    assert K.setup_img2img_steps(20, 1.0) == (20, 19)


def test_samplers_on_analytic_denoiser():
    """With the ideal denoiser of a point mass at 0 (D(x, sigma) = 0): Euler-a with eta=0 lands exactly on
    x * sigma_last/sigma_0 after each step, DPM++ 2M gives x -> (sigma_next/sigma) x."""
    from oracle import kdiffusion as K

    sig = K.get_sigmas_karras(8, 0.03, 14.6)
    x0 = torch.randn(2, 4, 8, 8)
    zero = lambda x, s, **kw: torch.zeros_like(x)  # noqa: E731
    out = K.sample_euler_ancestral(zero, x0.clone(), sig, eta=0.0, noise_sampler=lambda: torch.zeros_like(x0))
    assert out.abs().max() < 1e-5  # last sigma is 0
    out = K.sample_dpmpp_2m(zero, x0.clone(), sig[:-1])
    assert torch.allclose(out, x0 * (sig[-2] / sig[0]), rtol=1e-4, atol=1e-6)


def test_combine_denoised():
    """sd_samplers_cfg_denoiser.py:74-82: u + (c - u) * scale."""
    from oracle.cfg_denoiser import combine_denoised

    x_out = torch.randn(6, 4, 2, 2)
    d = combine_denoised(x_out, [[(0, 1.0)], [(1, 1.0)], [(2, 1.0)]], 3, 7.0)
    assert torch.allclose(d, x_out[3:] + (x_out[:3] - x_out[3:]) * 7.0, atol=1e-6)


def test_tiny_oracle_regression_fixture():
    """Oracle self-consistency against tests/golden/tiny_oracle.npz (guards against accidental edits of the oracle)."""
    from oracle.pipeline import OraclePipeline, SamplingParams
    from oracle.synth import init_module_, synthetic_context
    from oracle.unet import UNetModel, tiny_config
    from oracle.vae import AutoencoderKLDecode, tiny_vae_config

    gold = np.load(os.path.join(GOLD, "tiny_oracle.npz"))
    ucfg, vcfg = tiny_config(), tiny_vae_config()
    unet = init_module_(UNetModel(ucfg), 1).eval()
    vae = init_module_(AutoencoderKLDecode(vcfg), 2).eval()
    c, u = synthetic_context(2, 77, ucfg.context_dim, 3), synthetic_context(2, 77, ucfg.context_dim, 4)
    x = torch.randn(2, 4, 16, 16, generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        assert np.allclose(unet(x, torch.tensor([801.0, 37.5]), context=c).numpy(), gold["unet_eps"], atol=2e-4)
        assert np.allclose(vae.decode(x[:1]).numpy(), gold["vae_img"], atol=2e-4)
    pipe = OraclePipeline(unet, vae, "cpu")
    lat = pipe.sample(SamplingParams(sampler="Euler a", steps=4, width=128, height=128, seeds=(1000, 1001), randn_source="NV"), c, u)
    assert np.allclose(lat.numpy(), gold["latent_euler_a"], rtol=2e-3, atol=2e-3)
    lat = pipe.sample(SamplingParams(sampler="DPM++ 2M", steps=5, width=128, height=128, seeds=(1000, 1001), randn_source="NV"), c, u)
    assert np.allclose(lat.numpy(), gold["latent_dpmpp_2m"], rtol=2e-3, atol=2e-3)


def test_config1_plumbing_cpu():
    """BASELINE config 1 shape: txt2img, 1 Euler step, batch 1, CPU fp32 — on the tiny architecture so it runs in
    seconds (the full SD1.5 CPU run is bench.py's cpu_baseline leg)."""
    from oracle.pipeline import OraclePipeline, SamplingParams
    from oracle.synth import init_module_, synthetic_context
    from oracle.unet import UNetModel, tiny_config
    from oracle.vae import AutoencoderKLDecode, tiny_vae_config

    ucfg = tiny_config()
    pipe = OraclePipeline(init_module_(UNetModel(ucfg), 1).eval(), init_module_(AutoencoderKLDecode(tiny_vae_config()), 2).eval(), "cpu")
    c, u = synthetic_context(1, 77, ucfg.context_dim, 3), synthetic_context(1, 77, ucfg.context_dim, 4)
    lat, img = pipe.txt2img(SamplingParams(sampler="Euler a", steps=1, width=64, height=64, seeds=(1,)), c, u)
    assert lat.shape == (1, 4, 8, 8) and img.shape[1] == 3 and torch.isfinite(img).all() and 0 <= img.min() and img.max() <= 1
