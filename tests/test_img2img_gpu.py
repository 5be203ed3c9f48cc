"""SURVEY §8(f) N1 on the GPU: VAE encoder (engine vs oracle) and the img2img latent path (init image -> encode ->
noise at denoising strength -> sampler -> optional latent mask blend) end to end against the oracle pipeline.
Same yardstick as test_engine_gpu.py: error vs the fp32 oracle below max(3 x the reference 16-bit path's error, floor)."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


def rel_err(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


def _floor(dtype):
    return 2e-3 if dtype == torch.float16 else 1.6e-2


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_tiny_vae_encoder(cuda, dtype):
    from oracle.synth import init_module_
    from oracle.vae import AutoencoderKLEncode, tiny_vae_config
    from sdwebui_b200.engine import VAEEncoderEngine, VAESpec

    cfg = tiny_vae_config()
    enc = init_module_(AutoencoderKLEncode(cfg), 33).eval().to(cuda)
    eng = VAEEncoderEngine(VAESpec.from_any(cfg), dtype=dtype, device=cuda)
    eng.load_state_dict(enc.state_dict())
    assert eng.param_count() == sum(p.numel() for p in enc.parameters())
    eng.finalize()
    g = torch.Generator(device="cuda").manual_seed(9)
    f = 2 ** (len(cfg.ch_mult) - 1)
    for (n, h, w) in [(1, 32, 32), (2, 64, 48), (1, 128, 128)]:
        x = torch.rand(n, 3, h, w, device=cuda, generator=g) * 2 - 1
        with torch.no_grad():
            ref32 = enc.encode_moments(x.to(dtype).float())
            ref16 = copy.deepcopy(enc).to(dtype).encode_moments(x.to(dtype))
        out = eng.encode_moments(x.to(dtype))
        assert out.shape == (n, 2 * cfg.z_channels, h // f, w // f)
        e_eng, e_ref = rel_err(out, ref32), rel_err(ref16, ref32)
        print(f"tiny vae encoder {dtype} {n}x{h}x{w}: engine {e_eng:.3e} ref16 {e_ref:.3e}")
        assert e_eng < max(3 * e_ref, _floor(dtype)), (e_eng, e_ref)
    with pytest.raises(Exception):
        eng.encode_moments(torch.zeros(1, 3, 33, 32, device=cuda, dtype=dtype))  # not a multiple of the downsampling factor
    eng.close()


def test_full_vae_encoder(cuda):
    """KL-f8 encoder + quant_conv (34,163,592 + 72 parameters), one 256x256 image -> 32x32 moments."""
    from oracle.synth import init_module_
    from oracle.vae import AutoencoderKLEncode, VAEConfig
    from sdwebui_b200.engine import VAEEncoderEngine, VAESpec

    dtype = torch.float16
    cfg = VAEConfig()
    enc = init_module_(AutoencoderKLEncode(cfg), 43).eval().to(cuda)
    eng = VAEEncoderEngine(VAESpec.from_any(cfg), dtype=dtype, device=cuda)
    eng.load_state_dict(enc.state_dict())
    assert eng.param_count() == 34163592 + 72
    eng.finalize()
    g = torch.Generator(device="cuda").manual_seed(10)
    x = torch.rand(1, 3, 256, 256, device=cuda, generator=g) * 2 - 1
    with torch.no_grad():
        ref32 = enc.encode_moments(x.to(dtype).float())
        ref16 = copy.deepcopy(enc).to(dtype).encode_moments(x.to(dtype))
    out = eng.encode_moments(x.to(dtype))
    assert out.shape == (1, 8, 32, 32)
    e_eng, e_ref = rel_err(out, ref32), rel_err(ref16, ref32)
    print(f"full vae encoder fp16: engine {e_eng:.3e} ref16 {e_ref:.3e}")
    assert e_eng < max(3 * e_ref, _floor(dtype)), (e_eng, e_ref)
    eng.close()


@pytest.mark.parametrize("masked", [False, True])
def test_tiny_img2img(cuda, masked):
    from oracle.pipeline import OraclePipeline, SamplingParams
    from oracle.synth import init_module_
    from oracle.unet import UNetModel, tiny_config
    from oracle.vae import AutoencoderKLDecode, AutoencoderKLEncode, tiny_vae_config
    from sdwebui_b200.engine import UNetSpec, VAEDecoderEngine, VAEEncoderEngine, VAESpec
    from sdwebui_b200.processing import SdModel, StableDiffusionProcessingImg2Img, process_images
    from sdwebui_b200.sd_unet import SdxeUnet

    dtype = torch.float16
    ucfg = tiny_config()
    vcfg = tiny_vae_config()
    vcfg.ch_mult = [1, 2, 2, 2]  # f = 8 like the real VAE so that the latent is H/8 x W/8
    unet = init_module_(UNetModel(ucfg), 51).eval().to(cuda)
    dec = init_module_(AutoencoderKLDecode(vcfg), 52).eval().to(cuda)
    enc = init_module_(AutoencoderKLEncode(vcfg), 53).eval().to(cuda)
    su = SdxeUnet(unet.state_dict(), spec=UNetSpec.from_any(ucfg), dtype=dtype, device=str(cuda))
    su.activate()
    vd = VAEDecoderEngine(VAESpec.from_any(vcfg), dtype=dtype, device=cuda)
    vd.load_state_dict(dec.state_dict()); vd.finalize()
    ve = VAEEncoderEngine(VAESpec.from_any(vcfg), dtype=dtype, device=cuda)
    ve.load_state_dict(enc.state_dict()); ve.finalize()
    model = SdModel(su, vd, is_sdxl=False, dtype_unet=dtype, device=str(cuda), vae_encoder=ve)

    B, H, W = 2, 128, 128
    g = torch.Generator(device="cuda").manual_seed(12)
    init = torch.rand(B, 3, H, W, device=cuda, generator=g)
    cond = torch.randn(B, 77, ucfg.context_dim, device=cuda, generator=g)
    uncond = torch.randn(B, 77, ucfg.context_dim, device=cuda, generator=g)
    enoise = torch.randn(B, 4, H // 8, W // 8, device=cuda, generator=g)
    lmask = None
    if masked:
        lmask = torch.zeros(1, 1, H // 8, W // 8, device=cuda)
        lmask[..., 4:12, 4:12] = 1.0
    seeds = [2000, 2001]
    sp = SamplingParams(sampler="Euler a", steps=12, cfg_scale=6.0, width=W, height=H, seeds=tuple(seeds), randn_source="GPU",
                        denoising_strength=0.6)
    o32 = OraclePipeline(unet, dec, cuda)
    lat32, init32 = o32.img2img(sp, enc, init, cond, uncond, encode_noise=enoise, latent_mask=lmask)
    o16 = OraclePipeline(copy.deepcopy(unet).to(dtype), copy.deepcopy(dec).to(dtype), cuda, dtype_unet=dtype, autocast=True)
    lat16, _ = o16.img2img(sp, copy.deepcopy(enc).to(dtype), init, cond, uncond, encode_noise=enoise, latent_mask=lmask)

    p = StableDiffusionProcessingImg2Img(sd_model=model, c=cond.to(dtype), uc=uncond.to(dtype), seeds=seeds, sampler_name="Euler a",
                                         steps=12, cfg_scale=6.0, width=W, height=H, randn_source="GPU", denoising_strength=0.6,
                                         init_images=init, encode_noise=enoise, latent_mask=lmask, do_not_decode=True)
    res = process_images(p)
    e_init = rel_err(p.init_latent, init32)
    e_eng, e_ref = rel_err(res.latents, lat32), rel_err(lat16, lat32)
    print(f"tiny img2img masked={masked}: init latent {e_init:.3e}; final latent engine {e_eng:.3e} ref16 {e_ref:.3e}")
    assert e_init < 5e-3
    assert e_eng < max(3 * e_ref, 5e-3), (e_eng, e_ref)
    if masked:  # outside the mask the result IS the init latent
        keep = (1 - lmask).expand_as(res.latents).bool()
        assert torch.equal(res.latents[keep], p.init_latent.float()[keep])
    p2 = StableDiffusionProcessingImg2Img(sd_model=model, c=cond.to(dtype), uc=uncond.to(dtype), seeds=seeds, sampler_name="Euler a",
                                          steps=12, cfg_scale=6.0, width=W, height=H, randn_source="GPU", denoising_strength=0.6,
                                          init_images=init, encode_noise=enoise, latent_mask=lmask)
    img = process_images(p2).images
    assert tuple(img.shape) == (B, H, W, 3) and img.dtype == torch.uint8
