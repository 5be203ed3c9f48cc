"""Checkpoint file -> engines on the GPU: a full-size SD1.5-layout checkpoint written with the webui's
key prefixes loads through sd_models.load_model and computes exactly what an engine fed the same tensors directly does."""
import pytest
import torch

import sdwebui_b200  # noqa: F401
from sdwebui_b200 import checkpoint as C
from sdwebui_b200 import sd_models as M
from sdwebui_b200.engine import UNetEngine, UNetSpec, VAESpec
from sdwebui_b200.processing import StableDiffusionProcessingTxt2Img, process_images

pytestmark = pytest.mark.gpu


def test_load_model_from_safetensors(cuda, tmp_path):
    spec = UNetSpec.sd15()
    usd = C.synthetic_state_dict(C.unet_param_shapes(spec), 11, device=cuda, dtype=torch.float16)
    vsd = C.synthetic_state_dict(C.vae_decoder_param_shapes(VAESpec()), 12, device=cuda, dtype=torch.float16)
    full = {M.UNET_PREFIX + k: v for k, v in usd.items()}
    full.update({M.VAE_PREFIX + k: v for k, v in vsd.items()})
    full["cond_stage_model.transformer.text_model.embeddings.position_ids"] = torch.arange(77).unsqueeze(0)
    path = str(tmp_path / "synthetic-sd15.safetensors")
    M.save_safetensors(full, path, metadata={"format": "pt"})
    del full

    model, info = M.load_model(path, dtype=torch.float16, device=str(cuda))
    assert info.kind == "sd15" and info.has_vae and model.vae is not None
    assert model.unet.engine.param_count() == 859_520_964

    direct = UNetEngine(spec, dtype=torch.float16, device=cuda)
    direct.load_state_dict(usd)
    direct.finalize()
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn(2, 4, 32, 32, device=cuda, generator=g).half()
    t = torch.tensor([801.0, 13.0], device=cuda).half()
    ctx = torch.randn(2, 77, 768, device=cuda, generator=g).half()
    a = model.unet.engine.forward(x, t, ctx, None)
    b = direct.forward(x, t, ctx, None)
    assert torch.isfinite(a).all() and torch.equal(a, b)

    cond = torch.randn(1, 77, 768, device=cuda, generator=g).half()
    p = StableDiffusionProcessingTxt2Img(sd_model=model, c=cond, uc=torch.zeros_like(cond), seeds=[7], sampler_name="Euler a",
                                         steps=2, cfg_scale=7.0, width=256, height=256, randn_source="GPU")
    img = process_images(p).images
    assert tuple(img.shape) == (1, 256, 256, 3) and img.dtype == torch.uint8
