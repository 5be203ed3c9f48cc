"""Checkpoint ingestion: container parsing, Lightning unwrapping, architecture guess — host logic only."""
import os

import pytest
import torch

import sdwebui_b200  # noqa: F401
from sdwebui_b200 import lib as L
from sdwebui_b200 import sd_models as M


def _rand_sd():
    g = torch.Generator().manual_seed(3)
    sd = {
        "a.weight": torch.randn(5, 7, generator=g).half(),
        "b.bias": torch.randn(9, generator=g).bfloat16(),
        "c": torch.randn(2, 3, 4, generator=g),
        "empty": torch.zeros(0, 4),
        "i": torch.arange(6, dtype=torch.int64).reshape(2, 3),
    }
    if hasattr(torch, "float8_e4m3fn"):
        sd["f8"] = torch.randn(4, 8, generator=g).to(torch.float8_e4m3fn)
    return sd


def test_safetensors_container_round_trip(tmp_path):
    sd = _rand_sd()
    p = str(tmp_path / "x.safetensors")
    M.save_safetensors(sd, p, metadata={"format": "pt"})
    back = M._read_safetensors(p)
    assert list(back) == list(sd)
    for k in sd:
        assert back[k].dtype == sd[k].dtype and back[k].shape == sd[k].shape
        assert torch.equal(back[k].view(torch.uint8) if "float8" in str(sd[k].dtype) else back[k],
                           sd[k].view(torch.uint8) if "float8" in str(sd[k].dtype) else sd[k])
    # the header is what the reference's own metadata reader expects (modules/sd_models.py:278-303)
    with open(p, "rb") as f:
        n = int.from_bytes(f.read(8), "little")
        assert n > 2 and f.read(2) in (b'{"', b"{'")


def test_safetensors_matches_the_package_reader(tmp_path):
    st = pytest.importorskip("safetensors.torch")
    sd = {k: v for k, v in _rand_sd().items() if v.numel()}
    p = str(tmp_path / "y.safetensors")
    M.save_safetensors(sd, p)
    ref = st.load_file(p)
    for k in sd:
        a, b = ref[k], sd[k]
        assert a.dtype == b.dtype and torch.equal(a.view(torch.uint8), b.view(torch.uint8))
    # and the other way round: a file written by the package parses with the built-in reader
    q = str(tmp_path / "z.safetensors")
    st.save_file({k: v.contiguous() for k, v in sd.items()}, q)
    mine = M._read_safetensors(q)
    for k in sd:
        assert torch.equal(mine[k].view(torch.uint8), sd[k].view(torch.uint8))


def test_read_state_dict_unwraps_lightning_and_renames(tmp_path):
    inner = {"cond_stage_model.transformer.embeddings.position_ids": torch.arange(4),
             "model.diffusion_model.out.2.bias": torch.ones(4)}
    p = str(tmp_path / "m.ckpt")
    torch.save({"state_dict": inner, "global_step": 7}, p)
    sd = M.read_state_dict(p)
    assert "cond_stage_model.transformer.text_model.embeddings.position_ids" in sd  # sd_models.py:243-247
    assert "model.diffusion_model.out.2.bias" in sd and "global_step" not in sd
    not_st = str(tmp_path / "bad.safetensors")
    with open(not_st, "wb") as f:
        f.write(b"\x00" * 64)
    with pytest.raises(Exception):
        M.read_state_dict(not_st)


def _skeleton(kind):
    u = M.UNET_PREFIX
    if kind == "sd15":
        return {u + "input_blocks.0.0.weight": torch.empty(320, 4, 3, 3),
                u + "input_blocks.1.1.transformer_blocks.0.attn2.to_k.weight": torch.empty(320, 768),
                M.VAE_PREFIX + "decoder.conv_in.weight": torch.empty(512, 4, 3, 3)}
    return {u + "input_blocks.0.0.weight": torch.empty(320, 4, 3, 3),
            u + "input_blocks.4.1.transformer_blocks.0.attn2.to_k.weight": torch.empty(640, 2048),
            u + "label_emb.0.0.weight": torch.empty(1280, 2816),
            "conditioner.embedders.1.model.ln_final.weight": torch.empty(1280)}


def test_guess_model_config():
    i = M.guess_model_config_from_state_dict(_skeleton("sd15"))
    assert i.kind == "sd15" and i.has_vae and i.unet.context_dim == 768
    j = M.guess_model_config_from_state_dict(_skeleton("sdxl"))
    assert j.kind == "sdxl" and not j.has_vae and j.unet.adm_in_channels == 2816
    bad = _skeleton("sd15")
    bad[M.UNET_PREFIX + "input_blocks.0.0.weight"] = torch.empty(320, 9, 3, 3)  # inpainting
    with pytest.raises(L.SdxeError):
        M.guess_model_config_from_state_dict(bad)
    sd3 = _skeleton("sd15")
    sd3[M.UNET_PREFIX + "x_embedder.proj.weight"] = torch.empty(1)
    with pytest.raises(L.SdxeError):
        M.guess_model_config_from_state_dict(sd3)
    sd2 = _skeleton("sd15")
    sd2["cond_stage_model.model.transformer.resblocks.0.attn.in_proj_weight"] = torch.empty(3072, 1024)
    with pytest.raises(L.SdxeError):
        M.guess_model_config_from_state_dict(sd2)
    with pytest.raises(L.SdxeError):
        M.guess_model_config_from_state_dict({"x": torch.empty(1)})


def test_prefix_split_and_fp8_upcast():
    sd = {M.UNET_PREFIX + "out.2.bias": torch.ones(4, dtype=torch.float64),
          M.VAE_PREFIX + "decoder.conv_in.bias": torch.ones(3).half(),
          M.VAE_PREFIX + "encoder.conv_in.bias": torch.ones(3),
          "cond_stage_model.x": torch.ones(1)}
    if hasattr(torch, "float8_e4m3fn"):
        w = torch.tensor([0.5, -1.75, 448.0, 0.015625]).to(torch.float8_e4m3fn)
        sd[M.UNET_PREFIX + "w8"] = w
    u, v = M.unet_state_dict(sd), M.vae_state_dict(sd)
    assert set(v) == {"decoder.conv_in.bias"} and u["out.2.bias"].dtype == torch.float32
    if "w8" in u:
        assert u["w8"].dtype == torch.float16 and torch.equal(u["w8"].float(), sd[M.UNET_PREFIX + "w8"].float())


def test_vae_encoder_layout_matches_the_oracle():
    """N1: the engine's expected encoder keys / shapes are exactly the oracle module's (ldm Encoder + quant_conv)."""
    from oracle.vae import AutoencoderKLEncode, VAEConfig, gaussian_sample, tiny_vae_config
    from sdwebui_b200 import checkpoint as C
    from sdwebui_b200.engine import VAESpec

    for cfg in (VAEConfig(), tiny_vae_config()):
        m = AutoencoderKLEncode(cfg)
        shapes = C.vae_encoder_param_shapes(VAESpec.from_any(cfg), cfg.embed_dim)
        sd = m.state_dict()
        assert list(shapes) == list(sd) and all(tuple(sd[k].shape) == tuple(v) for k, v in shapes.items())
    assert C.param_count(C.vae_encoder_param_shapes(VAESpec())) == 34163592 + 72
    mom = torch.cat([torch.full((1, 4, 2, 2), 0.5), torch.full((1, 4, 2, 2), 40.0)], 1)  # logvar above the clamp
    assert torch.equal(gaussian_sample(mom), torch.full((1, 4, 2, 2), 0.5))
    s = gaussian_sample(mom, torch.ones(1, 4, 2, 2))
    assert torch.allclose(s, torch.full((1, 4, 2, 2), 0.5 + float(torch.exp(torch.tensor(10.0)))))
