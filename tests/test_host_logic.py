"""CPU tests of the product's host side: the C-ABI library loads and exports every declared symbol, config struct
layout, checkpoint layout, sampler host arithmetic, RNG, sharding, and 'fails loudly without CUDA'."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "sdxe.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sdxe_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from sdwebui_b200 import lib as L

    lib = L.load()
    names = _header_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/sdxe.h but not exported by libsdxe.so"
        assert n in L.SYMBOLS, f"{n} has no ctypes signature in lib.py"
    assert lib.sdxe_version() >= 1


def test_config_struct_matches_header():
    from sdwebui_b200 import lib as L

    src = open(os.path.join(ROOT, "include", "sdxe.h")).read()
    body = src[src.index("typedef struct sdxe_config {"):src.index("} sdxe_config;")]
    fields = re.findall(r"int32_t\s+([a-z_]+)(\[[A-Z_0-9]+\])?;", body)
    py = [f[0] for f in L.SdxeConfig._fields_]
    assert [f[0] for f in fields] == py
    n_ints = sum((1 if not dim else (8 if dim == "[SDXE_MAX_LEVELS]" else int(dim[1:-1]))) for _, dim in fields)
    assert ctypes.sizeof(L.SdxeConfig) == 4 * n_ints


def test_compute_calls_fail_loudly_without_cuda():
    from sdwebui_b200 import lib as L
    from sdwebui_b200 import ops
    from sdwebui_b200.engine import UNetEngine, UNetSpec

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(L.SdxeError):
        UNetEngine(UNetSpec.sd15())
    with pytest.raises(L.SdxeError):
        ops.attention(torch.zeros(1, 1, 8, 8, dtype=torch.float16), torch.zeros(1, 1, 8, 8, dtype=torch.float16),
                      torch.zeros(1, 1, 8, 8, dtype=torch.float16))
    cfg = L.SdxeConfig()
    cfg.kind, cfg.dtype, cfg.num_levels, cfg.model_channels = 0, 0, 1, 64
    h = ctypes.c_void_p()
    assert L.load().sdxe_create(ctypes.byref(cfg), ctypes.byref(h)) != 0
    assert b"CUDA" in L.load().sdxe_last_error()


def test_product_does_not_import_oracle():
    """The product path must never route through the oracle (or any CPU fallback)."""
    pkg = os.path.join(ROOT, "stable-diffusion-webui_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f


def test_checkpoint_layout_matches_architecture():
    from oracle.unet import UNetModel, sd15_config, sdxl_config, tiny_config
    from oracle.vae import AutoencoderKLDecode, VAEConfig
    from sdwebui_b200 import checkpoint as C
    from sdwebui_b200.engine import UNetSpec, VAESpec

    for spec, cfg, total in ((UNetSpec.sd15(), sd15_config(), 859_520_964), (UNetSpec.sdxl(), sdxl_config(), 2_567_463_684),
                             (UNetSpec.from_any(tiny_config(True, 96)), tiny_config(True, 96), None)):
        shapes = C.unet_param_shapes(spec)
        with torch.device("meta"):
            ref = {k: tuple(v.shape) for k, v in UNetModel(cfg).state_dict().items()}
        assert dict(shapes) == ref
        if total:
            assert C.param_count(shapes) == total
    with torch.device("meta"):
        ref = {k: tuple(v.shape) for k, v in AutoencoderKLDecode(VAEConfig()).state_dict().items()}
    vs = C.vae_decoder_param_shapes(VAESpec())
    assert dict(vs) == ref and C.param_count(vs) == 49_490_199
    sd = C.synthetic_state_dict(C.unet_param_shapes(UNetSpec.from_any(tiny_config())), seed=3)
    sd2 = C.synthetic_state_dict(C.unet_param_shapes(UNetSpec.from_any(tiny_config())), seed=3)
    assert all(torch.equal(sd[k], sd2[k]) for k in sd)


def test_sampler_host_arithmetic_matches_oracle():
    from oracle import kdiffusion as K
    from sdwebui_b200 import samplers as S

    ac = S.make_alphas_cumprod()
    assert torch.equal(ac, K.make_alphas_cumprod())
    prod, ora = S.DiscreteSchedule(ac, "cpu"), K.DiscreteSchedule(ac)
    assert torch.equal(prod.get_sigmas(20), ora.get_sigmas(20))
    sig = torch.tensor([14.6, 3.3, 0.7, 0.05])
    assert torch.equal(prod.sigma_to_t(sig), ora.sigma_to_t(sig))
    assert torch.equal(S.get_sigmas_karras(30, 0.0292, 14.6146), K.get_sigmas_karras(30, 0.0292, 14.6146))
    for a, b in ((14.6, 11.0), (1.0, 0.5), (0.1, 0.0)):
        assert np.allclose(S.get_ancestral_step(a, b), K.get_ancestral_step(a, b))

    class P:
        denoising_strength, steps = 0.75, 20

    assert S.setup_img2img_steps(P(), 20) == K.setup_img2img_steps(20, 0.75) == (26, 19)


def test_sampler_registry_and_errors():
    from sdwebui_b200 import lib as L
    from sdwebui_b200 import samplers as S

    assert S._sampler_map["euler a"][0] == "Euler a" and S._sampler_map["k_dpmpp_2m"][0] == "DPM++ 2M"
    assert S._sampler_map["dpm++ 2m"][2]["scheduler"] == "karras"   # sd_samplers_kdiffusion.py:12

    class FakeModel:
        alphas_cumprod = S.make_alphas_cumprod()
        device = torch.device("cpu")

    with pytest.raises(L.SdxeError):
        S.KDiffusionSampler("DDIM", FakeModel())
    smp = S.KDiffusionSampler("Euler a", FakeModel())

    class P:
        scheduler = "Automatic"

    assert smp.get_sigmas(P(), 20).shape == (21,)
    assert S.KDiffusionSampler("DPM++ 2M", FakeModel()).get_sigmas(P(), 30)[0] > 14.6


def test_image_rng_batch_equals_singles():
    from oracle.rng import ImageRNG as OracleRNG
    from sdwebui_b200.rng import ImageRNG

    a = ImageRNG((4, 8, 8), [5, 6, 7], source="NV", device="cpu")
    b = [ImageRNG((4, 8, 8), [s], source="NV", device="cpu") for s in (5, 6, 7)]
    o = OracleRNG((4, 8, 8), [5, 6, 7], source="NV", device="cpu")
    for _ in range(3):
        xa = a.next()
        assert torch.equal(xa, torch.cat([g.next() for g in b]))
        assert torch.equal(xa, o.next())
    c = ImageRNG((4, 8, 8), [5], source="CPU", device="cpu")
    g = torch.Generator().manual_seed(5)
    assert torch.equal(c.next()[0], torch.randn((4, 8, 8), generator=g))


def test_shard_indices_partition():
    from sdwebui_b200.parallel import shard, shard_indices

    for n, world in ((64, 8), (32, 8), (7, 4), (3, 8)):
        parts = [shard_indices(n, r, world) for r in range(world)]
        assert sorted(sum(parts, [])) == list(range(n))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    assert shard(list(range(10, 18)), 1, 4) == [12, 13]
    assert torch.equal(shard(torch.arange(8), 3, 4), torch.tensor([6, 7]))


def test_prompt_attention_matches_reference_golden():
    """sdwebui_b200.prompt_parser.parse_prompt_attention == modules/prompt_parser.py:370-458 on every golden prompt
    (hand-written cases incl. the reference's own doctest examples, plus 300 fuzzed strings; generator:
    tests/golden/make_golden.py, which imports the reference module from /root/reference)."""
    import json
    import os

    from sdwebui_b200.prompt_parser import parse_prompt_attention

    with open(os.path.join(os.path.dirname(__file__), "golden", "prompt_attention.json")) as f:
        cases = json.load(f)
    assert len(cases) > 300
    for c in cases:
        if "error" in c:
            try:
                parse_prompt_attention(c["prompt"])
            except Exception as e:  # noqa: BLE001
                assert type(e).__name__ == c["error"], (c["prompt"], type(e).__name__, c["error"])
            else:
                raise AssertionError(f"expected {c['error']} for {c['prompt']!r}")
        else:
            assert parse_prompt_attention(c["prompt"]) == c["result"], c["prompt"]
