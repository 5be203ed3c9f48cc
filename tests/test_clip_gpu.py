"""Row N4 on the GPU: the sdxe CLIP text transformer (sdxe_clip_forward) against oracle/clip.py — itself pinned to
transformers.CLIPTextModel, the class the reference runs (tests/test_clip_oracle_cpu.py) — on identical random weights:
every hidden state the webui can ask for (last / clip-skip through final_layer_norm, SDXL's un-normed hidden_states[11]),
both activations, and the whole prompt -> conditioning path with emphasis and multiple chunks."""
import re

import pytest
import torch

pytestmark = pytest.mark.gpu


def rel_err(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


def _models(cuda, cfg, dtype, act="quick_gelu", seed=3):
    from oracle.clip import CLIPTextModel
    from sdwebui_b200.engine import CLIPTextEngine, CLIPTextSpec

    torch.manual_seed(seed)
    m = CLIPTextModel(cfg).eval()
    with torch.no_grad():  # default nn init gives near-zero residual updates: widen so that every layer matters
        for n, p in m.named_parameters():
            if p.ndim == 2 and "embedding" not in n:
                p.mul_(3.0)
            if n.endswith("bias"):
                p.add_(0.05 * torch.randn_like(p))
    if act == "gelu":
        for layer in m.text_model.encoder.layers:
            mlp = layer.mlp
            mlp.forward = (lambda x, mlp=mlp: mlp.fc2(torch.nn.functional.gelu(mlp.fc1(x))))
    m = m.to(cuda)
    spec = CLIPTextSpec(vocab_size=cfg.vocab_size, hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size,
                        num_layers=cfg.num_layers, num_heads=cfg.num_heads, max_positions=cfg.max_positions, act=act)
    eng = CLIPTextEngine(spec, dtype=dtype, device=cuda)
    eng.load_state_dict(m.state_dict())
    eng.finalize()
    return m, eng


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("which", ["tiny", "clip_l", "gelu"])
def test_clip_hidden_states(cuda, dtype, which):
    from oracle.clip import CLIPTextConfig, tiny_clip_config

    if which == "tiny":
        cfg = tiny_clip_config()
        cfg.hidden_size, cfg.num_heads, cfg.intermediate_size = 128, 2, 512  # head dim 64, as CLIP-L / bigG
    elif which == "clip_l":
        cfg = CLIPTextConfig()                                           # 12 layers x 768, 12 heads: the SD1.x encoder
    else:
        cfg = CLIPTextConfig(vocab_size=2000, hidden_size=256, intermediate_size=1024, num_layers=4, num_heads=4, id_start=1998, id_end=1999)
    m, eng = _models(cuda, cfg, dtype, act="gelu" if which == "gelu" else "quick_gelu")
    g = torch.Generator().manual_seed(1)
    ids = torch.randint(0, cfg.vocab_size - 2, (5, 77), generator=g)
    ids[:, 0] = cfg.id_start
    ids[0, 9:] = cfg.id_end
    ids[3, 76] = cfg.id_end
    with torch.no_grad():
        hs = m.hidden_states(ids.to(cuda))
        fin = m.text_model.final_layer_norm
        L = cfg.num_layers
        cases = {"last": (L, True, fin(hs[-1])), "clip skip 2": (L - 1, True, fin(hs[-2])), "hidden[L-1] no norm": (L - 1, False, hs[L - 1]),
                 "embeddings": (0, False, hs[0])}
        m16 = m.to(dtype)
        hs16 = m16.hidden_states(ids.to(cuda))
        ref16 = {"last": m16.text_model.final_layer_norm(hs16[-1]), "clip skip 2": m16.text_model.final_layer_norm(hs16[-2]),
                 "hidden[L-1] no norm": hs16[L - 1], "embeddings": hs16[0]}
    tol = 3e-3 if dtype == torch.float16 else 2e-2
    for name, (layer, fnorm, want) in cases.items():
        out = eng.forward(ids, layer=layer, final_norm=fnorm)
        e, e_ref = rel_err(out, want), rel_err(ref16[name], want)
        print(f"clip {which} {str(dtype)[6:]} {name}: sdxe {e:.3e}  torch {str(dtype)[6:]} {e_ref:.3e}")
        assert out.shape == want.shape and e < max(2 * e_ref, tol), (name, e, e_ref)
    # fp32 output path
    out32 = eng.forward(ids, out_dtype=torch.float32)
    assert out32.dtype == torch.float32 and rel_err(out32, cases["last"][2]) < max(2 * rel_err(ref16["last"], cases["last"][2]), tol)
    eng.close()


def test_prompt_to_conditioning(cuda):
    """FrozenCLIPEmbedderWithCustomWords.forward (chunks + emphasis + clip skip) vs oracle get_learned_conditioning."""
    import types

    from oracle.clip import CLIPTextConfig, CLIPTextModel, chunk_tokens, emphasis_original  # noqa: F401
    from sdwebui_b200.prompt_parser import parse_prompt_attention
    from sdwebui_b200.sd_hijack_clip import FrozenCLIPEmbedderForSDXLWithCustomWords, FrozenCLIPEmbedderWithCustomWords, TextOptions

    cfg = CLIPTextConfig(vocab_size=3000, num_layers=3, id_start=2998, id_end=2999)
    torch.manual_seed(5)
    m = CLIPTextModel(cfg).eval().to(cuda)
    with torch.no_grad():  # EmphasisOriginal divides by the chunk's mean: a freshly initialised final LayerNorm (bias 0) makes it ~0
        m.text_model.final_layer_norm.bias.add_(0.3)

    def tok(texts, truncation=False, add_special_tokens=False):
        return {"input_ids": [[7 if w == "," else 10 + (sum((i + 1) * ord(c) for i, c in enumerate(w)) % 2900) for w in re.findall(r"[A-Za-z0-9]+|,", t)] for t in texts]}

    tokenizer_callable = type("Tok", (), {"__call__": staticmethod(tok), "get_vocab": staticmethod(lambda: {",</w>": 7}),
                                          "bos_token_id": cfg.id_start, "eos_token_id": cfg.id_end})()
    for skip in (1, 2):
        opts = TextOptions()
        opts.CLIP_stop_at_last_layers = skip
        emb = FrozenCLIPEmbedderWithCustomWords(m.state_dict(), tokenizer_callable, dtype=torch.float16, device=cuda, opts=opts)
        prompts = ["a (red:1.4) crown, jeweled", " ".join(f"w{i}" for i in range(90)) + ", (tail:0.6)"]
        z = emb.forward(prompts)
        assert z.shape == (2, 154, 768)
        # oracle: same chunks (taken from the product's own, reference-pinned tokenize_line), encoded chunk batch by chunk batch
        bc, _ = emb.process_texts(prompts)
        want = []
        with torch.no_grad():
            for i in range(2):
                chunks = [c[i] if i < len(c) else emb.empty_chunk() for c in bc]
                ids = torch.tensor([c.tokens for c in chunks], device=cuda)
                mul = torch.tensor([c.multipliers for c in chunks], device=cuda)
                want.append(emphasis_original(m.encode_with_transformers(ids, skip), mul))
        want = torch.hstack(want)
        e = rel_err(z, want)
        print(f"prompt -> cond, clip skip {skip}: rel err {e:.3e}")
        assert e < 4e-3
        emb.close()
    xl = FrozenCLIPEmbedderForSDXLWithCustomWords(m.state_dict(), tokenizer_callable, dtype=torch.float16, device=cuda, layer="hidden", layer_idx=2)
    z = xl.forward(["plain prompt"])
    with torch.no_grad():
        ids = torch.tensor([xl.tokenize_line("plain prompt")[0][0].tokens], device=cuda)
        want = m.hidden_states(ids)[2]
    assert rel_err(z, want) < 4e-3
    xl.close()


def test_open_clip2_pooled_and_sdxl_conditioning(cuda):
    """FrozenOpenCLIPEmbedder2WithCustomWords on an open_clip-named state dict (small bigG-like tower: GELU, fused in_proj,
    text_projection): penultimate hidden state without ln_final + pooled = ln_final(last)[eot] @ text_projection, against the
    torch restatement; then the SDXL {"crossattn", "vector"} assembly (modules/sd_models_xl.py:12-34) shapes."""
    import types

    from oracle.clip import CLIPTextConfig, CLIPTextModel
    from sdwebui_b200.sd_hijack_clip import (FrozenCLIPEmbedderForSDXLWithCustomWords, FrozenOpenCLIPEmbedder2WithCustomWords,
                                             sdxl_get_learned_conditioning)

    C, Lr, Hh = 256, 4, 4
    cfg = CLIPTextConfig(vocab_size=3000, hidden_size=C, intermediate_size=4 * C, num_layers=Lr, num_heads=Hh, id_start=2998, id_end=2999)
    torch.manual_seed(7)
    m = CLIPTextModel(cfg).eval()
    for layer in m.text_model.encoder.layers:
        mlp = layer.mlp
        mlp.forward = (lambda x, mlp=mlp: mlp.fc2(torch.nn.functional.gelu(mlp.fc1(x))))
    with torch.no_grad():
        m.text_model.final_layer_norm.bias.add_(0.2)
    m = m.to(cuda)
    hf = m.state_dict()
    oc = {"model.token_embedding.weight": hf["text_model.embeddings.token_embedding.weight"],
          "model.positional_embedding": hf["text_model.embeddings.position_embedding.weight"],
          "model.ln_final.weight": hf["text_model.final_layer_norm.weight"], "model.ln_final.bias": hf["text_model.final_layer_norm.bias"],
          "model.text_projection": torch.randn(C, 320, generator=torch.Generator().manual_seed(3)).to(cuda) * C ** -0.5}
    for n in range(Lr):
        s_, d_ = f"text_model.encoder.layers.{n}.", f"model.transformer.resblocks.{n}."
        oc[d_ + "attn.in_proj_weight"] = torch.cat([hf[s_ + f"self_attn.{k}_proj.weight"] for k in "qkv"])
        oc[d_ + "attn.in_proj_bias"] = torch.cat([hf[s_ + f"self_attn.{k}_proj.bias"] for k in "qkv"])
        for a, b in (("self_attn.out_proj", "attn.out_proj"), ("layer_norm1", "ln_1"), ("layer_norm2", "ln_2"), ("mlp.fc1", "mlp.c_fc"), ("mlp.fc2", "mlp.c_proj")):
            oc[d_ + b + ".weight"], oc[d_ + b + ".bias"] = hf[s_ + a + ".weight"], hf[s_ + a + ".bias"]

    def enc(text):
        return [10 + (sum((i + 1) * ord(c) for i, c in enumerate(w)) % 2900) for w in re.findall(r"[A-Za-z0-9]+", text)]

    tok = types.SimpleNamespace(encode=enc, encoder={",</w>": 7, "<start_of_text>": cfg.id_start, "<end_of_text>": cfg.id_end})
    g = FrozenOpenCLIPEmbedder2WithCustomWords(oc, tok, num_heads=Hh, dtype=torch.float16, device=cuda)
    texts = ["a castle on a hill", "portrait"]
    z, pooled = g.forward(texts)
    assert z.shape == (2, 77, C) and pooled.shape == (2, 320)
    chunks = [g.tokenize_line(t)[0][0] for t in texts]
    ids = torch.tensor([c.tokens for c in chunks], device=cuda)
    for b, c in enumerate(chunks):  # pad id 0 after end-of-text, as process_tokens does for open_clip
        ids[b, c.tokens.index(cfg.id_end) + 1:] = 0
    with torch.no_grad():
        hs = m.hidden_states(ids)
        want_z = hs[-2]
        last = m.text_model.final_layer_norm(hs[-1])
        want_p = last[torch.arange(2), ids.argmax(-1)] @ oc["model.text_projection"]
    ez, ep = rel_err(z, want_z), rel_err(pooled, want_p)
    print(f"open_clip2: penultimate rel err {ez:.3e}, pooled rel err {ep:.3e}")
    assert ez < 4e-3 and ep < 6e-3
    # SDXL assembly with a CLIP-L-like partner (hidden[layer_idx], no final norm)
    cfg_l = CLIPTextConfig(vocab_size=3000, hidden_size=128, intermediate_size=512, num_layers=3, num_heads=2, id_start=2998, id_end=2999)
    ml = CLIPTextModel(cfg_l).eval().to(cuda)
    tok_l = type("Tok", (), {"__call__": staticmethod(lambda texts, truncation=False, add_special_tokens=False: {"input_ids": [enc(t) for t in texts]}),
                             "get_vocab": staticmethod(lambda: {",</w>": 7}), "bos_token_id": cfg_l.id_start, "eos_token_id": cfg_l.id_end})()
    l = FrozenCLIPEmbedderForSDXLWithCustomWords(ml.state_dict(), tok_l, dtype=torch.float16, device=cuda, layer="hidden", layer_idx=2)
    c = sdxl_get_learned_conditioning(l, g, texts, width=1024, height=768)
    assert c["crossattn"].shape == (2, 77, 128 + C) and c["vector"].shape == (2, 320 + 6 * 256)
    uc = sdxl_get_learned_conditioning(l, g, ["", ""], is_negative_prompt=True)
    assert float(uc["crossattn"].abs().max()) == 0.0 and float(uc["vector"][:, :320].abs().max()) == 0.0
    g.close()
    l.close()


def test_prompt_driven_txt2img(cuda):
    """process_images with PROMPTS: setup_conds -> prompt_parser (editing, AND, weights) -> CLIP on the engine -> CFGDenoiser
    with per-step reconstruct_*_batch -> UNet / VAE engines. Must equal the same job fed the containers built by hand, and
    prompt editing must actually switch the conditioning mid-sampling."""
    from oracle.clip import CLIPTextModel, tiny_clip_config
    from oracle.synth import init_module_
    from oracle.unet import UNetModel, tiny_config
    from oracle.vae import AutoencoderKLDecode, tiny_vae_config
    from sdwebui_b200 import prompt_parser as P
    from sdwebui_b200.engine import UNetSpec, VAEDecoderEngine, VAESpec
    from sdwebui_b200.processing import SdModel, StableDiffusionProcessingTxt2Img, process_images
    from sdwebui_b200.sd_hijack_clip import FrozenCLIPEmbedderWithCustomWords
    from sdwebui_b200.sd_unet import SdxeUnet

    dtype = torch.float16
    ucfg, vcfg = tiny_config(), tiny_vae_config()
    unet = init_module_(UNetModel(ucfg), 1).eval().to(cuda)
    vae = init_module_(AutoencoderKLDecode(vcfg), 2).eval().to(cuda)
    su = SdxeUnet(unet.state_dict(), UNetSpec.from_any(ucfg), dtype=dtype, device=cuda)
    su.activate()
    ve = VAEDecoderEngine(VAESpec.from_any(vcfg), dtype=dtype, device=cuda)
    ve.load_state_dict(vae.state_dict())
    ve.finalize()
    model = SdModel(su, ve, is_sdxl=False, dtype_unet=dtype, device=cuda)
    cfg = tiny_clip_config()
    cfg.hidden_size, cfg.num_heads, cfg.intermediate_size = ucfg.context_dim, 2, 4 * ucfg.context_dim
    torch.manual_seed(11)
    clip = CLIPTextModel(cfg).eval().to(cuda)
    with torch.no_grad():
        clip.text_model.final_layer_norm.bias.add_(0.3)   # EmphasisOriginal divides by the chunk mean

    def tok(texts, truncation=False, add_special_tokens=False):
        return {"input_ids": [[7 if w == "," else 10 + (sum((i + 1) * ord(c) for i, c in enumerate(w)) % (cfg.vocab_size - 20)) for w in re.findall(r"[A-Za-z0-9]+|,", t)] for t in texts]}

    tokenizer = type("Tok", (), {"__call__": staticmethod(tok), "get_vocab": staticmethod(lambda: {",</w>": 7}),
                                 "bos_token_id": cfg.id_start, "eos_token_id": cfg.id_end})()
    model.cond_stage_model = FrozenCLIPEmbedderWithCustomWords(clip.state_dict(), tokenizer, dtype=dtype, device=cuda)
    T2I = StableDiffusionProcessingTxt2Img
    for cache in (T2I.cached_uc, T2I.cached_c, T2I.cached_hr_uc, T2I.cached_hr_c):
        cache[0] = cache[1] = None
    kw = dict(sd_model=model, seeds=[5, 6], steps=6, sampler_name="Euler a", cfg_scale=6.0, width=64, height=64, randn_source="NV")
    prompts, negatives = ["a [red:blue:0.5] (crown:1.3) AND jeweled :0.6", "a red crown"], ["blurry", ""]
    a = process_images(T2I(prompts=prompts, negative_prompts=negatives, **kw), to_host=True)
    # the same job with the containers built by hand
    c = P.get_multicond_learned_conditioning(model, P.SdConditioning(prompts, width=64, height=64), 6)
    uc = P.get_learned_conditioning(model, P.SdConditioning(negatives, is_negative_prompt=True, width=64, height=64), 6)
    assert [s.end_at_step for s in c.batch[0][0].schedules] == [3, 6] and len(c.batch[0]) == 2 and c.batch[0][1].weight == 0.6
    b = process_images(T2I(c=c, uc=uc, **kw), to_host=True)
    assert torch.equal(a.images, b.images) and torch.equal(a.latents, b.latents)
    # without the edit the first image changes (second prompt is the same in both jobs: it must not)
    plain = process_images(T2I(prompts=["a red (crown:1.3) AND jeweled :0.6", "a red crown"], negative_prompts=negatives, **kw), to_host=True)
    assert not torch.equal(plain.latents[0], a.latents[0]) and torch.equal(plain.latents[1], a.latents[1])
    assert torch.isfinite(a.latents).all()
    model.cond_stage_model.close()


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_textual_inversion_fixes(cuda, dtype):
    """sdxe_clip_forward_fixes vs the oracle (pinned to transformers.CLIPTextModel with its token embedding wrapped as the
    reference wraps it, tests/test_clip_oracle_cpu.py): rows replaced before the position embedding, last fix wins, a fix that
    runs past the sequence is cut; then the prompt path: a registered embedding changes exactly the prompts that name it."""
    import types

    from oracle.clip import CLIPTextConfig, tiny_clip_config
    from sdwebui_b200.sd_hijack_clip import FrozenCLIPEmbedderWithCustomWords
    from sdwebui_b200.textual_inversion import Embedding

    cfg = tiny_clip_config()
    cfg.hidden_size, cfg.num_heads, cfg.intermediate_size = 128, 2, 512
    m, eng = _models(cuda, cfg, dtype)
    g = torch.Generator().manual_seed(4)
    ids = torch.randint(0, cfg.vocab_size - 2, (3, 77), generator=g)
    ids[:, 0] = cfg.id_start
    vec_a, vec_b, vec_c = (torch.randn(k, cfg.hidden_size, generator=g) for k in (3, 2, 5))
    per_row = [[(4, vec_a)], [], [(0, vec_b), (1, vec_a), (74, vec_c)]]      # (1, vec_a) overwrites part of (0, vec_b); vec_c is cut
    flat = []
    for b, fixes in enumerate(per_row):
        for off, v in fixes:
            for j in range(min(77 - off - 1, v.shape[0])):
                flat.append((b * 77 + off + 1 + j, v[j]))
    # the oracle sees the vectors as the engine stores them (16-bit)
    per_row16 = [[(off, v.to(dtype).float()) for off, v in fixes] for fixes in per_row]
    with torch.no_grad():
        hs = m.hidden_states(ids.to(cuda), per_row16)
        plain = m.hidden_states(ids.to(cuda))
    tol = 4e-3 if dtype == torch.float16 else 3e-2
    for layer in (1, cfg.num_layers):
        got = eng.forward(ids, layer=layer, final_norm=False, fixes=flat)
        e = rel_err(got, hs[layer])
        print(f"TI fixes {dtype} layer {layer}: rel err {e:.3e} (without the fixes {rel_err(got, plain[layer]):.3e})")
        assert e < tol and rel_err(got, plain[layer]) > 10 * tol
    assert torch.equal(eng.forward(ids, layer=2, final_norm=True), eng.forward(ids, layer=2, final_norm=True, fixes=[]))
    eng.close()
    if dtype != torch.float16:
        return

    def tok(texts, truncation=False, add_special_tokens=False):
        return {"input_ids": [[7 if w == "," else 10 + (sum((i + 1) * ord(c) for i, c in enumerate(w)) % (cfg.vocab_size - 20)) for w in re.findall(r"[A-Za-z0-9]+|,", t)] for t in texts]}

    tokenizer = type("Tok", (), {"__call__": staticmethod(tok), "get_vocab": staticmethod(lambda: {",</w>": 7}),
                                 "bos_token_id": cfg.id_start, "eos_token_id": cfg.id_end})()
    with torch.no_grad():
        m.text_model.final_layer_norm.bias.add_(0.3)
    emb = FrozenCLIPEmbedderWithCustomWords(m.state_dict(), tokenizer, dtype=dtype, device=cuda)
    before = emb.forward(["a photo of wizardstyle", "a photo of a cat"])
    e = Embedding(vec_a, "wizardstyle")
    e.vectors, e.shape = 3, cfg.hidden_size
    emb.embedding_db.register_embedding(e, emb.tokenize)
    after = emb.forward(["a photo of wizardstyle", "a photo of a cat"])
    assert torch.equal(before[1], after[1]) and not torch.allclose(before[0], after[0], atol=1e-2)
    chunks, count = emb.tokenize_line("a photo of wizardstyle")
    assert count == 6 and chunks[0].fixes[0][0] == 3 and chunks[0].tokens[4:7] == [0, 0, 0]
    with torch.no_grad():
        ids1 = torch.tensor([chunks[0].tokens], device=cuda)
        want = m.encode_with_transformers(ids1, 1, [[(3, vec_a.to(dtype).float())]])
        unfixed = m.encode_with_transformers(ids1, 1)
    mult = torch.tensor([chunks[0].multipliers], device=cuda)
    from oracle.clip import emphasis_original

    # (the widened test weights put ~7e-3 of fp16 error through the final LayerNorm; the unfixed encoding is 50x further away)
    e_fix, e_plain = rel_err(after[0:1], emphasis_original(want, mult)), rel_err(after[0:1], emphasis_original(unfixed, mult))
    print(f"prompt with embedding -> cond: rel err {e_fix:.3e} (against the encoding without the embedding {e_plain:.3e})")
    assert e_fix < 1.2e-2 and e_plain > 20 * e_fix
    emb.close()
