"""sdwebui_b200/sd_hijack_clip.py host logic == the reference's modules/sd_hijack_clip.py + sd_emphasis.py on the same toy
tokenizer / toy transformer (golden: tests/golden/clip_host_ref.json, produced by tests/golden/make_golden_clip.py which
executes the reference files unmodified): chunking, BREAK, comma backtracking, emphasis modes, multi-prompt batches."""
import json
import os
import re

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "clip_host_ref.json")))
COMMA, BOS, EOS = 5, 998, 999


def toy_tokenize(texts):
    out = []
    for t in texts:
        ids = []
        for w in re.findall(r"[A-Za-z0-9]+|,|[^\sA-Za-z0-9,]", t):
            ids.append(COMMA if w == "," else 10 + (sum((i + 1) * ord(c) for i, c in enumerate(w)) % 890))
        out.append(ids)
    return out


def toy_transformer(tokens):
    t = tokens.float()
    pos = torch.arange(tokens.shape[1]).float()[None, :, None]
    c = torch.arange(8).float()[None, None, :]
    return torch.sin(t[:, :, None] * 0.013 + pos * 0.07 + c * 0.9) + 0.25


def _toy(emphasis="Original", backtrack=20):
    from sdwebui_b200.sd_hijack_clip import TextConditionalModel, TextOptions

    class Toy(TextConditionalModel):
        def tokenize(self, texts):
            return toy_tokenize(texts)

        def encode_with_transformers(self, tokens):
            return toy_transformer(tokens)

    o = TextOptions()
    o.emphasis, o.comma_padding_backtrack = emphasis, backtrack
    m = Toy(o)
    m.comma_token, m.id_start, m.id_end, m.id_pad = COMMA, BOS, EOS, EOS
    return m


def test_tokenize_line_matches_reference():
    assert len(GOLD["lines"]) >= 25
    for row in GOLD["lines"]:
        m = _toy(row["emphasis"], row.get("backtrack", 20))
        chunks, count = m.tokenize_line(row["prompt"])
        assert count == row["token_count"], row["prompt"][:40]
        assert [[c.tokens, c.multipliers] for c in chunks] == row["chunks"], (row["emphasis"], row["prompt"][:40])
        assert all(len(c.tokens) == 77 and len(c.multipliers) == 77 for c in chunks)


def test_forward_matches_reference():
    prompts = None
    import importlib.util

    spec = importlib.util.spec_from_file_location("mk", os.path.join(HERE, "golden", "make_golden_clip.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)   # only for the PROMPTS list (its main() needs /root/reference and is not run)
    prompts = mk.PROMPTS
    for row in GOLD["batches"]:
        m = _toy(row["emphasis"])
        z = m.forward([prompts[i] for i in row["prompts"]])
        assert list(z.shape) == row["shape"]
        assert torch.allclose(z.sum(-1).flatten(), torch.tensor(row["z_sum"]), atol=2e-4), row["prompts"]
        assert torch.allclose(z[..., 3].flatten(), torch.tensor(row["z_ch3"]), atol=2e-5), row["prompts"]


def test_target_token_count_and_empty_chunk():
    m = _toy()
    assert [m.get_target_prompt_token_count(n) for n in (0, 1, 75, 76, 150, 151)] == [75, 75, 75, 150, 150, 225]
    e = m.empty_chunk()
    assert e.tokens == [BOS] + [EOS] * 76 and e.multipliers == [1.0] * 77


def test_open_clip_key_mapping():
    from sdwebui_b200.sd_hijack_clip import open_clip_to_hf_state_dict

    C = 16
    sd = {"model.token_embedding.weight": torch.zeros(10, C), "model.positional_embedding": torch.zeros(77, C),
          "model.ln_final.weight": torch.ones(C), "model.ln_final.bias": torch.zeros(C)}
    for n in range(2):
        p = f"model.transformer.resblocks.{n}."
        sd[p + "attn.in_proj_weight"] = torch.arange(3 * C * C, dtype=torch.float32).reshape(3 * C, C)
        sd[p + "attn.in_proj_bias"] = torch.arange(3 * C, dtype=torch.float32)
        for k, shape in (("attn.out_proj", (C, C)), ("mlp.c_fc", (4 * C, C)), ("mlp.c_proj", (C, 4 * C))):
            sd[p + k + ".weight"], sd[p + k + ".bias"] = torch.zeros(shape), torch.zeros(shape[0])
        for k in ("ln_1", "ln_2"):
            sd[p + k + ".weight"], sd[p + k + ".bias"] = torch.ones(C), torch.zeros(C)
    hf = open_clip_to_hf_state_dict(sd)
    assert hf["text_model.encoder.layers.1.self_attn.k_proj.weight"].shape == (C, C)
    assert torch.equal(hf["text_model.encoder.layers.0.self_attn.v_proj.bias"], torch.arange(2 * C, 3 * C, dtype=torch.float32))
    assert "text_model.encoder.layers.1.mlp.fc2.weight" in hf and "text_model.final_layer_norm.weight" in hf
    from sdwebui_b200.engine import CLIPTextSpec

    spec = CLIPTextSpec.from_state_dict(hf, num_heads=2)
    assert (spec.hidden_size, spec.num_layers, spec.intermediate_size, spec.act) == (C, 2, 4 * C, "gelu")
