"""Textual-inversion host logic == the reference (golden: tests/golden/ti_host_ref.json, produced by
tests/golden/make_golden_ti.py from the reference's own sd_hijack_clip.py, EmbeddingDatabase and EmbeddingsWithFixes):
name lookup (longest first, shared first token), placeholder tokens and fixes per chunk (chunk-boundary and comma-backtracking
cases), SDXL clip_l / clip_g selection, and the rows the engine is asked to overwrite."""
import json
import os

import pytest
import torch

from test_clip_host_cpu import BOS, COMMA, EOS, toy_tokenize

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "ti_host_ref.json")))
DIM = 8
EMBEDDINGS = [("myemb", 3), ("my emb long", 2), ("my", 1), ("bigone", 40), ("xl style", 2)]
PROMPTS = [
    "a photo of myemb", "myemb", "(myemb:1.3) and my emb long, my cat", "my emb longer",
    " ".join(f"w{i}" for i in range(73)) + " myemb tail", " ".join(f"w{i}" for i in range(50)) + " bigone end", "bigone bigone",
    ", ".join(f"t{i}" for i in range(36)) + " myemb, after", "xl style painting BREAK myemb",
]


def toy_embed(tokens):
    c = torch.arange(DIM).float()[None, None, :]
    return torch.sin(tokens.float()[:, :, None] * 0.013 + c * 0.9)


def toy_rest(e):
    pos = torch.arange(e.shape[1]).float()[None, :, None]
    h = e + 0.1 * torch.cos(pos * 0.07)
    return h + 0.5 * torch.roll(h, 1, dims=1) + 0.25


def embedding_vec(name, vectors, dim=DIM, seed_shift=0):
    g = torch.Generator().manual_seed(sum(ord(ch) for ch in name) + seed_shift)
    return torch.randn(vectors, dim, generator=g)


def _toy(key):
    from sdwebui_b200.sd_hijack_clip import TextConditionalModel, TextOptions
    from sdwebui_b200.textual_inversion import Embedding

    class Toy(TextConditionalModel):
        def tokenize(self, texts):
            return toy_tokenize(texts)

        def encode_with_transformers(self, tokens, fixes=None):
            e = toy_embed(tokens)                                   # what the engine does on the device: overwrite rows, in order
            flat = e.reshape(-1, DIM)
            for row, vec in self.fix_rows(fixes, tokens.shape[1]):
                flat[row] = vec
            return toy_rest(flat.reshape(e.shape))

    m = Toy(TextOptions())
    m.comma_token, m.id_start, m.id_end, m.id_pad = COMMA, BOS, EOS, EOS
    m.textual_inversion_key = key
    for name, vectors in EMBEDDINGS:
        if name == "xl style":
            vec = {"clip_l": embedding_vec(name, vectors), "clip_g": embedding_vec(name, vectors, seed_shift=7)}
        else:
            vec = embedding_vec(name, vectors)
        e = Embedding(vec, name)
        e.vectors, e.shape = vectors, DIM
        m.embedding_db.register_embedding(e, m.tokenize)
    return m


def test_lookup_table_matches_reference():
    m = _toy("clip_l")
    got = {str(k): [[ids, e.name] for ids, e in v] for k, v in m.embedding_db.ids_lookup.items()}
    assert got == GOLD["lookup"]
    from sdwebui_b200.textual_inversion import Embedding

    e2 = Embedding(embedding_vec("myemb", 2, seed_shift=3), "myemb")
    e2.vectors, e2.shape = 2, DIM
    m.embedding_db.register_embedding(e2, m.tokenize)                # same name: replaces
    m.embedding_db.register_embedding_by_name(None, m.tokenize, "my")   # None: unregisters
    after = {str(k): [[ids, e.name, int(e.vectors)] for ids, e in v] for k, v in m.embedding_db.ids_lookup.items()}
    assert after == GOLD["lookup_after"] and sorted(m.embedding_db.word_embeddings) == GOLD["words_after"]


@pytest.mark.parametrize("key", ["clip_l", "clip_g"])
def test_tokenize_line_with_embeddings_matches_reference(key):
    m = _toy(key)
    lines = [x for x in GOLD["lines"] if x["key"] == key]
    assert len(lines) == len(PROMPTS)
    for ref in lines:
        chunks, count = m.tokenize_line(ref["prompt"])
        assert count == ref["token_count"], ref["prompt"]
        assert [[c.tokens, c.multipliers, [[off, emb.name] for off, emb in c.fixes]] for c in chunks] == ref["chunks"], ref["prompt"]


@pytest.mark.parametrize("key", ["clip_l", "clip_g"])
def test_forward_with_fixes_matches_reference(key):
    m = _toy(key)
    for ref in [x for x in GOLD["batches"] if x["key"] == key]:
        z = m.forward([PROMPTS[i] for i in ref["prompts"]])
        assert list(z.shape) == ref["shape"]
        assert torch.allclose(z.sum(-1).flatten(), torch.tensor(ref["z_sum"]), atol=2e-4)
        assert torch.allclose(z[..., 3].flatten(), torch.tensor(ref["z_ch3"]), atol=2e-5)


def test_file_layouts_and_width_check(tmp_path):
    from sdwebui_b200.lib import SdxeError
    from sdwebui_b200.sd_models import save_safetensors
    from sdwebui_b200.textual_inversion import EmbeddingDatabase, create_embedding_from_data

    a = create_embedding_from_data({"string_to_param": {"*": torch.ones(3, 768)}, "step": 500, "sd_checkpoint_name": "x"}, "a1111")
    assert (a.vectors, a.shape, a.step, a.sd_checkpoint_name) == (3, 768, 500, "x") and a.vec.dtype == torch.float32
    xl = create_embedding_from_data({"clip_l": torch.ones(2, 768), "clip_g": torch.ones(2, 1280)}, "xl")
    assert (xl.vectors, xl.shape) == (2, 2048) and set(xl.vec) == {"clip_l", "clip_g"}
    d = create_embedding_from_data({"<concept>": torch.ones(768)}, "diffusers")
    assert (d.vectors, d.shape) == (1, 768)
    with pytest.raises(SdxeError):
        create_embedding_from_data({"foo": 1}, "junk")
    db = EmbeddingDatabase()
    f = str(tmp_path / "wizard.safetensors")
    save_safetensors({"emb_params": torch.randn(4, 768)}, f)
    assert db.load_from_file(f, toy_tokenize, expected_shape=768).name == "wizard"
    assert db.find_embedding_at_position(toy_tokenize(["a wizard"])[0], 1)[1] == 1
    g = str(tmp_path / "wide.pt")
    torch.save({"string_to_param": {"*": torch.randn(2, 1024)}}, g)
    assert db.load_from_file(g, toy_tokenize, expected_shape=768) is None and "wide" in db.skipped_embeddings
