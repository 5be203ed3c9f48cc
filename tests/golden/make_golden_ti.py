"""Pins the textual-inversion host logic (textual_inversion.EmbeddingDatabase, TextConditionalModel.tokenize_line with
embeddings, the fix -> row replacement list) to the REFERENCE: modules/sd_hijack_clip.py executed unmodified with stub
`modules.*` packages, plus the reference's own `EmbeddingDatabase` (modules/textual_inversion/textual_inversion.py:108-256)
and `EmbeddingsWithFixes` (modules/sd_hijack.py:340-366) class bodies, taken from the files by AST and executed as they are.

    python tests/golden/make_golden_ti.py   ->   tests/golden/ti_host_ref.json
"""
import ast
import importlib.util
import json
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
sys.path.insert(0, HERE)
from make_golden_clip import BOS, COMMA, EOS, toy_tokenize  # noqa: E402

DIM = 8
EMBEDDINGS = [("myemb", 3), ("my emb long", 2), ("my", 1), ("bigone", 40), ("xl style", 2)]   # name, vectors
PROMPTS = [
    "a photo of myemb",
    "myemb",
    "(myemb:1.3) and my emb long, my cat",
    "my emb longer",                                                       # longest-first lookup: "my" then plain words
    " ".join(f"w{i}" for i in range(73)) + " myemb tail",                  # 3 vectors do not fit the 2 free slots: new chunk
    " ".join(f"w{i}" for i in range(50)) + " bigone end",                  # 40 vectors do not fit after 50 tokens
    "bigone bigone",                                                       # 40 + 40 > 75
    ", ".join(f"t{i}" for i in range(36)) + " myemb, after",              # comma backtracking with a fix in the moved part
    "xl style painting BREAK myemb",
]
BATCHES = [[0, 2], [4, 1], [5, 6, 3], [7, 8]]


def toy_embed(tokens: torch.Tensor) -> torch.Tensor:
    """[B, 77] ids -> [B, 77, DIM] 'token embedding'"""
    c = torch.arange(DIM).float()[None, None, :]
    return torch.sin(tokens.float()[:, :, None] * 0.013 + c * 0.9)


def toy_rest(e: torch.Tensor) -> torch.Tensor:
    """the 'transformer' after the embedding: position term + a mix over neighbouring positions"""
    pos = torch.arange(e.shape[1]).float()[None, :, None]
    h = e + 0.1 * torch.cos(pos * 0.07)
    return h + 0.5 * torch.roll(h, 1, dims=1) + 0.25


def embedding_vec(name, vectors, dim=DIM, seed_shift=0):
    g = torch.Generator().manual_seed(sum(ord(ch) for ch in name) + seed_shift)
    return torch.randn(vectors, dim, generator=g)


def class_source(path, name):
    src = open(path).read()
    for node in ast.parse(src).body:
        if isinstance(node, ast.ClassDef) and node.name == name:
            return ast.get_source_segment(src, node)
    raise KeyError(name)


def main():
    def pkg(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        return m

    opts = types.SimpleNamespace(emphasis="Original", comma_padding_backtrack=20, use_old_emphasis_implementation=False,
                                 textual_inversion_add_hashes_to_infotext=False, CLIP_stop_at_last_layers=1)
    spec = importlib.util.spec_from_file_location("ref_prompt_parser", os.path.join(REF, "modules/prompt_parser.py"))
    pp = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(pp)
    spec = importlib.util.spec_from_file_location("ref_sd_emphasis", os.path.join(REF, "modules/sd_emphasis.py"))
    emph = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(emph)
    devices = pkg("modules.devices", device=torch.device("cpu"), torch_npu_set_device=lambda: None, cond_cast_unet=lambda x: x)
    shared = pkg("modules.shared", opts=opts)
    # the reference's own classes, bodies unchanged
    ns = {"torch": torch, "devices": devices, "shared": shared, "os": os}
    exec(class_source(os.path.join(REF, "modules/textual_inversion/textual_inversion.py"), "Embedding"), ns)
    exec(class_source(os.path.join(REF, "modules/textual_inversion/textual_inversion.py"), "EmbeddingDatabase"), ns)
    exec(class_source(os.path.join(REF, "modules/sd_hijack.py"), "EmbeddingsWithFixes"), ns)
    db = ns["EmbeddingDatabase"]()
    hijack = types.SimpleNamespace(embedding_db=db, fixes=None, extra_generation_params={})
    sd_hijack = pkg("modules.sd_hijack", model_hijack=hijack)
    stubs = {"modules": pkg("modules", prompt_parser=pp, devices=devices, sd_hijack=sd_hijack, sd_emphasis=emph, shared=shared),
             "modules.prompt_parser": pp, "modules.devices": devices, "modules.sd_hijack": sd_hijack, "modules.sd_emphasis": emph,
             "modules.shared": shared}
    saved = {k: sys.modules.get(k) for k in stubs}
    sys.modules.update(stubs)
    try:
        spec = importlib.util.spec_from_file_location("ref_sd_hijack_clip", os.path.join(REF, "modules/sd_hijack_clip.py"))
        ref = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v

    class Toy(ref.TextConditionalModel):
        def __init__(self, key):
            super().__init__()
            self.comma_token, self.id_start, self.id_end, self.id_pad = COMMA, BOS, EOS, EOS
            self.embed = ns["EmbeddingsWithFixes"](toy_embed, hijack, textual_inversion_key=key)

        def tokenize(self, texts):
            return toy_tokenize(texts)

        def encode_with_transformers(self, tokens):
            return toy_rest(self.embed(tokens))

    model = types.SimpleNamespace(cond_stage_model=types.SimpleNamespace(tokenize=toy_tokenize))
    for name, vectors in EMBEDDINGS:
        if name == "xl style":
            vec = {"clip_l": embedding_vec(name, vectors), "clip_g": embedding_vec(name, vectors, seed_shift=7)}
        else:
            vec = embedding_vec(name, vectors)
        e = ns["Embedding"](vec, name)
        e.vectors = vectors
        e.shape = DIM
        db.register_embedding(e, model)
    out = {"lookup": {str(k): [[ids, e.name] for ids, e in v] for k, v in db.ids_lookup.items()}, "lines": [], "batches": []}
    for key in ("clip_l", "clip_g"):
        m = Toy(key)
        for p in PROMPTS:
            chunks, count = m.tokenize_line(p)
            out["lines"].append({"key": key, "prompt": p, "token_count": count,
                                 "chunks": [[c.tokens, c.multipliers, [[f.offset, f.embedding.name] for f in c.fixes]] for c in chunks]})
        for b in BATCHES:
            z = m.forward([PROMPTS[i] for i in b])
            out["batches"].append({"key": key, "prompts": b, "shape": list(z.shape),
                                   "z_sum": [round(v, 5) for v in z.sum(-1).flatten().tolist()],
                                   "z_ch3": [round(v, 6) for v in z[..., 3].flatten().tolist()]})
    # re-registration replaces, None unregisters (:129-150)
    e2 = ns["Embedding"](embedding_vec("myemb", 2, seed_shift=3), "myemb")
    e2.vectors, e2.shape = 2, DIM
    db.register_embedding(e2, model)
    db.register_embedding_by_name(None, model, "my")
    out["lookup_after"] = {str(k): [[ids, e.name, int(e.vectors)] for ids, e in v] for k, v in db.ids_lookup.items()}
    out["words_after"] = sorted(db.word_embeddings)
    path = os.path.join(HERE, "ti_host_ref.json")
    with open(path, "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print(path, os.path.getsize(path) // 1024, "KiB", len(out["lines"]), "lines", len(out["batches"]), "batches")


if __name__ == "__main__":
    main()
