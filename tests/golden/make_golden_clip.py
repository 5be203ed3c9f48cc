"""Pins sdwebui_b200/sd_hijack_clip.py (prompt chunking, BREAK, comma backtracking, emphasis, multi-prompt batches) to the
REFERENCE's modules/sd_hijack_clip.py + modules/sd_emphasis.py, executed unmodified from /root/reference with stub
`modules.*` packages, a toy tokenizer and a toy "transformer" (both restated identically in tests/test_clip_host_cpu.py):

    python tests/golden/make_golden_clip.py   ->   tests/golden/clip_host_ref.json
"""
import importlib.util
import json
import os
import re
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

COMMA, BOS, EOS = 5, 998, 999


def toy_tokenize(texts):
    """words -> ids in [10, 900), ',' -> COMMA; deterministic, no vocabulary file."""
    out = []
    for t in texts:
        ids = []
        for w in re.findall(r"[A-Za-z0-9]+|,|[^\sA-Za-z0-9,]", t):
            ids.append(COMMA if w == "," else 10 + (sum((i + 1) * ord(c) for i, c in enumerate(w)) % 890))
        out.append(ids)
    return out


def toy_transformer(tokens: torch.Tensor) -> torch.Tensor:
    """[B, 77] ids -> [B, 77, 8] floats that depend on id, position and channel."""
    t = tokens.float()
    pos = torch.arange(tokens.shape[1]).float()[None, :, None]
    c = torch.arange(8).float()[None, None, :]
    return torch.sin(t[:, :, None] * 0.013 + pos * 0.07 + c * 0.9) + 0.25


PROMPTS = [
    "",
    "a photo of a cat",
    "a (very:1.3) detailed [painting] of ((mountains)), lake",
    "first part BREAK second part, with (emphasis:0.7)",
    ", ".join(f"word{i} tag{i}" for i in range(40)),                      # > 75 tokens with commas: backtracking
    " ".join(f"w{i}" for i in range(160)),                                 # three chunks, no commas
    ", ".join(f"t{i}" for i in range(37)) + " " + " ".join(f"x{i}" for i in range(30)),  # comma too far back to cut at
    "(" + ", ".join(f"e{i}" for i in range(50)) + ":1.2) tail, end",
    "a BREAK b BREAK c",
]
BATCHES = [[1, 2], [0, 4], [5, 3, 1], [7, 8]]


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
    hijack = types.SimpleNamespace(embedding_db=types.SimpleNamespace(find_embedding_at_position=lambda tokens, pos: (None, None)),
                                   fixes=None, extra_generation_params={})
    devices = pkg("modules.devices", device=torch.device("cpu"), torch_npu_set_device=lambda: None)
    shared = pkg("modules.shared", opts=opts)
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
        def __init__(self):
            super().__init__()
            self.comma_token, self.id_start, self.id_end, self.id_pad = COMMA, BOS, EOS, EOS

        def tokenize(self, texts):
            return toy_tokenize(texts)

        def encode_with_transformers(self, tokens):
            return toy_transformer(tokens)

    out = {"lines": [], "batches": []}
    for emphasis in ("Original", "No norm", "None"):
        opts.emphasis = emphasis
        m = Toy()
        for p in PROMPTS:
            chunks, count = m.tokenize_line(p)
            out["lines"].append({"emphasis": emphasis, "prompt": p, "token_count": count,
                                 "chunks": [[c.tokens, c.multipliers] for c in chunks]})
        for b in BATCHES:
            z = m.forward([PROMPTS[i] for i in b])
            out["batches"].append({"emphasis": emphasis, "prompts": b, "shape": list(z.shape),
                                   "z_sum": [round(v, 5) for v in z.sum(-1).flatten().tolist()],       # per (prompt, token)
                                   "z_ch3": [round(v, 6) for v in z[..., 3].flatten().tolist()]})
    for bt in (0, 5):
        opts.emphasis, opts.comma_padding_backtrack = "Original", bt
        m = Toy()
        chunks, count = m.tokenize_line(PROMPTS[4])
        out["lines"].append({"emphasis": "Original", "backtrack": bt, "prompt": PROMPTS[4], "token_count": count,
                             "chunks": [[c.tokens, c.multipliers] for c in chunks]})
    path = os.path.join(HERE, "clip_host_ref.json")
    with open(path, "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print(path, os.path.getsize(path) // 1024, "KiB", len(out["lines"]), "lines", len(out["batches"]), "batches")


if __name__ == "__main__":
    main()
