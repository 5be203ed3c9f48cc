"""TEST INFRASTRUCTURE (see oracle/__init__.py) — groundwork for SURVEY §8(f) row N4, not used by the product.

CLIP-L text encoder as the webui runs it for SD1.x (`FrozenCLIPEmbedder.transformer` = transformers `CLIPTextModel`,
wrapped by `modules/sd_hijack_clip.py`), restated with the Hugging Face state-dict key names
(`text_model.embeddings.token_embedding.weight`, `text_model.encoder.layers.N.self_attn.q_proj.weight`, ...), so that
`cond_stage_model.transformer.*` checkpoint tensors load directly.

PINNED: tests/test_clip_oracle_cpu.py loads the same random weights into `transformers.CLIPTextModel` (the class the
reference itself calls, modules/sd_hijack_clip.py:351-360) and requires agreement to fp32 round-off, including the
`CLIP_stop_at_last_layers` ("clip skip") path and the per-chunk emphasis of `modules/sd_emphasis.py:38-49`.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F


@dataclass
class CLIPTextConfig:
    vocab_size: int = 49408
    hidden_size: int = 768
    intermediate_size: int = 3072
    num_layers: int = 12
    num_heads: int = 12
    max_positions: int = 77
    eps: float = 1e-5
    id_start: int = 49406
    id_end: int = 49407


def tiny_clip_config() -> CLIPTextConfig:
    return CLIPTextConfig(vocab_size=1000, hidden_size=64, intermediate_size=256, num_layers=3, num_heads=4, id_start=998, id_end=999)


class _Embeddings(nn.Module):
    def __init__(self, c: CLIPTextConfig):
        super().__init__()
        self.token_embedding = nn.Embedding(c.vocab_size, c.hidden_size)
        self.position_embedding = nn.Embedding(c.max_positions, c.hidden_size)

    def forward(self, ids, fixes=None):
        """`fixes`: textual-inversion replacements per batch row, [[(offset, vectors [n, C]), ...], ...] — the token embedding
        of positions offset + 1 ... is replaced before the position embedding is added (modules/sd_hijack.py:347-366)."""
        pos = torch.arange(ids.shape[1], device=ids.device)
        tok = self.token_embedding(ids)
        if fixes:
            rows = []
            for row_fixes, tensor in zip(fixes, tok):
                for offset, vec in row_fixes:
                    vec = vec.to(tensor.device, tensor.dtype)
                    n = min(tensor.shape[0] - offset - 1, vec.shape[0])
                    tensor = torch.cat([tensor[0:offset + 1], vec[0:n], tensor[offset + 1 + n:]])
                rows.append(tensor)
            tok = torch.stack(rows)
        return tok + self.position_embedding(pos)[None]


class _Attention(nn.Module):
    def __init__(self, c: CLIPTextConfig):
        super().__init__()
        self.h = c.num_heads
        self.q_proj = nn.Linear(c.hidden_size, c.hidden_size)
        self.k_proj = nn.Linear(c.hidden_size, c.hidden_size)
        self.v_proj = nn.Linear(c.hidden_size, c.hidden_size)
        self.out_proj = nn.Linear(c.hidden_size, c.hidden_size)

    def forward(self, x):
        B, T, C = x.shape
        d = C // self.h
        q = self.q_proj(x).view(B, T, self.h, d).transpose(1, 2)
        k = self.k_proj(x).view(B, T, self.h, d).transpose(1, 2)
        v = self.v_proj(x).view(B, T, self.h, d).transpose(1, 2)
        s = (q @ k.transpose(-1, -2)) / math.sqrt(d)
        causal = torch.full((T, T), float("-inf"), device=x.device, dtype=s.dtype).triu(1)  # token t sees tokens <= t
        p = torch.softmax(s + causal, dim=-1)
        return self.out_proj((p @ v).transpose(1, 2).reshape(B, T, C))


class _MLP(nn.Module):
    def __init__(self, c: CLIPTextConfig):
        super().__init__()
        self.fc1 = nn.Linear(c.hidden_size, c.intermediate_size)
        self.fc2 = nn.Linear(c.intermediate_size, c.hidden_size)

    def forward(self, x):
        h = self.fc1(x)
        return self.fc2(h * torch.sigmoid(1.702 * h))  # quick_gelu


class _Layer(nn.Module):
    def __init__(self, c: CLIPTextConfig):
        super().__init__()
        self.self_attn = _Attention(c)
        self.layer_norm1 = nn.LayerNorm(c.hidden_size, eps=c.eps)
        self.mlp = _MLP(c)
        self.layer_norm2 = nn.LayerNorm(c.hidden_size, eps=c.eps)

    def forward(self, x):
        x = x + self.self_attn(self.layer_norm1(x))
        return x + self.mlp(self.layer_norm2(x))


class _Encoder(nn.Module):
    def __init__(self, c: CLIPTextConfig):
        super().__init__()
        self.layers = nn.ModuleList([_Layer(c) for _ in range(c.num_layers)])


class _TextModel(nn.Module):
    def __init__(self, c: CLIPTextConfig):
        super().__init__()
        self.embeddings = _Embeddings(c)
        self.encoder = _Encoder(c)
        self.final_layer_norm = nn.LayerNorm(c.hidden_size, eps=c.eps)


class CLIPTextModel(nn.Module):
    def __init__(self, c: CLIPTextConfig):
        super().__init__()
        self.cfg = c
        self.text_model = _TextModel(c)

    def hidden_states(self, ids, fixes=None) -> List[torch.Tensor]:
        """[embeddings, after layer 1, ..., after layer L] — `output_hidden_states` of the HF model."""
        x = self.text_model.embeddings(ids, fixes)
        hs = [x]
        for layer in self.text_model.encoder.layers:
            x = layer(x)
            hs.append(x)
        return hs

    def encode_with_transformers(self, ids, stop_at_last_layers: int = 1, fixes=None):
        """modules/sd_hijack_clip.py:351-360: last_hidden_state, or hidden_states[-n] through the final LayerNorm."""
        hs = self.hidden_states(ids, fixes)
        z = hs[-1] if stop_at_last_layers <= 1 else hs[-stop_at_last_layers]
        return self.text_model.final_layer_norm(z)


# ---- prompt chunks and emphasis (modules/sd_hijack_clip.py:86-197, modules/sd_emphasis.py) -----------------------------
def chunk_tokens(tokens: Sequence[int], multipliers: Sequence[float], cfg: CLIPTextConfig, chunk_length: int = 75):
    """Token ids (already BPE-encoded, without specials) -> list of (77 ids, 77 multipliers): 75-token chunks wrapped in
    <start> ... <end>, padded with <end> (SD1: id_pad == id_end), multiplier 1.0 on specials and padding. An empty prompt
    is one chunk. (The comma-backtracking heuristic of tokenize_line is upstream of this restatement.)"""
    out = []
    toks, muls = list(tokens), list(multipliers)
    if not toks:
        toks, muls = [], []
    for i in range(0, max(len(toks), 1), chunk_length):
        t, m = toks[i:i + chunk_length], muls[i:i + chunk_length]
        pad = chunk_length - len(t)
        out.append(([cfg.id_start] + t + [cfg.id_end] * (pad + 1), [1.0] + m + [1.0] * (pad + 1)))
    return out


def emphasis_original(z: torch.Tensor, multipliers: torch.Tensor) -> torch.Tensor:
    """sd_emphasis.EmphasisOriginal: scale each token's vector, then restore the chunk batch's mean."""
    original_mean = z.mean()
    z = z * multipliers.reshape(multipliers.shape + (1,)).expand(z.shape)
    return z * (original_mean / z.mean())


@torch.no_grad()
def get_learned_conditioning(model: CLIPTextModel, tokens: Sequence[int], multipliers: Sequence[float],
                             stop_at_last_layers: int = 1) -> torch.Tensor:
    """One prompt -> [1, 77 * chunks, C]: FrozenCLIPEmbedderWithCustomWords.forward for a single text
    (sd_hijack_clip.py:199-251): every chunk encoded separately, emphasised, then concatenated along tokens."""
    zs = []
    for ids, muls in chunk_tokens(tokens, multipliers, model.cfg):
        dev = next(model.parameters()).device
        z = model.encode_with_transformers(torch.tensor([ids], device=dev), stop_at_last_layers)
        zs.append(emphasis_original(z, torch.tensor([muls], device=dev, dtype=z.dtype)))
    return torch.hstack(zs)
