"""Textual-inversion embeddings for the text conditioner (the "custom words" of FrozenCLIPEmbedderWithCustomWords). Host mirror
of the inference half of modules/textual_inversion/textual_inversion.py:

  Embedding                       :28-75     learned vectors [n_vectors, dim] (SDXL: {"clip_l": ..., "clip_g": ...}) + a name
  create_embedding_from_data      :287-323   the three file layouts (A1111 `string_to_param`, SDXL clip_l / clip_g, diffusers concept)
  EmbeddingDatabase               :108-256   name -> embedding, token-id prefix lookup, `find_embedding_at_position`

`TextConditionalModel.tokenize_line` asks the database at every token position; a hit reserves `vectors` placeholder tokens
and records a fix (offset, embedding); `encode_with_transformers` hands the fixes to the engine, which overwrites those rows of
the token embedding (`sdxe_clip_forward_fixes`). Training, PNG-embedded embeddings and file hashing stay with the webui.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Tuple

import torch

from . import lib as L


class Embedding:
    def __init__(self, vec, name, step=None):
        self.vec = vec            # tensor [vectors, dim], or {"clip_l": [v, 768], "clip_g": [v, 1280]}
        self.name = name
        self.step = step
        self.shape = None         # dim (SDXL: the sum of both)
        self.vectors = 0
        self.sd_checkpoint = None
        self.sd_checkpoint_name = None
        self.filename = None


def create_embedding_from_data(data, name, filename="unknown embedding file", filepath=None) -> Embedding:
    """:287-323. Vectors are kept in fp32 on the host; the engine wrapper casts the ones a prompt uses."""
    if "string_to_param" in data:                                        # textual inversion embeddings
        param_dict = data["string_to_param"]
        param_dict = getattr(param_dict, "_parameters", param_dict)
        if len(param_dict) != 1:
            raise L.SdxeError("embedding file has multiple terms in it")
        emb = next(iter(param_dict.items()))[1]
        vec = emb.detach().to(dtype=torch.float32)
        shape, vectors = vec.shape[-1], vec.shape[0]
    elif type(data) == dict and "clip_g" in data and "clip_l" in data:   # SDXL embedding
        vec = {k: v.detach().to(dtype=torch.float32) for k, v in data.items()}
        shape = data["clip_g"].shape[-1] + data["clip_l"].shape[-1]
        vectors = data["clip_g"].shape[0]
    elif type(data) == dict and type(next(iter(data.values()))) == torch.Tensor:   # diffuser concepts
        if len(data.keys()) != 1:
            raise L.SdxeError("embedding file has multiple terms in it")
        emb = next(iter(data.values()))
        if len(emb.shape) == 1:
            emb = emb.unsqueeze(0)
        vec = emb.detach().to(dtype=torch.float32)
        shape, vectors = vec.shape[-1], vec.shape[0]
    else:
        raise L.SdxeError(f"Couldn't identify {filename} as neither textual inversion embedding nor diffuser concept.")
    embedding = Embedding(vec, name)
    embedding.step = data.get("step", None)
    embedding.sd_checkpoint = data.get("sd_checkpoint", None)
    embedding.sd_checkpoint_name = data.get("sd_checkpoint_name", None)
    embedding.vectors = vectors
    embedding.shape = shape
    embedding.filename = filepath
    return embedding


class EmbeddingDatabase:
    """:108-256 (lookup side). `tokenize` is the conditioner's `tokenize([name]) -> [[ids]]`."""

    def __init__(self):
        self.ids_lookup: Dict[int, List[Tuple[List[int], Embedding]]] = {}
        self.word_embeddings: Dict[str, Embedding] = {}
        self.skipped_embeddings: Dict[str, Embedding] = {}
        self.expected_shape = -1

    def clear(self):
        self.ids_lookup.clear()
        self.word_embeddings.clear()
        self.skipped_embeddings.clear()

    def register_embedding(self, embedding: Embedding, tokenize):
        return self.register_embedding_by_name(embedding, tokenize, embedding.name)

    def register_embedding_by_name(self, embedding: Optional[Embedding], tokenize, name: str):
        """:129-150 — `embedding=None` unregisters `name`. Candidates sharing a first token are tried longest name first."""
        ids = tokenize([name])[0]
        first_id = ids[0]
        if first_id not in self.ids_lookup:
            self.ids_lookup[first_id] = []
        if name in self.word_embeddings:
            lookup = [x for x in self.ids_lookup[first_id] if x[1].name != name]   # replace the old entry
        else:
            lookup = self.ids_lookup[first_id]
        if embedding is not None:
            lookup += [(ids, embedding)]
        self.ids_lookup[first_id] = sorted(lookup, key=lambda x: len(x[0]), reverse=True)
        if embedding is None:
            if name in self.word_embeddings:
                del self.word_embeddings[name]
            if len(self.ids_lookup[first_id]) == 0:
                del self.ids_lookup[first_id]
            return None
        self.word_embeddings[name] = embedding
        return embedding

    def find_embedding_at_position(self, tokens, offset):
        """:245-256 -> (embedding, number of prompt tokens its name spans) or (None, None)."""
        possible_matches = self.ids_lookup.get(tokens[offset], None)
        if possible_matches is None:
            return None, None
        for ids, embedding in possible_matches:
            if tokens[offset:offset + len(ids)] == ids:
                return embedding, len(ids)
        return None, None

    def load_from_file(self, path: str, tokenize, expected_shape: int = -1) -> Optional[Embedding]:
        """:157-203 for .pt / .bin / .safetensors: the embedding is named after the file; one whose width does not fit the
        loaded text encoder (`expected_shape`, -1 = anything) is parked in `skipped_embeddings`, as the reference does."""
        name, ext = os.path.splitext(os.path.basename(path))
        ext = ext.upper()
        if ext in (".BIN", ".PT"):
            data = torch.load(path, map_location="cpu", weights_only=True)
        elif ext == ".SAFETENSORS":
            from .sd_models import _read_safetensors

            data = _read_safetensors(path)
        else:
            return None
        embedding = create_embedding_from_data(data, name, filename=os.path.basename(path), filepath=path)
        if expected_shape == -1 or expected_shape == embedding.shape:
            return self.register_embedding(embedding, tokenize)
        self.skipped_embeddings[name] = embedding
        return None
