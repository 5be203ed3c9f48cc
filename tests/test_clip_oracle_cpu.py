"""Pins oracle/clip.py (groundwork for SURVEY §8(f) N4) against transformers.CLIPTextModel — the class the reference
itself runs (modules/sd_hijack_clip.py:351-360) — with identical random weights, on CPU."""
import pytest
import torch

from oracle.clip import (CLIPTextConfig, CLIPTextModel, chunk_tokens, emphasis_original, get_learned_conditioning,
                         tiny_clip_config)

transformers = pytest.importorskip("transformers")


def _hf(cfg):
    hc = transformers.CLIPTextConfig(vocab_size=cfg.vocab_size, hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size,
                                     num_hidden_layers=cfg.num_layers, num_attention_heads=cfg.num_heads,
                                     max_position_embeddings=cfg.max_positions, hidden_act="quick_gelu", layer_norm_eps=cfg.eps,
                                     bos_token_id=cfg.id_start, eos_token_id=cfg.id_end, pad_token_id=cfg.id_end,
                                     attn_implementation="eager")
    return transformers.CLIPTextModel(hc).eval()


@pytest.mark.parametrize("cfg", [tiny_clip_config(), CLIPTextConfig(num_layers=2)])
def test_matches_transformers_clip_text_model(cfg):
    torch.manual_seed(0)
    hf = _hf(cfg)
    mine = CLIPTextModel(cfg).eval()
    sd = {k: v for k, v in hf.state_dict().items() if "position_ids" not in k}
    missing, unexpected = mine.load_state_dict(sd, strict=False)
    assert not missing and not unexpected, (missing, unexpected)  # same key names as cond_stage_model.transformer.*
    g = torch.Generator().manual_seed(1)
    ids = torch.randint(0, cfg.vocab_size - 2, (2, 77), generator=g)
    ids[:, 0] = cfg.id_start
    ids[0, 20:] = cfg.id_end
    ids[1, 76] = cfg.id_end
    with torch.no_grad():
        out = hf(input_ids=ids, output_hidden_states=True)
        hs = mine.hidden_states(ids)
        assert len(hs) == len(out.hidden_states)
        for a, b in zip(hs, out.hidden_states):
            assert torch.allclose(a, b, atol=2e-5, rtol=1e-5), (a - b).abs().max()
        assert torch.allclose(mine.encode_with_transformers(ids, 1), out.last_hidden_state, atol=2e-5, rtol=1e-5)
        # clip skip 2: hidden_states[-2] through the final LayerNorm (sd_hijack_clip.py:354-356)
        want = hf.text_model.final_layer_norm(out.hidden_states[-2])
        assert torch.allclose(mine.encode_with_transformers(ids, 2), want, atol=2e-5, rtol=1e-5)


def test_textual_inversion_fixes_match_wrapped_token_embedding():
    """The reference applies embeddings by wrapping the HF model's token_embedding (modules/sd_hijack.py:340-366, installed at
    :196-231). Same wrapping here around transformers.CLIPTextModel vs the oracle's `fixes` argument."""
    cfg = tiny_clip_config()
    torch.manual_seed(3)
    hf = _hf(cfg)
    mine = CLIPTextModel(cfg).eval()
    mine.load_state_dict({k: v for k, v in hf.state_dict().items() if "position_ids" not in k}, strict=False)
    g = torch.Generator().manual_seed(2)
    ids = torch.randint(0, cfg.vocab_size - 2, (3, 77), generator=g)
    ids[:, 0] = cfg.id_start
    fixes = [[(4, torch.randn(3, cfg.hidden_size, generator=g))], [],
             [(0, torch.randn(2, cfg.hidden_size, generator=g)), (74, torch.randn(5, cfg.hidden_size, generator=g))]]  # the last one is cut at position 77

    class Wrapped(torch.nn.Module):
        def __init__(self, inner):
            super().__init__()
            self.inner = inner

        def forward(self, input_ids):
            e = self.inner(input_ids)
            rows = []
            for row_fixes, tensor in zip(fixes, e):
                for offset, vec in row_fixes:
                    n = min(tensor.shape[0] - offset - 1, vec.shape[0])
                    tensor = torch.cat([tensor[0:offset + 1], vec[0:n], tensor[offset + 1 + n:]])
                rows.append(tensor)
            return torch.stack(rows)

    hf.text_model.embeddings.token_embedding = Wrapped(hf.text_model.embeddings.token_embedding)
    with torch.no_grad():
        out = hf(input_ids=ids, output_hidden_states=True)
        hs = mine.hidden_states(ids, fixes)
        for a, b in zip(hs, out.hidden_states):
            assert torch.allclose(a, b, atol=2e-5, rtol=1e-5), (a - b).abs().max()
        assert not torch.allclose(hs[-1], mine.hidden_states(ids)[-1], atol=1e-3)   # the fixes matter


def test_chunks_and_emphasis():
    cfg = tiny_clip_config()
    ch = chunk_tokens([], [], cfg)
    assert len(ch) == 1 and ch[0][0] == [cfg.id_start] + [cfg.id_end] * 76 and ch[0][1] == [1.0] * 77
    toks = list(range(100))
    muls = [1.1] * 100
    ch = chunk_tokens(toks, muls, cfg)
    assert len(ch) == 2 and all(len(t) == 77 and len(m) == 77 for t, m in ch)
    assert ch[0][0][0] == cfg.id_start and ch[0][0][1:76] == toks[:75] and ch[0][0][76] == cfg.id_end
    assert ch[1][0][1:26] == toks[75:] and ch[1][0][26:] == [cfg.id_end] * 51 and ch[1][1][26:] == [1.0] * 51
    z = torch.randn(1, 77, 8, generator=torch.Generator().manual_seed(2)) + 0.3
    m = torch.ones(1, 77)
    assert torch.allclose(emphasis_original(z, m), z)
    m[0, 5] = 1.3
    e = emphasis_original(z, m)
    assert torch.allclose(e.mean(), z.mean(), atol=1e-6)                          # the mean is restored ...
    assert torch.allclose(e[0, 5] / e[0, 6].norm(), 1.3 * z[0, 5] / z[0, 6].norm(), rtol=1e-5)  # ... the ratio kept
    model = CLIPTextModel(cfg).eval()
    c = get_learned_conditioning(model, toks, muls)
    assert tuple(c.shape) == (1, 154, cfg.hidden_size)
