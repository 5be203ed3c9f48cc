"""Prompt-editing schedules and prompt -> conditioning containers against the reference itself: tests/golden/
prompt_sched_ref.json holds the outputs of /root/reference/modules/prompt_parser.py (lark grammar) for its doctest prompts,
edge cases and a 1200-prompt fuzz corpus (tests/golden/make_golden_prompt_sched.py); the product's hand-written parser must
reproduce every schedule exactly."""
import json
import os

import torch

from sdwebui_b200 import prompt_parser as P

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "prompt_sched_ref.json")))


def test_schedules_match_reference_corpus():
    assert len(GOLD["cases"]) > 1500
    for prompt, steps, hires, old, want in GOLD["cases"]:
        got = P.get_learned_conditioning_prompt_schedules([prompt], steps, hires, old)[0]
        assert got == want, (prompt, steps, hires, old)


def test_reference_doctests():
    """modules/prompt_parser.py:30-62 as written (the one doctest the reference itself no longer satisfies — the "{b|d{"
    prompt, marked "not handling this right now" — is pinned to what the reference actually returns, in the corpus)."""
    g = lambda p: P.get_learned_conditioning_prompt_schedules([p], 10)[0]  # noqa: E731
    assert g("test") == [[10, "test"]]
    assert g("a [b:3]") == [[3, "a "], [10, "a b"]]
    assert g("a [[[b]]:2]") == [[2, "a "], [10, "a [[b]]"]]
    assert g("[(a:2):3]") == [[3, ""], [10, "(a:2)"]]
    assert g("a [b : c : 1] d") == [[1, "a b  d"], [10, "a  c  d"]]
    assert g("a[b:[c:d:2]:1]e") == [[1, "abe"], [2, "ace"], [10, "ade"]]
    assert g("a [unbalanced") == [[10, "a [unbalanced"]]
    assert g("((a][:b:c [d:3]") == [[3, "((a][:b:c "], [10, "((a][:b:c d"]]
    assert g("[fe|||]male")[:5] == [[1, "female"], [2, "male"], [3, "male"], [4, "male"], [5, "female"]]
    h = lambda p: P.get_learned_conditioning_prompt_schedules([p], 10, 10)[0]  # noqa: E731
    assert h("a [b:.5] c") == [[10, "a b c"]] and h("a [b:1.5] c") == [[5, "a  c"], [10, "a b c"]]


def test_duplicate_prompts_share_one_schedule_object():
    a, b, c = P.get_learned_conditioning_prompt_schedules(["[x:y:3]", "z", "[x:y:3]"], 10)
    assert a is c and a == [[3, "x"], [10, "y"]] and b == [[10, "z"]]


class _Stub:
    def __init__(self):
        self.seen, self.calls = [], []

    def get_learned_conditioning(self, texts):
        self.calls.append([list(texts), bool(getattr(texts, "is_negative_prompt", False)), getattr(texts, "width", None), getattr(texts, "height", None)])
        for t in texts:
            if t not in self.seen:
                self.seen.append(t)
        return torch.tensor([[float(self.seen.index(t))] for t in texts])


def test_multicond_containers_match_reference():
    for ref in GOLD["multicond"]:
        stub = _Stub()
        sd = P.SdConditioning(ref["prompts"], is_negative_prompt=True, width=640, height=768)
        mc = P.get_multicond_learned_conditioning(stub, sd, ref["steps"], ref["hires_steps"])
        assert list(mc.shape) == ref["shape"] and stub.seen == ref["seen"] and stub.calls == ref["calls"]
        got = [[[cp.weight, [[s.end_at_step, int(s.cond.item())] for s in cp.schedules]] for cp in per] for per in mc.batch]
        assert got == ref["batch"]
        # what the sampler consumes: the per-step gather follows the schedule boundaries
        conds_list, stacked = P.reconstruct_multicond_batch(mc, 1)
        assert stacked.shape[0] == sum(len(per) for per in mc.batch) and len(conds_list) == len(ref["prompts"])


def test_dict_conditioning_is_split_per_schedule_entry():
    class XL:
        def get_learned_conditioning(self, texts):
            return {"crossattn": torch.arange(len(texts) * 6, dtype=torch.float32).reshape(len(texts), 3, 2), "vector": torch.arange(len(texts) * 4, dtype=torch.float32).reshape(len(texts), 4)}

    (sched,) = P.get_learned_conditioning(XL(), ["[a:b:2]"], 4)
    assert [s.end_at_step for s in sched] == [2, 4]
    assert sched[1].cond["crossattn"].shape == (3, 2) and sched[1].cond["vector"].tolist() == [4.0, 5.0, 6.0, 7.0]
    batch = P.reconstruct_cond_batch([sched], 3)
    assert batch["crossattn"].shape == (1, 3, 2) and batch.shape == (1, 3, 2)


def _cpu_model(stub):
    from sdwebui_b200.processing import SdModel

    m = SdModel(None, None, is_sdxl=False, device="cpu")
    m.cond_stage_model = stub
    return m


class _Encoder:
    """stands in for FrozenCLIPEmbedderWithCustomWords: one [3, 2] 'conditioning' per text, value = index of first sight"""

    def __init__(self):
        self.seen, self.calls = [], 0

    def __call__(self, texts):
        self.calls += 1
        for t in texts:
            if t not in self.seen:
                self.seen.append(t)
        return torch.stack([torch.full((3, 2), float(self.seen.index(t))) for t in texts])


def test_setup_conds_first_pass_hires_and_cache():
    """StableDiffusionProcessing.setup_conds / calculate_hr_conds / get_conds (modules/processing.py:460-506, 1498-1542)."""
    from sdwebui_b200.processing import StableDiffusionProcessingTxt2Img as T2I

    for cache in (T2I.cached_uc, T2I.cached_c, T2I.cached_hr_uc, T2I.cached_hr_c):
        cache[0] = cache[1] = None
    enc = _Encoder()
    model = _cpu_model(enc)
    kw = dict(sd_model=model, seeds=[1, 2], prompts=["a [red:blue:0.5] hat AND a dog :0.5", "plain"], negative_prompts=["ugly", "[x:y:15]"],
              steps=10, sampler_name="Heun", enable_hr=True, hr_second_pass_steps=6, width=64, height=96)
    p = T2I(**kw)
    p.setup_conds()
    assert p.step_multiplier == 2 and p.firstpass_steps == 20      # Heun is second order: the denoiser is called twice per step
    assert isinstance(p.c, P.MulticondLearnedConditioning) and p.c.shape == (2,)
    first = p.c.batch[0]
    assert [cp.weight for cp in first] == [1.0, 0.5]
    assert [s.end_at_step for s in first[0].schedules] == [10, 20]  # 0.5 of the 20 denoiser calls
    assert [s.end_at_step for s in p.uc[1]] == [15, 20]
    assert p.get_conds() == (p.c, p.uc)
    calls = enc.calls
    q = T2I(**kw)
    q.setup_conds()                                                  # same prompts, same steps: served from the class-level cache
    assert enc.calls == calls and q.c is p.c and q.uc is p.uc
    r = T2I(**{**kw, "steps": 12})
    r.setup_conds()
    assert enc.calls > calls and r.c is not p.c
    # second pass: whole-number boundaries count on from the first pass (20), fractions from 1.0; 6 Heun steps = 12 calls
    p.setup_conds()
    p.is_hr_pass = True
    p.calculate_hr_conds()
    hr_c, hr_uc = p.get_conds()
    assert hr_c is p.hr_c and [s.end_at_step for s in hr_c.batch[0][0].schedules] == [12]       # 0.5 < 1.0: already "blue"
    assert float(hr_c.batch[0][0].schedules[0].cond[0, 0]) == float(enc.seen.index("a blue hat"))
    assert [s.end_at_step for s in hr_uc[1]] == [12]                                           # step 15 belonged to the first pass
    p2 = T2I(**{**kw, "negative_prompts": ["ugly", "[x:y:25]"]})
    p2.setup_conds()
    p2.is_hr_pass = True
    p2.calculate_hr_conds()
    assert [s.end_at_step for s in p2.hr_uc[1]] == [5, 12]                                     # 25 - 20 = 5 calls into the second pass


def test_setup_conds_without_prompts_keeps_given_conds():
    from sdwebui_b200.processing import StableDiffusionProcessingTxt2Img as T2I

    c, u = torch.zeros(1, 3, 2), torch.ones(1, 3, 2)
    p = T2I(sd_model=_cpu_model(None), c=c, uc=u, seeds=[1])
    p.setup_conds()
    assert p.c is c and p.uc is u and p.get_conds() == (c, u)


def test_model_without_conditioner_fails_loudly():
    import pytest

    from sdwebui_b200.lib import SdxeError
    from sdwebui_b200.processing import StableDiffusionProcessingTxt2Img as T2I

    for cache in (T2I.cached_uc, T2I.cached_c):
        cache[0] = cache[1] = None
    p = T2I(sd_model=_cpu_model(None), prompts=["a"], seeds=[1])
    with pytest.raises(SdxeError):
        p.setup_conds()
