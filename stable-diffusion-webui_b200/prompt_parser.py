"""Prompt emphasis syntax -> (text, weight) runs: host mirror of `modules/prompt_parser.py:352-458 parse_prompt_attention`
(groundwork for SURVEY §8(f) row N4; the weights become the per-token multipliers of `sd_hijack_clip.process_tokens`).

    (abc)        weight x 1.1          [abc]     weight / 1.1         (abc:3.12)   weight x 3.12
    \\( \\) \\[ \\] \\\\   literal characters       BREAK (a word of its own)  ->  ["BREAK", -1] separator

Pinned bit-for-bit against the reference function imported from /root/reference (tests/golden/prompt_attention.json,
generator tests/golden/make_golden.py).
"""
from __future__ import annotations

import re
from typing import List

_ROUND, _SQUARE = 1.1, 1 / 1.1
_ESCAPABLE = "()[]\\"
_WEIGHT = re.compile(r":\s*([+-]?[.\d]+)\s*\)")
_BREAK = re.compile(r"\s*\bBREAK\b\s*", re.S)


def parse_prompt_attention(text: str) -> List[list]:
    runs: List[list] = []
    open_round: List[int] = []   # index into `runs` where each still-open ( started
    open_square: List[int] = []

    def scale_from(start: int, factor: float):
        for r in runs[start:]:
            r[1] *= factor

    def add_text(piece: str):
        parts = _BREAK.split(piece)
        for k, part in enumerate(parts):
            if k:
                runs.append(["BREAK", -1])
            runs.append([part, 1.0])

    i, n = 0, len(text)
    while i < n:
        ch = text[i]
        if ch == "\\":
            if i + 1 < n and text[i + 1] in _ESCAPABLE:
                runs.append([text[i + 1], 1.0])
                i += 2
            else:  # a lone backslash is dropped (it matches the pattern's bare "\\" alternative: text[1:] == "")
                runs.append(["", 1.0])
                i += 1
        elif ch == "(":
            open_round.append(len(runs))
            i += 1
        elif ch == "[":
            open_square.append(len(runs))
            i += 1
        elif ch == ":":
            m = _WEIGHT.match(text, i)
            if m:
                if open_round:
                    scale_from(open_round.pop(), float(m.group(1)))
                else:  # ":1.2)" with nothing open is plain text
                    add_text(m.group(0))
                i = m.end()
            else:
                add_text(":")
                i += 1
        elif ch == ")":
            if open_round:
                scale_from(open_round.pop(), _ROUND)
            else:
                add_text(")")
            i += 1
        elif ch == "]":
            if open_square:
                scale_from(open_square.pop(), _SQUARE)
            else:
                add_text("]")
            i += 1
        else:
            j = i
            while j < n and text[j] not in "\\()[]:":
                j += 1
            add_text(text[i:j])
            i = j
    for start in open_round:
        scale_from(start, _ROUND)
    for start in open_square:
        scale_from(start, _SQUARE)
    if not runs:
        runs = [["", 1.0]]
    # merge neighbours of equal weight (BREAK separators have weight -1 and merge like any other run, as upstream)
    k = 0
    while k + 1 < len(runs):
        if runs[k][1] == runs[k + 1][1]:
            runs[k][0] += runs[k + 1][0]
            runs.pop(k + 1)
        else:
            k += 1
    return runs


# ----------------------------------------------------------------------------------------------------------------------
# Conditioning containers handed to the samplers (modules/prompt_parser.py:140,242-349). `p.c` is a
# MulticondLearnedConditioning (one list of AND-composed, weighted, step-scheduled conds per image), `p.uc` a plain list
# of schedules; CFGDenoiser.forward rebuilds the per-step tensors from them with the two reconstruct_* functions.
# ----------------------------------------------------------------------------------------------------------------------
import collections  # noqa: E402

import torch  # noqa: E402

ScheduledPromptConditioning = collections.namedtuple("ScheduledPromptConditioning", ["end_at_step", "cond"])


class ComposableScheduledPromptConditioning:
    def __init__(self, schedules, weight=1.0):
        self.schedules = schedules   # list[ScheduledPromptConditioning]
        self.weight = weight


class MulticondLearnedConditioning:
    def __init__(self, shape, batch):
        self.shape = shape           # (number of prompts,) — what DDIM / PLMS look at
        self.batch = batch           # list (per image) of list[ComposableScheduledPromptConditioning]


class DictWithShape(dict):
    """SDXL conditioning {"crossattn": [B,T,2048], "vector": [B,2816]} that still answers `.shape` (:269-277)."""

    def __init__(self, x, shape=None):
        super().__init__()
        self.update(x)

    @property
    def shape(self):
        return self["crossattn"].shape


def _active(schedule, step):
    """index of the first entry whose end_at_step has not passed (entry 0 when all have, as the reference)."""
    for idx, entry in enumerate(schedule):
        if step <= entry.end_at_step:
            return idx
    return 0


def reconstruct_cond_batch(c, current_step):
    """:280-303 — list (per image) of schedules -> [B, T, C] tensor (or DictWithShape of them) for this step."""
    proto = c[0][0].cond
    picked = [sched[_active(sched, current_step)].cond for sched in c]
    if isinstance(proto, dict):
        return DictWithShape({k: torch.stack([p[k] for p in picked]).to(device=v.device, dtype=v.dtype) for k, v in proto.items()})
    return torch.stack(picked).to(device=proto.device, dtype=proto.dtype)


def stack_conds(tensors):
    """:306-317 — conds of different token counts are padded by repeating their last vector."""
    tensors = list(tensors)
    longest = max(t.shape[0] for t in tensors)
    for i, t in enumerate(tensors):
        if t.shape[0] != longest:
            tensors[i] = torch.vstack([t, t[-1:].repeat([longest - t.shape[0], 1])])
    return torch.stack(tensors)


def reconstruct_multicond_batch(c: MulticondLearnedConditioning, current_step):
    """:321-349 -> (conds_list, stacked): conds_list[i] = [(row in `stacked`, weight), ...] for image i."""
    proto = c.batch[0][0].schedules[0].cond
    rows, conds_list = [], []
    for composable_prompts in c.batch:
        mine = []
        for cp in composable_prompts:
            mine.append((len(rows), cp.weight))
            rows.append(cp.schedules[_active(cp.schedules, current_step)].cond)
        conds_list.append(mine)
    if isinstance(rows[0], dict):
        return conds_list, DictWithShape({k: stack_conds([r[k] for r in rows]) for k in rows[0].keys()})
    return conds_list, stack_conds(rows).to(device=proto.device, dtype=proto.dtype)


_AND = re.compile(r"\bAND\b")
_AND_WEIGHT = re.compile(r"^((?:\s|.)*?)(?:\s*:\s*([-+]?(?:\d+\.?|\d*\.\d+)))?\s*$")


def get_multicond_prompt_list(prompts):
    """:208-239 — split every prompt on the word AND, peel an optional `:weight` off each part, deduplicate the texts.
    -> (per-prompt [(flat index, weight)...], flat list of distinct texts, text -> flat index)."""
    per_prompt, flat, index_of = [], [], {}
    for prompt in prompts:
        parts = []
        for sub in _AND.split(prompt):
            m = _AND_WEIGHT.search(sub)
            text, weight = m.groups() if m is not None else (sub, 1.0)
            weight = float(weight) if weight is not None else 1.0
            if text not in index_of:
                index_of[text] = len(flat)
                flat.append(text)
            parts.append((index_of[text], weight))
        per_prompt.append(parts)
    return per_prompt, flat, index_of
