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


# ----------------------------------------------------------------------------------------------------------------------
# Prompt editing: "[from:to:when]", "[to:when]", "[from::when]", alternation "[a|b|...]" -> per-prompt step schedules
# (modules/prompt_parser.py:14-147), then prompt -> conditioning schedules (:155-203, :246-262).
#
# The reference parses with a lark (Earley) grammar. Restated here as a hand-written packrat parser of the same language —
# no parser generator on the product path:
#   start     : (prompt | one of the characters [ ] ( ) : )*                 stray delimiters are legal at top level only
#   prompt    : (emphasized | scheduled | alternate | plain)*
#   emphasized: "(" prompt ")" | "(" prompt ":" prompt ")" | "[" prompt "]"   (kept verbatim in the output text)
#   scheduled : "[" [prompt ":"] prompt ":" ws? NUMBER ws? "]"
#   alternate : "[" prompt ("|" prompt?)+ "]"
#   plain     : runs of characters other than \ [ ] ( ) : |  or backslash-escaped characters (whitespace included)
# A bracket that cannot be completed as one of the constructs is a stray delimiter; a bare "|" (or a trailing backslash)
# makes the prompt unparseable and it is then used unscheduled, as the reference does on a parse error. Pinned against the
# reference's doctests (:30-62) and a fuzz corpus run through the reference itself (tests/golden/prompt_sched_ref.json).
# ----------------------------------------------------------------------------------------------------------------------
_PLAIN = re.compile(r"(?:[^\\\[\]():|]|\\.)+")
_WS = re.compile(r"\s+")
_SIGNED_NUMBER = re.compile(r"[+-]?(?:\d+\.\d*(?:[eE][+-]?\d+)?|\.\d+(?:[eE][+-]?\d+)?|\d+[eE][+-]?\d+|\d+)")
_STRAY = "[]():"


class _ScheduleSyntax:
    """Packrat recursive descent over one prompt. Nodes: str (verbatim text), ("seq", [nodes]),
    ("sched", before | None, after, number_text), ("alt", [option | None, ...])."""

    def __init__(self, text: str):
        self.text = text
        self.memo = {}

    def prompt(self, pos: int):
        """Longest run of elements starting at pos -> (("seq", nodes), end). Never fails (may be empty)."""
        nodes = []
        text = self.text
        while pos < len(text):
            ch = text[pos]
            if ch == "(" or ch == "[":
                got = self.construct(pos)
                if got is None:
                    break
                node, pos = got
                nodes.append(node)
                continue
            m = _PLAIN.match(text, pos)
            if m is None:
                break
            nodes.append(m.group(0))
            pos = m.end()
        return ("seq", nodes), pos

    def construct(self, pos: int):
        if pos not in self.memo:
            self.memo[pos] = self._paren(pos) if self.text[pos] == "(" else self._bracket(pos)
        return self.memo[pos]

    def _at(self, pos: int) -> str:
        return self.text[pos] if pos < len(self.text) else ""

    def _paren(self, pos: int):
        inner, p = self.prompt(pos + 1)
        if self._at(p) == ")":
            return ("seq", ["(", inner, ")"]), p + 1
        if self._at(p) == ":":
            second, q = self.prompt(p + 1)
            if self._at(q) == ")":
                return ("seq", ["(", inner, ":", second, ")"]), q + 1
        return None

    def _when(self, pos: int):
        """ws? NUMBER ws? "]" -> (number text, end) | None"""
        m = _WS.match(self.text, pos)
        if m:
            pos = m.end()
        m = _SIGNED_NUMBER.match(self.text, pos)
        if m is None:
            return None
        number, pos = m.group(0), m.end()
        m = _WS.match(self.text, pos)
        if m:
            pos = m.end()
        return (number, pos + 1) if self._at(pos) == "]" else None

    def _bracket(self, pos: int):
        first, p = self.prompt(pos + 1)
        ch = self._at(p)
        if ch == "]":
            return ("seq", ["[", first, "]"]), p + 1
        if ch == "|":
            options = [first if first[1] else None]
            while self._at(p) == "|":
                option, p = self.prompt(p + 1)
                options.append(option if option[1] else None)
            return (("alt", options), p + 1) if self._at(p) == "]" else None
        if ch == ":":
            when = self._when(p + 1)
            if when is not None:                       # "[to:when]"
                return ("sched", None, first, when[0]), when[1]
            second, q = self.prompt(p + 1)
            if self._at(q) == ":":
                when = self._when(q + 1)
                if when is not None:                   # "[from:to:when]"
                    return ("sched", first if first[1] else None, second, when[0]), when[1]
        return None

    def parse(self):
        """-> top-level node list, or None when the prompt is not in the language."""
        nodes, pos, text = [], 0, self.text
        while pos < len(text):
            (_, part), pos = self.prompt(pos)
            nodes.extend(part)
            if pos >= len(text):
                break
            if text[pos] in _STRAY:
                nodes.append(text[pos])
                pos += 1
            else:
                return None
        return nodes


def get_learned_conditioning_prompt_schedules(prompts, base_steps, hires_steps=None, use_old_scheduling=False):
    """modules/prompt_parser.py:26-147: every prompt -> [[end_at_step, text], ...]. Whole-number `when` is a step, a `when`
    with a decimal point a fraction of the pass; in the hires pass (hires_steps given, new scheduling) steps count on from
    base_steps and fractions from 1.0."""
    if hires_steps is None or use_old_scheduling:
        int_offset, flt_offset, steps = 0, 0.0, base_steps
    else:
        int_offset, flt_offset, steps = base_steps, 1.0, hires_steps

    def when_step(number: str) -> int:
        v = float(number)
        if use_old_scheduling:
            v = v * steps if v < 1 else v
        elif "." in number:
            v = (v - flt_offset) * steps
        else:
            v = v - int_offset
        return min(steps, int(v))

    def boundaries(node, out):
        if isinstance(node, str) or node is None:
            return
        if node[0] == "seq":
            for child in node[1]:
                boundaries(child, out)
        elif node[0] == "sched":
            w = when_step(node[3])
            if w >= 1:
                out.add(w)
            boundaries(node[1], out)
            boundaries(node[2], out)
        else:  # alt
            out.update(range(1, steps + 1))
            for option in node[1]:
                boundaries(option, out)

    def render(node, step, out):
        if node is None:
            return
        if isinstance(node, str):
            out.append(node)
        elif node[0] == "seq":
            for child in node[1]:
                render(child, step, out)
        elif node[0] == "sched":
            render(node[1] if step <= when_step(node[3]) else node[2], step, out)
        else:
            render(node[1][(step - 1) % len(node[1])], step, out)

    def schedule_of(prompt):
        nodes = _ScheduleSyntax(prompt).parse()
        if nodes is None:
            return [[steps, prompt]]
        tree = ("seq", nodes)
        ends = {steps}
        boundaries(tree, ends)
        result = []
        for t in sorted(ends):
            out = []
            render(tree, t, out)
            result.append([t, "".join(out)])
        return result

    by_prompt = {prompt: schedule_of(prompt) for prompt in set(prompts)}
    return [by_prompt[prompt] for prompt in prompts]


class SdConditioning(list):
    """modules/prompt_parser.py:140-152: the prompts handed to the conditioner, plus what SDXL needs to know about the job."""

    def __init__(self, prompts, is_negative_prompt=False, width=None, height=None, copy_from=None):
        super().__init__()
        self.extend(prompts)
        if copy_from is None:
            copy_from = prompts
        self.is_negative_prompt = is_negative_prompt or getattr(copy_from, "is_negative_prompt", False)
        self.width = width or getattr(copy_from, "width", None)
        self.height = height or getattr(copy_from, "height", None)


def get_learned_conditioning(model, prompts, steps, hires_steps=None, use_old_scheduling=False):
    """:155-203 — prompts -> list (per prompt) of [ScheduledPromptConditioning(end_at_step, cond), ...]; all texts of one
    prompt's schedule are encoded in one `model.get_learned_conditioning` call, equal prompts share their schedule."""
    res, cache = [], {}
    for prompt, schedule in zip(prompts, get_learned_conditioning_prompt_schedules(prompts, steps, hires_steps, use_old_scheduling)):
        if prompt in cache:
            res.append(cache[prompt])
            continue
        conds = model.get_learned_conditioning(SdConditioning([text for _, text in schedule], copy_from=prompts))
        cond_schedule = []
        for i, (end_at_step, _) in enumerate(schedule):
            cond = {k: v[i] for k, v in conds.items()} if isinstance(conds, dict) else conds[i]
            cond_schedule.append(ScheduledPromptConditioning(end_at_step, cond))
        cache[prompt] = cond_schedule
        res.append(cond_schedule)
    return res


def get_multicond_learned_conditioning(model, prompts, steps, hires_steps=None, use_old_scheduling=False) -> MulticondLearnedConditioning:
    """:246-262 — AND-split prompts -> MulticondLearnedConditioning (what `p.c` is)."""
    per_prompt, flat, _ = get_multicond_prompt_list(prompts)
    flat = SdConditioning(flat, copy_from=prompts)
    learned = get_learned_conditioning(model, flat, steps, hires_steps, use_old_scheduling)
    batch = [[ComposableScheduledPromptConditioning(learned[i], weight) for i, weight in parts] for parts in per_prompt]
    return MulticondLearnedConditioning(shape=(len(prompts),), batch=batch)
