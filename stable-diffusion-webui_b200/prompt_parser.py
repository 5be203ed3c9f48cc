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
