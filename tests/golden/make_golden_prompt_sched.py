"""Golden vectors for the prompt-editing schedules: the REFERENCE's own `get_learned_conditioning_prompt_schedules`
(modules/prompt_parser.py:26, lark Earley grammar) run in this container over its doctest prompts, hand-picked edge cases
and a seeded fuzz corpus (random delimiter soup + generated well-formed nestings).

    python tests/golden/make_golden_prompt_sched.py        # needs /root/reference and lark; writes prompt_sched_ref.json
"""
import json
import os
import random
import sys

sys.path.insert(0, "/root/reference")
from modules import prompt_parser as R  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
DOCTEST = ["test", "a [b:3]", "a [b: 3]", "a [[[b]]:2]", "[(a:2):3]", "a [b : c : 1] d", "a[b:[c:d:2]:1]e", "a [unbalanced", "a [b:.5] c",
           "a [{b|d{:.5] c", "((a][:b:c [d:3]", "[a|(b:1.1)]", "[fe|]male", "[fe|||]male", "a [b:1.5] c"]
EDGE = ["a | b", "[a|b:3]", "[a:b]", "(a:b:c)", "[a:b:c:3]", "x [a:3x] y", "[a\\]:3]", "a\\", "[[a:2]|b]", "[a:1e0]", "[a:+3]", "[a:-3]", "[a:3.]",
        "[ :3]", "[:3]", "[::3]", "[a::3]", "[|]", "[a|]", "[|a]", "()", "[]", "(:)", "[a:b:3] AND [c|d]", "", " ", "[a:b: 0.5 ]", "[a:b:\t7\n]",
        "(a [b:c:0.3] d:1.2)", "[(a|b):3]", "[[a|b]:[c:d:4]:6]", "a\\[b:3\\]", "[a:3][b:4][c|d|e]", "fantasy landscape with a [mountain:lake:0.25] and "
        "[an oak:a christmas tree:0.75][ in foreground::0.6][: in background:0.25] [shoddy:masterful:0.5]", "[a:b:0]", "[a:b:1]", "[a:b:0.0]",
        "[a:b:1.0]", "[a:b:99]", "[a:b:0.999]", "[a:b:.05]", "[a:b:10]", "[a:b:11]", "[a:b:20]", "[a:b:15]", "[a:b:1.25]", "[a:b:2.0]"]
SOUP = ["a", "b", "cat", "dog ", " ", "  ", "[", "]", "(", ")", ":", "|", "3", ".5", "1.5", "0.25", "12", "\\(", "\\]", "\\:", " AND ", "-1", "+2", "1e1", "{", ",", "7", "0.8"]


def soup(rng):
    return "".join(rng.choice(SOUP) for _ in range(rng.randint(1, 12)))


def wellformed(rng, depth=0):
    parts = []
    for _ in range(rng.randint(1, 4)):
        k = rng.random()
        if depth > 2 or k < 0.4:
            parts.append(rng.choice(["a", "bb ", " c", "d e", "x", " ", "\\[y\\]"]))
        elif k < 0.6:
            when = rng.choice(["3", "0.5", ".25", "8", "1.5", " 4 ", "12", "0.95", "15", "-2"])
            form = rng.randint(0, 2)
            if form == 0:
                parts.append(f"[{wellformed(rng, depth + 1)}:{when}]")
            elif form == 1:
                parts.append(f"[{wellformed(rng, depth + 1)}:{wellformed(rng, depth + 1)}:{when}]")
            else:
                parts.append(f"[{wellformed(rng, depth + 1)}::{when}]")
        elif k < 0.75:
            opts = [wellformed(rng, depth + 1) if rng.random() < 0.8 else "" for _ in range(rng.randint(2, 4))]
            parts.append("[" + "|".join(opts) + "]")
        elif k < 0.9:
            parts.append(f"({wellformed(rng, depth + 1)}:{rng.choice(['1.1', '0.8', '2'])})" if rng.random() < 0.5 else f"({wellformed(rng, depth + 1)})")
        else:
            parts.append(f"[{wellformed(rng, depth + 1)}]")
    return "".join(parts)


def main():
    rng = random.Random(20240917)
    prompts = list(DOCTEST) + list(EDGE)
    prompts += [soup(rng) for _ in range(700)]
    prompts += [wellformed(rng) for _ in range(500)]
    modes = [(10, None, False), (20, None, False), (7, None, False), (10, 10, False), (20, 12, False), (10, None, True), (20, 8, True)]
    cases = []
    for i, p in enumerate(prompts):
        for (steps, hires, old) in (modes if i < len(DOCTEST) + len(EDGE) else [modes[i % len(modes)]]):
            try:
                out = R.get_learned_conditioning_prompt_schedules([p], steps, hires, old)[0]
            except OverflowError:
                continue
            cases.append([p, steps, hires, old, out])
    # conditioning containers: the reference's get_multicond_learned_conditioning over a stub conditioner that returns,
    # for each text, a 1-vector holding that text's index in `seen` (so the structure can be stored as plain numbers)
    import torch

    class Stub:
        def __init__(self):
            self.seen, self.calls = [], []

        def get_learned_conditioning(self, texts):
            self.calls.append([list(texts), bool(getattr(texts, "is_negative_prompt", False)), getattr(texts, "width", None), getattr(texts, "height", None)])
            for t in texts:
                if t not in self.seen:
                    self.seen.append(t)
            return torch.tensor([[float(self.seen.index(t))] for t in texts])

    multi = []
    for prompts_, steps, hires in [(["a red crown", "a [blue:green:5] jeweled crown"], 20, None),
                                   (["a cat AND a dog :0.5 AND [x|y] :1.5", "a cat", "a cat AND a dog :0.5 AND [x|y] :1.5"], 6, None),
                                   (["[a:b:0.5] AND c: -1", "d:2 AND d"], 10, 8)]:
        stub = Stub()
        sd = R.SdConditioning(prompts_, is_negative_prompt=True, width=640, height=768)
        mc = R.get_multicond_learned_conditioning(stub, sd, steps, hires)
        multi.append({"prompts": prompts_, "steps": steps, "hires_steps": hires, "shape": list(mc.shape), "seen": stub.seen, "calls": stub.calls,
                      "batch": [[[cp.weight, [[s.end_at_step, int(s.cond.item())] for s in cp.schedules]] for cp in per] for per in mc.batch]})
    with open(os.path.join(HERE, "prompt_sched_ref.json"), "w") as f:
        json.dump({"source": "modules/prompt_parser.py:26 get_learned_conditioning_prompt_schedules (lark %s)" % R.lark.__version__,
                   "columns": ["prompt", "base_steps", "hires_steps", "use_old_scheduling", "schedule"], "cases": cases, "multicond": multi}, f, separators=(",", ":"))
    print(len(cases), "cases")


if __name__ == "__main__":
    main()
