"""Generates the committed golden fixtures. Run in the build container (needs /root/reference):

    python tests/golden/make_golden.py

philox_ref.npz   outputs of the REFERENCE's own modules/rng_philox.py (imported from /root/reference) for a few
                 (seed, shape, call-index) triples: pins oracle/rng.py and the product's rng.py bit-for-bit.
prompt_attention.json  outputs of the REFERENCE's modules/prompt_parser.py parse_prompt_attention (module loaded from
                 /root/reference) for hand-written and fuzzed prompts: pins sdwebui_b200/prompt_parser.py exactly.
tiny_oracle.npz  outputs of the oracle itself on the tiny UNet / VAE / samplers with seeded weights (regression
                 fixture: guards the oracle against accidental edits; it is NOT an external ground truth).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def philox():
    sys.path.insert(0, "/root/reference")
    from modules import rng_philox  # the reference's module, unmodified

    out = {}
    for seed in (0, 1000, 123456789012, 2 ** 32 + 5):
        g = rng_philox.Generator(seed)
        for call in range(3):
            out[f"s{seed}_c{call}"] = g.randn((4, 8, 8))
    g = rng_philox.Generator(0)
    out["doc_3x4"] = g.randn((3, 4))
    np.savez_compressed(os.path.join(HERE, "philox_ref.npz"), **out)
    print("philox_ref.npz", len(out))


def prompt_attention():
    import importlib.util
    import json
    import random
    import warnings

    warnings.simplefilter("ignore")
    spec = importlib.util.spec_from_file_location("ref_prompt_parser", "/root/reference/modules/prompt_parser.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)  # the reference's module, unmodified
    prompts = ["", "normal text", "an (important) word", "(unbalanced", r"\(literal\]", "(unnecessary)(parens)",
               "a (((house:1.3)) [on] a (hill:0.5), sun, (((sky))).", "a cat BREAK a dog", "(a BREAK b:1.4) c", "x BREAKING y",
               "[[deep] er] and ((more:0.25) text)", "colon: alone :1.2) ) ] closing", "(neg:-0.5) (sp : 1.5 ) (plus:+2)",
               r"back\\slash \ lone \x", "masterpiece, (best quality:1.2), [lowres], ((detailed face)), 8k",
               "(((", "]]]", "(a:1.1", "(a:)", "a:b:c", "(::1.5)", "[a:b:0.5]", "((x):2)(y:3)", " lead and trail "]
    rnd = random.Random(7)
    alphabet = ["(", ")", "[", "]", ":", "\\", " ", "a", "bc", "1.2", ".5", "BREAK", ",", "-", "+"]
    for _ in range(300):
        prompts.append("".join(rnd.choice(alphabet) for _ in range(rnd.randint(1, 24))))
    out = []
    for p in prompts:
        try:
            out.append({"prompt": p, "result": ref.parse_prompt_attention(p)})
        except Exception as e:  # e.g. float("..") — the restatement must fail the same way
            out.append({"prompt": p, "error": type(e).__name__})
    with open(os.path.join(HERE, "prompt_attention.json"), "w") as f:
        json.dump(out, f, indent=0)
    print("prompt_attention.json", len(out), sum(1 for o in out if "error" in o), "error cases")


def tiny():
    from oracle.pipeline import OraclePipeline, SamplingParams
    from oracle.synth import init_module_, synthetic_context
    from oracle.unet import UNetModel, tiny_config
    from oracle.vae import AutoencoderKLDecode, tiny_vae_config

    torch.manual_seed(0)
    torch.set_num_threads(1)
    ucfg, vcfg = tiny_config(), tiny_vae_config()
    unet = init_module_(UNetModel(ucfg), 1).eval()
    vae = init_module_(AutoencoderKLDecode(vcfg), 2).eval()
    c, u = synthetic_context(2, 77, ucfg.context_dim, 3), synthetic_context(2, 77, ucfg.context_dim, 4)
    x = torch.randn(2, 4, 16, 16, generator=torch.Generator().manual_seed(5))
    out = {}
    with torch.no_grad():
        out["unet_eps"] = unet(x, torch.tensor([801.0, 37.5]), context=c).numpy()
        out["vae_img"] = vae.decode(x[:1]).numpy()
    pipe = OraclePipeline(unet, vae, "cpu")
    for name, sampler, steps in (("euler_a", "Euler a", 4), ("dpmpp_2m", "DPM++ 2M", 5)):
        lat = pipe.sample(SamplingParams(sampler=sampler, steps=steps, width=128, height=128, seeds=(1000, 1001), randn_source="NV"), c, u)
        out[f"latent_{name}"] = lat.numpy()
    np.savez_compressed(os.path.join(HERE, "tiny_oracle.npz"), **out)
    print("tiny_oracle.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    philox()
    prompt_attention()
    tiny()
