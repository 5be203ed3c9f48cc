"""Generates the committed golden fixtures. Run in the build container (needs /root/reference):

    python tests/golden/make_golden.py

philox_ref.npz   outputs of the REFERENCE's own modules/rng_philox.py (imported from /root/reference) for a few
                 (seed, shape, call-index) triples: pins oracle/rng.py and the product's rng.py bit-for-bit.
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
    tiny()
