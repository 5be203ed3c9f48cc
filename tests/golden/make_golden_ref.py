"""Pins the oracle to code the REFERENCE itself ships in-tree. Run in the build container (needs /root/reference):

    python tests/golden/make_golden_ref.py        ->  tests/golden/ref_pins.npz

The hot path's arithmetic mostly lives in un-vendored upstream repositories (ldm / sgm / k-diffusion), but the reference
tree does hold its own copies of several pieces. Each is EXECUTED here unmodified (module loaded from /root/reference
with stub `modules.shared` / `ldm` / `sgm` packages, or a single function compiled from the file's AST) on seeded
inputs, and the outputs are committed; tests/test_oracle_pins_cpu.py replays the same inputs through oracle/ and through
the product's host code. Nothing here is imported by the product.

  vae_*        modules/models/sd3/sd3_impls.py:171-355   VAEDecoder / VAEEncoder (z_channels=4 == the SD1.x/SDXL KL-f8 VAE:
                                                          same key names, same ops)                     -> oracle/vae.py
  attn_*       modules/sd_hijack_optimizations.py         scaled_dot_product / Doggettx / sub-quadratic / v1 / InvokeAI
                                                          CrossAttention forwards, sdp / sub-quad / bmm AttnBlock
                                                          forwards                                        -> oracle CrossAttention, AttnBlock
  temb         modules/sd_hijack_unet.py:58-78            timestep_embedding                              -> oracle/unet.py
  st_*         modules/sd_hijack_unet.py:81-101           spatial_transformer_forward                     -> oracle SpatialTransformer
  cfg_*        modules/sd_samplers_cfg_denoiser.py:74-82  CFGDenoiser.combine_denoised (weights != 1)     -> oracle + product combine
  rng_*        modules/rng.py:85-163                      slerp + ImageRNG (subseeds, seed resize, NV/CPU) -> oracle/rng.py, product rng.py
  i2i_steps    modules/sd_samplers_common.py:22-31        setup_img2img_steps                             -> product samplers.py
"""
import ast
import importlib.util
import math
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)


def ref_nodes(relpath, names, glb, first_line=None, last_line=None):
    """Compile the named top-level defs / classes (or `Class.method`) of a reference file, unmodified, into `glb`."""
    src = open(os.path.join(REF, relpath)).read()
    tree = ast.parse(src)
    picked = []
    for n in tree.body:
        if isinstance(n, (ast.FunctionDef, ast.ClassDef)):
            if first_line is not None and not (first_line <= n.lineno <= last_line):
                continue
            if names is None or n.name in names:
                picked.append(n)
            elif isinstance(n, ast.ClassDef):
                for m in n.body:
                    if isinstance(m, ast.FunctionDef) and f"{n.name}.{m.name}" in names:
                        picked.append(m)
    mod = ast.Module(body=picked, type_ignores=[])
    exec(compile(mod, os.path.join(REF, relpath), "exec"), glb)
    return glb


def load_ref_module(relpath, name, stubs):
    """Import a reference module file under `name` with `stubs` patched into sys.modules for the duration."""
    saved = {k: sys.modules.get(k) for k in stubs}
    sys.modules.update(stubs)
    try:
        spec = importlib.util.spec_from_file_location(name, os.path.join(REF, relpath))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod


def T(x):
    return x.detach().cpu().numpy()


# ----------------------------------------------------------------------------------------------------------------------
def vae_pins(out):
    import einops

    from oracle.synth import init_module_
    from oracle.vae import AutoencoderKLDecode, AutoencoderKLEncode, VAEConfig, tiny_vae_config

    g = {"torch": torch, "math": math, "einops": einops}
    ref_nodes("modules/models/sd3/sd3_impls.py", None, g, first_line=171, last_line=355)
    for tag, cfg, hw in (("tiny", tiny_vae_config(), 16), ("full", VAEConfig(), 4)):
        dec = init_module_(AutoencoderKLDecode(cfg), 11).eval()
        enc = init_module_(AutoencoderKLEncode(cfg), 12).eval()
        rd = g["VAEDecoder"](ch=cfg.ch, out_ch=cfg.out_ch, ch_mult=tuple(cfg.ch_mult), num_res_blocks=cfg.num_res_blocks, z_channels=cfg.z_channels)
        re = g["VAEEncoder"](ch=cfg.ch, ch_mult=tuple(cfg.ch_mult), num_res_blocks=cfg.num_res_blocks, in_channels=cfg.in_channels, z_channels=cfg.z_channels)
        rd.load_state_dict({k[len("decoder."):]: v for k, v in dec.state_dict().items() if k.startswith("decoder.")}, strict=True)
        re.load_state_dict({k[len("encoder."):]: v for k, v in enc.state_dict().items() if k.startswith("encoder.")}, strict=True)
        gen = torch.Generator().manual_seed(5)
        z = torch.randn(1 if tag == "full" else 2, cfg.z_channels, hw, hw, generator=gen)
        f = 2 ** (len(cfg.ch_mult) - 1)
        x = torch.rand(z.shape[0], 3, hw * f, hw * f, generator=gen) * 2 - 1
        with torch.no_grad():
            # AutoencoderKL.decode = decoder(post_quant_conv(z)); encode = quant_conv(encoder(x)) -- the 1x1 convs around the
            # reference classes are applied with the oracle's weights through F.conv2d (no oracle module code involved)
            zq = torch.nn.functional.conv2d(z, dec.state_dict()["post_quant_conv.weight"], dec.state_dict()["post_quant_conv.bias"])
            out[f"vae_{tag}_decode"] = T(rd(zq))
            h = re(x)
            out[f"vae_{tag}_moments"] = T(torch.nn.functional.conv2d(h, enc.state_dict()["quant_conv.weight"], enc.state_dict()["quant_conv.bias"]))
    print("vae pins done")


def attention_pins(out):
    from oracle.synth import init_module_
    from oracle.unet import CrossAttention
    from oracle.vae import AttnBlock

    class _Dummy:
        def forward(self, *a, **k):
            raise NotImplementedError

    def pkg(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        return m

    opts = types.SimpleNamespace(upcast_attn=False, sub_quad_q_chunk_size=1024, sub_quad_kv_chunk_size=None, sub_quad_chunk_threshold=None)
    cmd_opts = types.SimpleNamespace(sub_quad_q_chunk_size=1024, sub_quad_kv_chunk_size=None, sub_quad_chunk_threshold=None,
                                     xformers=False, force_enable_xformers=False, opt_sdp_attention=True, opt_sdp_no_mem_attention=False,
                                     opt_sub_quad_attention=False, opt_split_attention_invokeai=False, opt_split_attention=False,
                                     opt_split_attention_v1=False, disable_opt_split_attention=False)
    shared = pkg("modules.shared", opts=opts, cmd_opts=cmd_opts, loaded_hypernetworks=[], device=torch.device("cpu"), xformers_available=False)
    hyper = pkg("modules.hypernetworks.hypernetwork", apply_hypernetworks=lambda hns, context, layer=None: (context, context))
    import contextlib

    devices = pkg("modules.devices", device=torch.device("cpu"), cpu=torch.device("cpu"), without_autocast=lambda disable=False: contextlib.nullcontext(),
                  torch_gc=lambda: None)
    sqa = load_ref_module("modules/sub_quadratic_attention.py", "ref_sub_quadratic_attention", {})
    ldm_attn = pkg("ldm.modules.attention", CrossAttention=_Dummy)
    ldm_model = pkg("ldm.modules.diffusionmodules.model", AttnBlock=_Dummy)
    stubs = {
        "ldm": pkg("ldm"), "ldm.util": pkg("ldm.util", default=lambda v, d: v if v is not None else (d() if callable(d) else d)),
        "ldm.modules": pkg("ldm.modules"), "ldm.modules.attention": ldm_attn, "ldm.modules.diffusionmodules": pkg("ldm.modules.diffusionmodules"),
        "ldm.modules.diffusionmodules.model": ldm_model,
        "sgm": pkg("sgm"), "sgm.modules": pkg("sgm.modules"), "sgm.modules.attention": pkg("sgm.modules.attention", CrossAttention=_Dummy),
        "sgm.modules.diffusionmodules": pkg("sgm.modules.diffusionmodules"),
        "sgm.modules.diffusionmodules.model": pkg("sgm.modules.diffusionmodules.model", AttnBlock=_Dummy),
        "modules": pkg("modules", shared=shared, errors=pkg("modules.errors"), devices=devices, sub_quadratic_attention=sqa),
        "modules.shared": shared, "modules.errors": pkg("modules.errors"), "modules.devices": devices,
        "modules.sub_quadratic_attention": sqa, "modules.hypernetworks": pkg("modules.hypernetworks", hypernetwork=hyper),
        "modules.hypernetworks.hypernetwork": hyper,
    }
    for parent, child in (("ldm", "modules"), ("ldm.modules", "attention"), ("ldm.modules", "diffusionmodules"), ("ldm.modules.diffusionmodules", "model"),
                          ("sgm", "modules"), ("sgm.modules", "attention"), ("sgm.modules", "diffusionmodules"), ("sgm.modules.diffusionmodules", "model")):
        setattr(stubs[parent], child, stubs[f"{parent}.{child}"])
    ref = load_ref_module("modules/sd_hijack_optimizations.py", "ref_sd_hijack_optimizations", stubs)
    # sub-quad reads shared.cmd_opts / psutil at call time: keep the stubs alive on the module
    ref.shared = shared
    gen = torch.Generator().manual_seed(21)
    # SD1.5 level-0 geometry in miniature: 8 heads x 40, 256 query tokens; cross: 77 context tokens of width 96
    for tag, (qd, cd, heads, dh, n, nk) in {"self": (320, None, 8, 40, 128, None), "cross": (320, 96, 8, 40, 128, 77),
                                            "sdxl": (128, 64, 2, 64, 64, 77)}.items():
        m = init_module_(CrossAttention(qd, cd, heads, dh), 31).eval()
        x = torch.randn(2, n, qd, generator=gen)
        ctx = None if cd is None else torch.randn(2, nk, cd, generator=gen)
        with torch.no_grad():
            sdp = ref.scaled_dot_product_attention_forward(m, x, context=ctx)
            out[f"attn_{tag}_sdp"] = T(sdp)
            # the other variants the reference ships compute the same function: keep only their max deviation from SDP
            # (what the sdxe attention seam replaces is ALL of them)
            devs = []
            for name in ("split_cross_attention_forward", "sub_quad_attention_forward", "split_cross_attention_forward_v1",
                         "split_cross_attention_forward_invokeAI"):
                devs.append(float((getattr(ref, name)(m, x, context=ctx) - sdp).abs().max() / sdp.abs().max()))
            out[f"attn_{tag}_variant_dev"] = np.array(devs)
    ab = init_module_(AttnBlock(64), 32).eval()
    xa = torch.randn(2, 64, 12, 12, generator=gen)
    with torch.no_grad():
        sdp = ref.sdp_attnblock_forward(ab, xa)
        out["attn_block_sdp"] = T(sdp)
        out["attn_block_variant_dev"] = np.array([float((getattr(ref, name)(ab, xa) - sdp).abs().max() / sdp.abs().max())
                                                  for name in ("cross_attention_attnblock_forward", "sub_quad_attnblock_forward")])
    print("attention pins done")


def unet_piece_pins(out):
    from oracle.synth import init_module_
    from oracle.unet import SpatialTransformer

    g = {"torch": torch, "math": math}
    ref_nodes("modules/sd_hijack_unet.py", ["timestep_embedding", "spatial_transformer_forward"], g)
    t = torch.tensor([0.0, 1.0, 17.5, 500.25, 999.0])
    out["temb_320"] = T(g["timestep_embedding"](None, t, 320))
    out["temb_256"] = T(g["timestep_embedding"](None, t.half(), 256))  # the fp16 `t` the apply_model patch hands over
    gen = torch.Generator().manual_seed(41)
    for tag, linear in (("conv", False), ("linear", True)):
        # depth 1: the patch is installed on ldm's SpatialTransformer only (sd_hijack_unet.py:127; SD1.x / SD2.x have depth 1),
        # and indexes `context[i]` of a one-element list; sgm (SDXL) keeps its own forward (same arithmetic, context shared)
        st = init_module_(SpatialTransformer(64, 2, 32, 1, 48, linear), 42).eval()
        x = torch.randn(2, 64, 8, 8, generator=gen)
        ctx = torch.randn(2, 77, 48, generator=gen)
        with torch.no_grad():
            out[f"st_{tag}"] = T(g["spatial_transformer_forward"](None, st, x, context=ctx))
    print("unet piece pins done")


def cfg_pins(out):
    g = {"torch": torch}
    ref_nodes("modules/sd_samplers_cfg_denoiser.py", ["CFGDenoiser.combine_denoised"], g)
    gen = torch.Generator().manual_seed(51)
    # 3 images; image 0 has two AND-composed conds (weights 0.7 / 1.3), image 1 one cond of weight 1, image 2 weight -0.5
    conds_list = [[(0, 0.7), (1, 1.3)], [(2, 1.0)], [(3, -0.5)]]
    x_out = torch.randn(4 + 3, 4, 8, 8, generator=gen)
    uncond = torch.zeros(3, 77, 8)
    out["cfg_combine"] = T(g["combine_denoised"](None, x_out, conds_list, uncond, 7.5))
    print("cfg pins done")


def rng_pins(out):
    def pkg(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        return m

    philox = load_ref_module("modules/rng_philox.py", "ref_rng_philox", {})
    for source in ("NV", "CPU"):
        opts = types.SimpleNamespace(randn_source=source, eta_noise_seed_delta=0)
        devices = pkg("modules.devices", device=torch.device("cpu"), cpu=torch.device("cpu"))
        shared = pkg("modules.shared", opts=opts, device=torch.device("cpu"))
        stubs = {"modules": pkg("modules", devices=devices, rng_philox=philox, shared=shared), "modules.devices": devices,
                 "modules.rng_philox": philox, "modules.shared": shared}
        ref = load_ref_module("modules/rng.py", "ref_rng", stubs)
        shape = (4, 8, 8)
        cases = {
            "plain": dict(),
            "sub": dict(subseeds=[7, 8, 9], subseed_strength=0.3),
            "sub_short": dict(subseeds=[7], subseed_strength=0.65),           # fewer subseeds than images -> subseed 0
            "resize_small": dict(seed_resize_from_h=48, seed_resize_from_w=32),   # source 6x4 latent into 8x8
            "resize_large": dict(seed_resize_from_h=96, seed_resize_from_w=80),   # source 12x10 latent cropped to 8x8
            "sub_resize": dict(subseeds=[3, 4, 5], subseed_strength=0.5, seed_resize_from_h=48, seed_resize_from_w=96),
        }
        for tag, kw in cases.items():
            r = ref.ImageRNG(shape, [1000, 1001, 1002], **kw)
            out[f"rng_{source}_{tag}_first"] = T(r.next())
            out[f"rng_{source}_{tag}_next"] = T(r.next())
        opts.eta_noise_seed_delta = 31337
        r = ref.ImageRNG(shape, [1000, 1001], subseeds=[1, 2], subseed_strength=0.1)
        out[f"rng_{source}_ensd_first"] = T(r.next())
        out[f"rng_{source}_ensd_next"] = T(r.next())
    lo, hi = torch.randn(3, 200, generator=torch.Generator().manual_seed(61)), torch.randn(3, 200, generator=torch.Generator().manual_seed(62))
    out["rng_slerp"] = T(ref.slerp(0.25, lo, hi))
    out["rng_slerp_close"] = T(ref.slerp(0.25, lo, lo * 1.00001))
    print("rng pins done")


def img2img_step_pins(out):
    g = {"opts": types.SimpleNamespace(img2img_fix_steps=False)}
    ref_nodes("modules/sd_samplers_common.py", ["setup_img2img_steps"], g)
    rows = []
    for steps in (1, 7, 20, 50):
        for ds in (0.0, 0.05, 0.3, 0.75, 1.0):
            p = types.SimpleNamespace(steps=steps, denoising_strength=ds)
            rows.append([steps, ds, -1, *g["setup_img2img_steps"](p)])
            for hr in (0, 10):
                rows.append([steps, ds, hr, *g["setup_img2img_steps"](p, hr)])
    out["i2i_steps"] = np.array(rows, dtype=np.float64)
    print("img2img step pins done")


if __name__ == "__main__":
    torch.manual_seed(0)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    out = {}
    vae_pins(out)
    attention_pins(out)
    unet_piece_pins(out)
    cfg_pins(out)
    rng_pins(out)
    img2img_step_pins(out)
    path = os.path.join(HERE, "ref_pins.npz")
    np.savez_compressed(path, **out)
    print(path, len(out), "arrays,", os.path.getsize(path) // 1024, "KiB")
