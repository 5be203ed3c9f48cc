"""LoRA for a weight-snapshotting UNet.

The reference patches `nn.Linear/Conv2d.forward` and merges `weight += updown` lazily inside the stock modules
(`extensions-builtin/Lora/networks.py:391-545`, `network.py:167-214`, `network_lora.py:66-84`); an `SdUnet` replacement
that snapshotted its weights at `activate()` never sees that. Here the same deltas are merged into the UNet state dict
*before* the engine ingests it (`SdxeUnet(..., loras=[(lora_state_dict, multiplier)])`), so the packed kernel layouts
(stacked q|k|v, GEGLU interleave, LayerNorm folds) are built from the merged weights.

Covered: kohya / diffusers-named and compvis-named LoRA for the UNet — Linear, conv1x1, conv3x3 (incl. the `lora_mid`
CP decomposition), `alpha` / `scale`, `lora_A` / `lora_B` naming, per-network UNet multiplier. Text-encoder parts are
reported and skipped (text encoders are upstream of the path); LoHa / LoKr / IA3 / OFT / DoRA / bias deltas raise.
"""
from __future__ import annotations

import re
from typing import Dict, Iterable, List, Tuple

import torch

from . import lib as L

_SUFFIX = {  # networks.py:43-53
    "attentions": {},
    "resnets": {"conv1": "in_layers_2", "conv2": "out_layers_3", "norm1": "in_layers_0", "norm2": "out_layers_0",
                "time_emb_proj": "emb_layers_1", "conv_shortcut": "skip_connection"},
}


def convert_diffusers_name_to_compvis(key: str) -> str:
    """UNet half of networks.py:56-119: 'lora_unet_down_blocks_0_attentions_0_…' -> 'diffusion_model_input_blocks_1_1_…'.
    Keys that are already compvis-style ('lora_unet_input_blocks_…', SDXL kohya files) or belong to a text encoder are
    returned unchanged and resolved (or not) by the caller."""
    m = re.match(r"lora_unet_conv_in(.*)", key)
    if m:
        return f"diffusion_model_input_blocks_0_0{m.group(1)}"
    m = re.match(r"lora_unet_conv_out(.*)", key)
    if m:
        return f"diffusion_model_out_2{m.group(1)}"
    m = re.match(r"lora_unet_time_embedding_linear_(\d+)(.*)", key)
    if m:
        return f"diffusion_model_time_embed_{int(m.group(1)) * 2 - 2}{m.group(2)}"
    m = re.match(r"lora_unet_down_blocks_(\d+)_(attentions|resnets)_(\d+)_(.+)", key)
    if m:
        a, kind, b, rest = int(m.group(1)), m.group(2), int(m.group(3)), m.group(4)
        return f"diffusion_model_input_blocks_{1 + a * 3 + b}_{1 if kind == 'attentions' else 0}_{_SUFFIX[kind].get(rest, rest)}"
    m = re.match(r"lora_unet_mid_block_(attentions|resnets)_(\d+)_(.+)", key)
    if m:
        kind, b, rest = m.group(1), int(m.group(2)), m.group(3)
        return f"diffusion_model_middle_block_{1 if kind == 'attentions' else b * 2}_{_SUFFIX[kind].get(rest, rest)}"
    m = re.match(r"lora_unet_up_blocks_(\d+)_(attentions|resnets)_(\d+)_(.+)", key)
    if m:
        a, kind, b, rest = int(m.group(1)), m.group(2), int(m.group(3)), m.group(4)
        return f"diffusion_model_output_blocks_{a * 3 + b}_{1 if kind == 'attentions' else 0}_{_SUFFIX[kind].get(rest, rest)}"
    m = re.match(r"lora_unet_down_blocks_(\d+)_downsamplers_0_conv", key)
    if m:
        return f"diffusion_model_input_blocks_{3 + int(m.group(1)) * 3}_0_op"
    m = re.match(r"lora_unet_up_blocks_(\d+)_upsamplers_0_conv", key)
    if m:
        a = int(m.group(1))
        return f"diffusion_model_output_blocks_{2 + a * 3}_{2 if a > 0 else 1}_conv"
    return key


def network_layer_mapping(weight_keys: Iterable[str]) -> Dict[str, str]:
    """What `assign_network_names_to_compvis_modules` (networks.py:122-146) yields for `sd_model.model`: module path with
    dots -> underscores, prefixed 'diffusion_model_'. Values here are the module's weight key in the UNet state dict."""
    out = {}
    for k in weight_keys:
        if k.endswith(".weight"):
            out["diffusion_model_" + k[: -len(".weight")].replace(".", "_")] = k
    return out


def match_lora_keys(lora_sd: Dict[str, torch.Tensor], mapping: Dict[str, str]):
    """networks.py:181-240 for the UNet: group the file's tensors per target module. -> (matched, skipped_te, unmatched)"""
    matched: Dict[str, Dict[str, torch.Tensor]] = {}
    skipped_te: List[str] = []
    unmatched: List[str] = []
    for key_network, w in lora_sd.items():
        head, _, part = key_network.partition(".")
        if head.startswith(("lora_te", "lora_te1", "lora_te2")):
            skipped_te.append(key_network)
            continue
        key = convert_diffusers_name_to_compvis(head)
        target = mapping.get(key)
        if target is None and "lora_unet" in head:  # SDXL-style files already carry compvis names (networks.py:218-221)
            target = mapping.get(head.replace("lora_unet", "diffusion_model"))
        if target is None:
            unmatched.append(key_network)
            continue
        matched.setdefault(target, {})[part] = w
    return matched, skipped_te, unmatched


def calc_updown(parts: Dict[str, torch.Tensor], orig_shape: Tuple[int, ...]) -> torch.Tensor:
    """network_lora.py:66-84 + network.py:167-214 (without the multiplier): the weight delta of one module, fp32."""
    if "lora_A.weight" in parts and "lora_B.weight" in parts:  # network_lora.py:15-20
        parts = dict(parts)
        parts["lora_up.weight"], parts["lora_down.weight"] = parts.pop("lora_B.weight"), parts.pop("lora_A.weight")
    if not ("lora_up.weight" in parts and "lora_down.weight" in parts):
        raise L.SdxeError("unsupported network module (only LoRA up/down is implemented; got " + ", ".join(sorted(parts)) + ")")
    for bad in ("dora_scale", "bias", "diff_b"):
        if bad in parts:
            raise L.SdxeError(f"LoRA module carries '{bad}': not implemented")
    up, down = parts["lora_up.weight"].float(), parts["lora_down.weight"].float()
    dim = down.shape[0]
    if "lora_mid.weight" in parts:  # CP decomposition, lyco_helpers.py:18-21
        mid = parts["lora_mid.weight"].float()
        updown = torch.einsum("n m k l, i n, m j -> i j k l", mid, up.reshape(up.shape[0], -1), down.reshape(dim, -1))
    else:  # lyco_helpers.py:9-15
        shape = [up.shape[0], down.shape[1]] + (list(down.shape[2:]) if down.ndim == 4 else [])
        updown = (up.reshape(up.shape[0], -1) @ down.reshape(dim, -1)).reshape(shape)
    if updown.numel() != int(torch.tensor(orig_shape).prod()):
        raise L.SdxeError(f"LoRA delta {tuple(updown.shape)} does not fit the target weight {tuple(orig_shape)}")
    updown = updown.reshape(orig_shape)
    if "scale" in parts:
        scale = float(parts["scale"])
    elif "alpha" in parts:
        scale = float(parts["alpha"]) / dim
    else:
        scale = 1.0
    return updown * scale


def merge_loras(unet_sd: Dict[str, torch.Tensor], loras: Iterable[Tuple[Dict[str, torch.Tensor], float]]):
    """-> (new state dict sharing untouched tensors, one report per network). All networks' deltas of a weight are
    summed in fp32 and the weight is rounded ONCE to its storage dtype (the reference adds each network's delta in
    the module dtype, networks.py:433-441)."""
    mapping = network_layer_mapping(unet_sd.keys())
    total: Dict[str, torch.Tensor] = {}
    reports = []
    for lora_sd, mult in loras:
        matched, skipped_te, unmatched = match_lora_keys(lora_sd, mapping)
        for target, parts in matched.items():
            w = unet_sd[target]
            delta = float(mult) * calc_updown(parts, tuple(w.shape)).to(w.device)
            total[target] = delta if target not in total else total[target] + delta
        reports.append({"merged": sorted(matched), "skipped_text_encoder": skipped_te, "unmatched": unmatched})
    out = dict(unet_sd)
    for target, delta in total.items():
        w = unet_sd[target]
        out[target] = (w.float() + delta).to(w.dtype)
    return out, reports


def merge_lora_into_state_dict(unet_sd: Dict[str, torch.Tensor], lora_sd: Dict[str, torch.Tensor], multiplier: float = 1.0):
    out, reports = merge_loras(unet_sd, [(lora_sd, multiplier)])
    return out, reports[0]
