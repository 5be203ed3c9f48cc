"""Checkpoint ingestion for the sdxe engines: the on-disk formats the webui loads
(`modules/sd_models.py:243-329 read_state_dict / get_state_dict_from_checkpoint`) and its architecture guess from
the checkpoint's own keys (`modules/sd_models_config.py:72-114 guess_model_config_from_state_dict`), restricted to what
the engines implement: SD1.x and SDXL-base UNets (eps-prediction, 4 input channels) plus the KL-VAE decoder.

Host-side only: tensors are read into torch tensors and handed to `sdxe_set_weight`, which repacks them on the GPU
into the kernel layouts (K-major 16-bit GEMM operands, tap-major conv weights, GEGLU interleave, LayerNorm folds).
fp8-stored checkpoints (`fp8_storage`, modules/sd_models.py:410-520: weights kept as float8 and upcast by autocast) are
upcast once at load: the engine's resident format is 16 bit.
"""
from __future__ import annotations

import json
import os
import struct
from dataclasses import dataclass, field
from typing import Dict, Optional

import torch

from . import lib as L
from .engine import UNetSpec, VAEDecoderEngine, VAESpec

UNET_PREFIX = "model.diffusion_model."
VAE_PREFIX = "first_stage_model."

# modules/sd_models.py:243-247 (only matters for the text encoder's keys, kept so the dict looks the same to callers)
_CHECKPOINT_DICT_REPLACEMENTS_SD1 = {
    "cond_stage_model.transformer.embeddings.": "cond_stage_model.transformer.text_model.embeddings.",
    "cond_stage_model.transformer.encoder.": "cond_stage_model.transformer.text_model.encoder.",
    "cond_stage_model.transformer.final_layer_norm.": "cond_stage_model.transformer.text_model.final_layer_norm.",
}

_ST_DTYPES = {
    "F16": torch.float16, "BF16": torch.bfloat16, "F32": torch.float32, "F64": torch.float64,
    "I64": torch.int64, "I32": torch.int32, "I16": torch.int16, "I8": torch.int8, "U8": torch.uint8, "BOOL": torch.bool,
}
for _n, _a in (("F8_E4M3", "float8_e4m3fn"), ("F8_E5M2", "float8_e5m2")):
    if hasattr(torch, _a):
        _ST_DTYPES[_n] = getattr(torch, _a)
_ST_NAMES = {v: k for k, v in _ST_DTYPES.items()}


def _read_safetensors(path: str) -> Dict[str, torch.Tensor]:
    """The safetensors container: u64 little-endian header length, JSON header {name: {dtype, shape, data_offsets}},
    raw little-endian tensor bytes. (Same header parse as modules/sd_models.py:278-303.)"""
    with open(path, "rb") as f:
        n = struct.unpack("<Q", f.read(8))[0]
        if n < 2 or n > 100 * 1024 * 1024:
            raise L.SdxeError(f"{path} is not a safetensors file")
        head = f.read(n)
        if head[:1] != b"{":
            raise L.SdxeError(f"{path} is not a safetensors file")
        meta = json.loads(head)
        blob = f.read()
    out = {}
    for name, info in meta.items():
        if name == "__metadata__":
            continue
        dt = _ST_DTYPES.get(info["dtype"])
        if dt is None:
            raise L.SdxeError(f"{path}: unsupported safetensors dtype {info['dtype']} for {name}")
        a, b = info["data_offsets"]
        if b == a:
            out[name] = torch.empty(info["shape"], dtype=dt)
            continue
        t = torch.frombuffer(bytearray(blob[a:b]), dtype=torch.uint8).view(dt)
        out[name] = t.reshape(info["shape"])
    return out


def save_safetensors(sd: Dict[str, torch.Tensor], path: str, metadata: Optional[Dict[str, str]] = None) -> None:
    """Writer for the same container (tests, tools); tensors are stored contiguous in key order."""
    head, chunks, off = {}, [], 0
    for k, v in sd.items():
        v = v.detach().cpu().contiguous()
        if v.dtype not in _ST_NAMES:
            raise L.SdxeError(f"cannot store dtype {v.dtype}")
        raw = v.reshape(-1).view(torch.uint8).numpy().tobytes() if v.numel() else b""
        head[k] = {"dtype": _ST_NAMES[v.dtype], "shape": list(v.shape), "data_offsets": [off, off + len(raw)]}
        chunks.append(raw)
        off += len(raw)
    if metadata:
        head["__metadata__"] = metadata
    hb = json.dumps(head, separators=(",", ":")).encode()
    hb += b" " * ((8 - len(hb) % 8) % 8)
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(hb)))
        f.write(hb)
        for c in chunks:
            f.write(c)


def get_state_dict_from_checkpoint(pl_sd: dict) -> dict:
    """modules/sd_models.py:261-276: unwrap a Lightning checkpoint and apply the SD1 key renames."""
    pl_sd = pl_sd.pop("state_dict", pl_sd)
    pl_sd.pop("state_dict", None)
    sd = {}
    for k, v in pl_sd.items():
        for text, repl in _CHECKPOINT_DICT_REPLACEMENTS_SD1.items():
            if k.startswith(text):
                k = repl + k[len(text):]
        sd[k] = v
    return sd


def read_state_dict(checkpoint_file: str, map_location: Optional[str] = None) -> Dict[str, torch.Tensor]:
    """modules/sd_models.py:312-329. `.safetensors` goes through the safetensors package when it is importable (mmap),
    else through the built-in parser; anything else is a torch pickle (loaded with weights_only=True)."""
    ext = os.path.splitext(checkpoint_file)[1].lower()
    if ext == ".safetensors":
        try:
            import safetensors.torch as st  # type: ignore

            pl_sd = st.load_file(checkpoint_file, device=map_location or "cpu")
        except ImportError:
            pl_sd = _read_safetensors(checkpoint_file)
            if map_location and map_location != "cpu":
                pl_sd = {k: v.to(map_location) for k, v in pl_sd.items()}
    else:
        pl_sd = torch.load(checkpoint_file, map_location=map_location or "cpu", weights_only=True)
    return get_state_dict_from_checkpoint(pl_sd)


@dataclass
class CheckpointInfo:
    kind: str                       # "sd15" | "sdxl"
    unet: UNetSpec
    vae: VAESpec
    has_vae: bool
    storage_dtypes: Dict[str, int] = field(default_factory=dict)
    notes: str = ""


def guess_model_config_from_state_dict(sd: Dict[str, torch.Tensor]) -> CheckpointInfo:
    """The subset of modules/sd_models_config.py:72-114 the engines cover. Every other family the reference recognises
    (SD3, SD2.x, inpainting / instruct-pix2pix input widths, refiner, unCLIP, depth, AltDiffusion) is rejected by name
    rather than mis-loaded."""
    din = sd.get(UNET_PREFIX + "input_blocks.0.0.weight")
    if UNET_PREFIX + "x_embedder.proj.weight" in sd:
        raise L.SdxeError("SD3 checkpoints are out of scope (modules/models/sd3)")
    if din is None:
        raise L.SdxeError("no UNet in this checkpoint (model.diffusion_model.input_blocks.0.0.weight missing)")
    if din.shape[1] != 4:
        raise L.SdxeError(f"UNet takes {din.shape[1]} input channels (inpainting / instruct-pix2pix): not implemented")
    if sd.get("conditioner.embedders.1.model.ln_final.weight") is not None:
        kind, unet = "sdxl", UNetSpec.sdxl()
    elif sd.get("conditioner.embedders.0.model.ln_final.weight") is not None:
        raise L.SdxeError("SDXL refiner checkpoints are not implemented")
    elif sd.get("cond_stage_model.model.transformer.resblocks.0.attn.in_proj_weight") is not None:
        raise L.SdxeError("SD2.x checkpoints (OpenCLIP-H, 1024-wide context) are not implemented")
    else:
        kind, unet = "sd15", UNetSpec.sd15()
    # the spec must agree with the tensors (catches fine-tunes with a different width early, with a readable message)
    k = UNET_PREFIX + "input_blocks.1.1.transformer_blocks.0.attn2.to_k.weight" if kind == "sd15" else \
        UNET_PREFIX + "input_blocks.4.1.transformer_blocks.0.attn2.to_k.weight"
    w = sd.get(k)
    if w is None or w.shape[1] != unet.context_dim or din.shape[0] != unet.model_channels:
        raise L.SdxeError(f"checkpoint does not match the {kind} UNet layout ({k}: {None if w is None else tuple(w.shape)})")
    has_vae = VAE_PREFIX + "decoder.conv_in.weight" in sd
    dts: Dict[str, int] = {}
    for kk, v in sd.items():
        if kk.startswith(UNET_PREFIX):
            dts[str(v.dtype)] = dts.get(str(v.dtype), 0) + 1
    return CheckpointInfo(kind=kind, unet=unet, vae=VAESpec(), has_vae=has_vae, storage_dtypes=dts)


def _compute_ready(t: torch.Tensor) -> torch.Tensor:
    """fp8 / fp64 storage -> a dtype sdxe_set_weight ingests (fp16 for fp8: every e4m3 / e5m2 value is exact in fp16)."""
    if t.dtype in (torch.float16, torch.bfloat16, torch.float32):
        return t
    if "float8" in str(t.dtype):
        return t.to(torch.float16)
    return t.float()


def unet_state_dict(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    return {k[len(UNET_PREFIX):]: _compute_ready(v) for k, v in sd.items() if k.startswith(UNET_PREFIX)}


def vae_state_dict(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    return {k[len(VAE_PREFIX):]: _compute_ready(v) for k, v in sd.items()
            if k.startswith(VAE_PREFIX + "decoder.") or k.startswith(VAE_PREFIX + "post_quant_conv.")}


def vae_encoder_state_dict(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    return {k[len(VAE_PREFIX):]: _compute_ready(v) for k, v in sd.items()
            if k.startswith(VAE_PREFIX + "encoder.") or k.startswith(VAE_PREFIX + "quant_conv.")}


def load_model(checkpoint_file: str, dtype=torch.float16, device="cuda:0", vae_file: Optional[str] = None,
               with_encoder: bool = False):
    """Checkpoint file -> a ready `processing.SdModel` (UNet engine activated, VAE decoder finalized; with_encoder also
    builds the VAE encoder engine img2img needs).
    `vae_file`: external VAE (`--vae-path` / sd_vae.py) whose keys are un-prefixed `decoder.*` / `post_quant_conv.*`."""
    from .processing import SdModel
    from .sd_unet import SdxeUnet

    sd = read_state_dict(checkpoint_file)
    info = guess_model_config_from_state_dict(sd)
    unet = SdxeUnet(unet_state_dict(sd), spec=info.unet, dtype=dtype, device=device)
    unet.activate()
    vsd = None
    if vae_file:
        ext = read_state_dict(vae_file)
        vsd = {k: _compute_ready(v) for k, v in ext.items() if k.startswith(("decoder.", "post_quant_conv."))}
    elif info.has_vae:
        vsd = vae_state_dict(sd)
    vae = None
    if vsd:
        vae = VAEDecoderEngine(info.vae, dtype=dtype, device=torch.device(device))
        vae.load_state_dict(vsd)
        vae.finalize()
    enc = None
    if with_encoder:
        from .engine import VAEEncoderEngine

        esd = ({k: _compute_ready(v) for k, v in ext.items() if k.startswith(("encoder.", "quant_conv."))} if vae_file
               else vae_encoder_state_dict(sd))
        if not esd:
            raise L.SdxeError("with_encoder: the checkpoint has no first_stage_model.encoder.* tensors")
        enc = VAEEncoderEngine(info.vae, dtype=dtype, device=torch.device(device))
        enc.load_state_dict(esd)
        enc.finalize()
    return SdModel(unet, vae, is_sdxl=info.kind == "sdxl", dtype_unet=dtype, device=device, vae_encoder=enc), info
