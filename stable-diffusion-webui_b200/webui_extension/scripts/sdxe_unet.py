"""webui extension entry point (drop this directory into <webui>/extensions/sdxe/): registers the sdxe UNet and the
sdxe attention optimisation through the reference's own plugin callbacks:
  script_callbacks.on_list_unets       modules/script_callbacks.py:602-606  -> modules/sd_unet.py:10-14
  script_callbacks.on_list_optimizers  modules/script_callbacks.py:594-599  -> modules/sd_hijack.py:48-56
  script_callbacks.on_model_loaded     modules/script_callbacks.py:467      -> wraps first_stage_model.decode (B3)
Nothing in the webui itself is modified.
"""
import os
import sys

from modules import script_callbacks, shared  # type: ignore

_ROOT = os.environ.get("SDXE_ROOT", os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

import sdwebui_b200  # noqa: E402,F401
from sdwebui_b200.engine import VAEDecoderEngine, VAESpec  # noqa: E402
from sdwebui_b200.sd_hijack_optimizations import SdOptimizationSdxe  # noqa: E402
from sdwebui_b200.sd_unet import SdxeUnetOption  # noqa: E402


def _list_unets(unets):
    info = getattr(shared.sd_model, "sd_checkpoint_info", None)
    if info is None:
        return
    unets.append(SdxeUnetOption(info.model_name, lambda: shared.sd_model.model.diffusion_model.state_dict()))


def _list_optimizers(optimizers):
    optimizers.append(SdOptimizationSdxe())


def _model_loaded(sd_model):
    fs = sd_model.first_stage_model
    sd = {k: v for k, v in fs.state_dict().items() if k.startswith(("decoder.", "post_quant_conv."))}
    eng = VAEDecoderEngine(VAESpec(), device=shared.device)
    eng.load_state_dict(sd)
    eng.finalize()
    fs.decode = lambda z, *a, **k: eng.decode(z)  # precedent: modules/lowvram.py:64-74,136-137


script_callbacks.on_list_unets(_list_unets)
script_callbacks.on_list_optimizers(_list_optimizers)
script_callbacks.on_model_loaded(_model_loaded)
