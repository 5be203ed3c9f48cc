"""webui extension entry point (drop this directory into <webui>/extensions/sdxe/): registers the sdxe UNet and the
sdxe attention optimisation through the reference's own plugin callbacks:
  script_callbacks.on_list_unets       modules/script_callbacks.py:602-606  -> modules/sd_unet.py:10-14
  script_callbacks.on_list_optimizers  modules/script_callbacks.py:594-599  -> modules/sd_hijack.py:48-56
  script_callbacks.on_model_loaded     modules/script_callbacks.py:467      -> wraps first_stage_model.decode (B3) and
                                                                              first_stage_model.encode (img2img init, N1)
Nothing in the webui itself is modified.
"""
import os
import sys

from modules import script_callbacks, shared  # type: ignore

_ROOT = os.environ.get("SDXE_ROOT", os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

import sdwebui_b200  # noqa: E402,F401
import torch  # noqa: E402

from sdwebui_b200.engine import VAEDecoderEngine, VAEEncoderEngine, VAESpec  # noqa: E402
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
    esd = {k: v for k, v in fs.state_dict().items() if k.startswith(("encoder.", "quant_conv."))}
    if esd:
        enc = VAEEncoderEngine(VAESpec(), device=shared.device)
        enc.load_state_dict(esd)
        enc.finalize()
        fs.encode = lambda x, *a, **k: _Posterior(enc.encode_moments(x))


class _Posterior:
    """What ldm's AutoencoderKL.encode returns (DiagonalGaussianDistribution), as far as get_first_stage_encoding and
    images_tensor_to_samples (modules/sd_samplers_common.py:87-112) use it: sample() / mode()."""

    def __init__(self, moments):
        self.mean, self.logvar = torch.chunk(moments.float(), 2, dim=1)
        self.logvar = torch.clamp(self.logvar, -30.0, 20.0)
        self.std = torch.exp(0.5 * self.logvar)

    def sample(self):
        return self.mean + self.std * torch.randn(self.mean.shape, device=self.mean.device)

    def mode(self):
        return self.mean


script_callbacks.on_list_unets(_list_unets)
script_callbacks.on_list_optimizers(_list_optimizers)
script_callbacks.on_model_loaded(_model_loaded)
