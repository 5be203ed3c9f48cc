"""The plugin seams as the webui would drive them, against a stub `modules` package (structural twins of
modules/script_callbacks.py:467,594-606, modules/sd_unet.py:63-83, modules/sd_hijack_optimizations.py:25-48,
modules/devices.py, modules/shared.py): the extension script imports, registers its three callbacks, offers the UNet
option only for checkpoints the engine implements, and subclasses the webui's own base classes."""
import importlib
import importlib.util
import os
import sys
import types

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXT = os.path.join(ROOT, "stable-diffusion-webui_b200", "webui_extension", "scripts", "sdxe_unet.py")


class _Callbacks(types.ModuleType):
    def __init__(self):
        super().__init__("modules.script_callbacks")
        self.unets, self.optimizers, self.model_loaded = [], [], []

    def on_list_unets(self, cb, *, name=None):
        self.unets.append(cb)

    def on_list_optimizers(self, cb, *, name=None):
        self.optimizers.append(cb)

    def on_model_loaded(self, cb, *, name=None):
        self.model_loaded.append(cb)


def _sd15_like_unet_sd(in_ch=4, ctx=768):
    return {"input_blocks.0.0.weight": torch.zeros(320, in_ch, 3, 3),
            "input_blocks.1.1.transformer_blocks.0.attn2.to_k.weight": torch.zeros(320, ctx)}


def _sd_model(unet_sd, **attrs):
    dm = types.SimpleNamespace(state_dict=lambda: unet_sd)
    m = types.SimpleNamespace(model=types.SimpleNamespace(diffusion_model=dm, conditioning_key="crossattn"),
                              sd_checkpoint_info=types.SimpleNamespace(model_name="ckpt"), parameterization="eps",
                              is_sd2=False, is_sdxl=False, is_sdxl_inpaint=False, first_stage_model=None)
    for k, v in attrs.items():
        setattr(m, k, v)
    return m


@pytest.fixture()
def webui(monkeypatch):
    """Installs the stub `modules` package, reloads the seam modules so that they bind to it, yields (callbacks, shared,
    extension module), then restores the headless bindings."""
    cb = _Callbacks()

    class SdUnetOption:
        model_name = None
        label = None

        def create_unet(self):
            raise NotImplementedError()

    class SdUnet(torch.nn.Module):
        def forward(self, x, timesteps, context, *args, **kwargs):
            raise NotImplementedError()

        def activate(self):
            pass

        def deactivate(self):
            pass

    class SdOptimization:
        name = None
        label = None
        cmd_opt = None
        priority = 0

        def title(self):
            return self.name if self.label is None else f"{self.name} - {self.label}"

        def is_available(self):
            return True

        def apply(self):
            pass

        def undo(self):
            pass

    shared = types.ModuleType("modules.shared")
    shared.device = torch.device("cpu")
    shared.sd_model = _sd_model(_sd15_like_unet_sd())
    shared.loaded_hypernetworks = []
    shared.opts = types.SimpleNamespace(upcast_attn=False)
    devices = types.ModuleType("modules.devices")
    devices.dtype_unet, devices.dtype_vae = torch.float16, torch.float32
    sd_unet = types.ModuleType("modules.sd_unet")
    sd_unet.SdUnetOption, sd_unet.SdUnet = SdUnetOption, SdUnet
    sd_opt = types.ModuleType("modules.sd_hijack_optimizations")
    sd_opt.SdOptimization = SdOptimization
    hyper = types.ModuleType("modules.hypernetworks.hypernetwork")
    hyper.calls = []

    def apply_hypernetworks(hns, context, layer=None):
        hyper.calls.append(len(hns))
        return context, context

    hyper.apply_hypernetworks = apply_hypernetworks
    hpkg = types.ModuleType("modules.hypernetworks")
    hpkg.hypernetwork = hyper
    pkg = types.ModuleType("modules")
    pkg.script_callbacks, pkg.shared, pkg.devices, pkg.sd_unet, pkg.sd_hijack_optimizations, pkg.hypernetworks = cb, shared, devices, sd_unet, sd_opt, hpkg
    stubs = {"modules": pkg, "modules.script_callbacks": cb, "modules.shared": shared, "modules.devices": devices, "modules.sd_unet": sd_unet,
             "modules.sd_hijack_optimizations": sd_opt, "modules.hypernetworks": hpkg, "modules.hypernetworks.hypernetwork": hyper}
    for k, v in stubs.items():
        monkeypatch.setitem(sys.modules, k, v)
    import sdwebui_b200  # noqa: F401
    from sdwebui_b200 import sd_hijack_optimizations as prod_opt
    from sdwebui_b200 import sd_unet as prod_unet

    importlib.reload(prod_unet)   # rebind the base classes to the stub webui's
    importlib.reload(prod_opt)
    spec = importlib.util.spec_from_file_location("sdxe_ext_under_test", EXT)
    ext = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ext)
    yield cb, shared, ext, (SdUnetOption, SdUnet, SdOptimization), hyper
    for k in stubs:
        monkeypatch.delitem(sys.modules, k, raising=False)
    importlib.reload(prod_unet)
    importlib.reload(prod_opt)


def test_extension_registers_and_subclasses(webui):
    cb, shared, ext, (SdUnetOption, SdUnet, SdOptimization), _ = webui
    assert len(cb.unets) == 1 and len(cb.optimizers) == 1 and len(cb.model_loaded) == 1
    unets, opts = [], []
    cb.unets[0](unets)
    cb.optimizers[0](opts)
    assert len(unets) == 1 and isinstance(unets[0], SdUnetOption)
    assert unets[0].model_name == "ckpt" and unets[0].label == "[sdxe] ckpt"      # "Automatic" matches on model_name
    assert len(opts) == 1 and isinstance(opts[0], SdOptimization) and opts[0].name == "sdxe"
    assert opts[0].priority < 70                                                     # never beats the stock CUDA choices by itself
    from sdwebui_b200.sd_unet import SdxeUnet

    assert issubclass(SdxeUnet, SdUnet)


@pytest.mark.parametrize("why,model", [
    ("v-prediction", dict(parameterization="v")),
    ("sd2", dict(is_sd2=True)),
    ("inpainting conditioning", dict(model=types.SimpleNamespace(diffusion_model=types.SimpleNamespace(state_dict=lambda: _sd15_like_unet_sd()), conditioning_key="hybrid"))),
    ("9-channel input", dict(model=types.SimpleNamespace(diffusion_model=types.SimpleNamespace(state_dict=lambda: _sd15_like_unet_sd(in_ch=9)), conditioning_key="crossattn"))),
    ("SD2 context width", dict(model=types.SimpleNamespace(diffusion_model=types.SimpleNamespace(state_dict=lambda: _sd15_like_unet_sd(ctx=1024)), conditioning_key="crossattn"))),
    ("sdxl inpaint", dict(is_sdxl_inpaint=True)),
])
def test_extension_skips_unsupported_checkpoints(webui, why, model):
    cb, shared, ext, _, _ = webui
    shared.sd_model = _sd_model(_sd15_like_unet_sd(), **model)
    unets = []
    cb.unets[0](unets)
    assert unets == [], why


def test_lora_provider_is_wired(webui, monkeypatch, tmp_path):
    """create_unet() hands the networks the built-in Lora extension has loaded (by file) to SdxeUnet(loras=...)."""
    cb, shared, ext, _, _ = webui
    from sdwebui_b200.sd_models import save_safetensors

    f = str(tmp_path / "l.safetensors")
    save_safetensors({"lora_unet_x.alpha": torch.tensor(4.0)}, f)
    nets = types.ModuleType("networks")
    nets.loaded_networks = [types.SimpleNamespace(network_on_disk=types.SimpleNamespace(filename=f), unet_multiplier=0.6)]
    monkeypatch.setitem(sys.modules, "networks", nets)
    unets = []
    cb.unets[0](unets)
    u = unets[0].create_unet()
    assert len(u.loras) == 1 and u.loras[0][1] == 0.6 and "lora_unet_x.alpha" in u.loras[0][0]


def test_vae_wrapper_leaves_fp32_vae_alone_and_restores(webui):
    cb, shared, ext, _, _ = webui
    calls = []
    fs = types.SimpleNamespace(state_dict=lambda: {}, decode=lambda z: calls.append("decode"), encode=lambda x: calls.append("encode"))
    m = _sd_model(_sd15_like_unet_sd(), first_stage_model=fs)
    orig = fs.decode
    cb.model_loaded[0](m)           # devices.dtype_vae is fp32 in the stub: nothing is wrapped
    assert fs.decode is orig
    sys.modules["modules.devices"].dtype_vae = torch.float16
    cb.model_loaded[0](m)           # unknown (empty) VAE layout: the attempt fails, the stock methods stay
    assert fs.decode is orig and not hasattr(fs, "_sdxe_orig_decode")


def test_attention_seam_calls_hypernetworks_and_rejects_upcast(webui):
    cb, shared, ext, _, hyper = webui
    from sdwebui_b200 import sd_hijack_optimizations as so
    from sdwebui_b200.lib import SdxeError

    hns, apply_hn, upcast = so._webui_state()
    assert apply_hn is hyper.apply_hypernetworks and upcast is False
    shared.opts.upcast_attn = True
    with pytest.raises(SdxeError):
        so._check_upcast(so._webui_state()[2])


def test_vaespec_from_state_dict_roundtrip():
    from sdwebui_b200 import checkpoint as C
    from sdwebui_b200.engine import VAESpec
    from sdwebui_b200.lib import SdxeError

    for spec in (VAESpec(), VAESpec(ch=64, ch_mult=[1, 2], num_res_blocks=1)):
        d = C.empty_state_dict(C.vae_decoder_param_shapes(spec), "cpu")
        e = C.empty_state_dict(C.vae_encoder_param_shapes(spec), "cpu")
        assert VAESpec.from_state_dict(d) == spec
        assert VAESpec.from_state_dict(e) == spec
    with pytest.raises(SdxeError):
        VAESpec.from_state_dict({"foo": torch.zeros(1)})
