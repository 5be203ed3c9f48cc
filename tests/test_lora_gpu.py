"""SURVEY §8(f) N3 on the GPU: a LoRA handed to SdxeUnet is merged before the engine packs its weights — the engine then
computes what the stock UNet with `weight += updown` (the reference's merged state) computes."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


def rel_err(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


def test_lora_merge_reaches_the_packed_weights(cuda):
    from oracle.synth import init_module_
    from oracle.unet import UNetModel, tiny_config
    from sdwebui_b200 import extra_networks_lora as X
    from sdwebui_b200.engine import UNetSpec
    from sdwebui_b200.sd_unet import SdxeUnet

    dtype = torch.float16
    cfg = tiny_config()
    unet = init_module_(UNetModel(cfg), 61).eval().to(cuda)
    sd = {k: v.to(dtype) for k, v in unet.state_dict().items()}
    g = torch.Generator(device="cuda").manual_seed(3)
    lora = {}

    def add(name, key, rank, alpha=None, conv3=False):
        w = sd[key]
        if w.ndim == 4 and conv3:
            down = torch.randn(rank, w.shape[1], 3, 3, device=cuda, generator=g) * 0.2
            up = torch.randn(w.shape[0], rank, 1, 1, device=cuda, generator=g) * 0.2
        elif w.ndim == 4:
            down = torch.randn(rank, w.shape[1], 1, 1, device=cuda, generator=g) * 0.2
            up = torch.randn(w.shape[0], rank, 1, 1, device=cuda, generator=g) * 0.2
        else:
            down = torch.randn(rank, w.shape[1], device=cuda, generator=g) * 0.2
            up = torch.randn(w.shape[0], rank, device=cuda, generator=g) * 0.2
        lora[name + ".lora_up.weight"], lora[name + ".lora_down.weight"] = up.half(), down.half()
        if alpha is not None:
            lora[name + ".alpha"] = torch.tensor(float(alpha))

    # every packed-layout family: stacked q|k|v, cross k|v (batched context projection), LayerNorm-folded q / ff1 (GEGLU
    # interleave), out-projection, proj_in conv1x1, resblock conv3x3, time-embedding linear
    p = "input_blocks.1.1.transformer_blocks.0."
    add("lora_unet_down_blocks_0_attentions_0_transformer_blocks_0_attn1_to_q", p + "attn1.to_q.weight", 4, 2)
    add("lora_unet_down_blocks_0_attentions_0_transformer_blocks_0_attn1_to_v", p + "attn1.to_v.weight", 4)
    add("lora_unet_down_blocks_0_attentions_0_transformer_blocks_0_attn2_to_q", p + "attn2.to_q.weight", 8, 8)
    add("lora_unet_down_blocks_0_attentions_0_transformer_blocks_0_attn2_to_k", p + "attn2.to_k.weight", 4, 1)
    add("lora_unet_down_blocks_0_attentions_0_transformer_blocks_0_ff_net_0_proj", p + "ff.net.0.proj.weight", 4, 4)
    add("lora_unet_down_blocks_0_attentions_0_transformer_blocks_0_attn1_to_out_0", p + "attn1.to_out.0.weight", 4, 4)
    add("lora_unet_down_blocks_0_attentions_0_proj_in", "input_blocks.1.1.proj_in.weight", 4, 4)
    add("lora_unet_mid_block_resnets_0_conv1", "middle_block.0.in_layers.2.weight", 4, 4, conv3=True)
    add("lora_unet_time_embedding_linear_1", "time_embed.0.weight", 4, 4)
    lora["lora_te_text_model_encoder_layers_0_mlp_fc1.lora_up.weight"] = torch.zeros(8, 4)

    mult = 0.8
    merged, rep = X.merge_lora_into_state_dict(sd, lora, mult)
    assert len(rep["merged"]) == 9 and not rep["unmatched"] and len(rep["skipped_text_encoder"]) == 1
    ref = copy.deepcopy(unet)
    ref.load_state_dict({k: v.float() for k, v in merged.items()})

    su = SdxeUnet(sd, spec=UNetSpec.from_any(cfg), dtype=dtype, device=str(cuda), loras=[(lora, mult)])
    su.activate()
    base = SdxeUnet(sd, spec=UNetSpec.from_any(cfg), dtype=dtype, device=str(cuda))
    base.activate()
    x = torch.randn(2, 4, 32, 32, device=cuda, generator=g).to(dtype)
    t = torch.tensor([700.0, 31.0], device=cuda).to(dtype)
    ctx = torch.randn(2, 77, cfg.context_dim, device=cuda, generator=g).to(dtype)
    out = su.forward(x, t, ctx)
    out0 = base.forward(x, t, ctx)
    with torch.no_grad():
        want = ref(x.float(), t.float(), context=ctx.float())
        plain = unet(x.float(), t.float(), context=ctx.float())
    e = rel_err(out, want)
    moved = rel_err(want, plain)
    print(f"lora: engine vs merged oracle {e:.3e}; the LoRA moves the output by {moved:.3e}; unmerged engine vs merged oracle {rel_err(out0, want):.3e}")
    assert moved > 5e-2                      # the network really changes the function
    assert e < 3e-3 and rel_err(out0, want) > 10 * e
    assert len(su.lora_reports) == 1 and len(su.lora_reports[0]["merged"]) == 9
