"""SURVEY §8(f) N3 host logic: LoRA key mapping (networks.py:56-119) and the merged-weight arithmetic
(network_lora.py:66-84, network.py:167-214), on CPU."""
import pytest
import torch

import sdwebui_b200  # noqa: F401
from sdwebui_b200 import checkpoint as C
from sdwebui_b200 import extra_networks_lora as X
from sdwebui_b200 import lib as L
from sdwebui_b200.engine import UNetSpec


def test_name_conversion_known_cases():
    cv = X.convert_diffusers_name_to_compvis
    assert cv("lora_unet_down_blocks_0_attentions_0_transformer_blocks_0_attn1_to_q") == \
        "diffusion_model_input_blocks_1_1_transformer_blocks_0_attn1_to_q"
    assert cv("lora_unet_down_blocks_2_attentions_1_proj_in") == "diffusion_model_input_blocks_8_1_proj_in"
    assert cv("lora_unet_down_blocks_1_resnets_0_conv1") == "diffusion_model_input_blocks_4_0_in_layers_2"
    assert cv("lora_unet_down_blocks_1_resnets_1_time_emb_proj") == "diffusion_model_input_blocks_5_0_emb_layers_1"
    assert cv("lora_unet_down_blocks_0_downsamplers_0_conv") == "diffusion_model_input_blocks_3_0_op"
    assert cv("lora_unet_mid_block_attentions_0_transformer_blocks_0_ff_net_2") == \
        "diffusion_model_middle_block_1_transformer_blocks_0_ff_net_2"
    assert cv("lora_unet_mid_block_resnets_1_conv2") == "diffusion_model_middle_block_2_out_layers_3"
    assert cv("lora_unet_up_blocks_1_attentions_2_transformer_blocks_0_attn2_to_v") == \
        "diffusion_model_output_blocks_5_1_transformer_blocks_0_attn2_to_v"
    assert cv("lora_unet_up_blocks_0_upsamplers_0_conv") == "diffusion_model_output_blocks_2_1_conv"
    assert cv("lora_unet_up_blocks_2_upsamplers_0_conv") == "diffusion_model_output_blocks_8_2_conv"
    assert cv("lora_unet_up_blocks_3_resnets_0_conv_shortcut") == "diffusion_model_output_blocks_9_0_skip_connection"
    assert cv("lora_unet_time_embedding_linear_2") == "diffusion_model_time_embed_2"
    assert cv("lora_unet_conv_in") == "diffusion_model_input_blocks_0_0"
    assert cv("lora_te_text_model_encoder_layers_0_mlp_fc1") == "lora_te_text_model_encoder_layers_0_mlp_fc1"


def test_every_diffusers_style_name_lands_on_a_real_sd15_weight():
    """Build the diffusers-style name of every attention / resnet / sampler module of SD1.5 and check that the converted
    name is one of the engine's weights (i.e. the mapping is onto the real ldm layout, not just syntactically right)."""
    shapes = C.unet_param_shapes(UNetSpec.sd15())
    mapping = X.network_layer_mapping(shapes.keys())
    names = ["lora_unet_conv_in", "lora_unet_conv_out", "lora_unet_time_embedding_linear_1", "lora_unet_time_embedding_linear_2"]
    att = ["proj_in", "proj_out"] + [f"transformer_blocks_0_{x}" for x in
                                      ("attn1_to_q", "attn1_to_k", "attn1_to_v", "attn1_to_out_0", "attn2_to_q", "attn2_to_k",
                                       "attn2_to_v", "attn2_to_out_0", "ff_net_0_proj", "ff_net_2")]
    for lvl in range(4):
        for j in range(2):
            names += [f"lora_unet_down_blocks_{lvl}_resnets_{j}_{s}" for s in ("conv1", "conv2", "time_emb_proj")]
            if lvl < 3:
                names += [f"lora_unet_down_blocks_{lvl}_attentions_{j}_{s}" for s in att]
        if lvl < 3:
            names.append(f"lora_unet_down_blocks_{lvl}_downsamplers_0_conv")
        for j in range(3):
            names += [f"lora_unet_up_blocks_{lvl}_resnets_{j}_{s}" for s in ("conv1", "conv2", "time_emb_proj", "conv_shortcut")]
            if lvl > 0:
                names += [f"lora_unet_up_blocks_{lvl}_attentions_{j}_{s}" for s in att]
        if lvl < 3:
            names.append(f"lora_unet_up_blocks_{lvl}_upsamplers_0_conv")
    names += [f"lora_unet_mid_block_resnets_{j}_{s}" for j in range(2) for s in ("conv1", "conv2", "time_emb_proj")]
    names += [f"lora_unet_mid_block_attentions_0_{s}" for s in att]
    missing = [n for n in names if X.convert_diffusers_name_to_compvis(n) not in mapping]
    assert not missing, missing[:5]
    assert len(set(X.convert_diffusers_name_to_compvis(n) for n in names)) == len(names)


def _lora_for(shape, rank, g, conv3=False):
    if len(shape) == 4 and conv3:
        down = torch.randn(rank, shape[1], 3, 3, generator=g) * 0.05
        up = torch.randn(shape[0], rank, 1, 1, generator=g) * 0.05
    elif len(shape) == 4:
        down = torch.randn(rank, shape[1], 1, 1, generator=g) * 0.05
        up = torch.randn(shape[0], rank, 1, 1, generator=g) * 0.05
    else:
        down = torch.randn(rank, shape[1], generator=g) * 0.05
        up = torch.randn(shape[0], rank, generator=g) * 0.05
    return up, down


def test_merge_arithmetic_and_reports():
    g = torch.Generator().manual_seed(1)
    sd = {"input_blocks.1.1.transformer_blocks.0.attn1.to_q.weight": torch.randn(64, 64, generator=g).half(),
          "input_blocks.1.1.proj_in.weight": torch.randn(64, 64, 1, 1, generator=g).half(),
          "input_blocks.1.0.in_layers.2.weight": torch.randn(64, 32, 3, 3, generator=g).half(),
          "input_blocks.1.0.in_layers.2.bias": torch.zeros(64).half(),
          "out.2.weight": torch.randn(4, 64, 3, 3, generator=g).half()}
    lo = {}
    up, down = _lora_for((64, 64), 4, g)
    lo["lora_unet_down_blocks_0_attentions_0_transformer_blocks_0_attn1_to_q.lora_up.weight"] = up
    lo["lora_unet_down_blocks_0_attentions_0_transformer_blocks_0_attn1_to_q.lora_down.weight"] = down
    lo["lora_unet_down_blocks_0_attentions_0_transformer_blocks_0_attn1_to_q.alpha"] = torch.tensor(2.0)
    up1, down1 = _lora_for((64, 64, 1, 1), 8, g)
    lo["lora_unet_down_blocks_0_attentions_0_proj_in.lora_B.weight"] = up1      # A/B naming, no alpha -> scale 1
    lo["lora_unet_down_blocks_0_attentions_0_proj_in.lora_A.weight"] = down1
    up3, down3 = _lora_for((64, 32, 3, 3), 4, g, conv3=True)
    lo["lora_unet_down_blocks_0_resnets_0_conv1.lora_up.weight"] = up3
    lo["lora_unet_down_blocks_0_resnets_0_conv1.lora_down.weight"] = down3
    lo["lora_unet_down_blocks_0_resnets_0_conv1.alpha"] = torch.tensor(4.0)
    lo["lora_te_text_model_encoder_layers_0_mlp_fc1.lora_up.weight"] = torch.zeros(8, 4)
    lo["lora_unet_down_blocks_3_attentions_9_nope.lora_up.weight"] = torch.zeros(8, 4)
    out, rep = X.merge_lora_into_state_dict(sd, lo, multiplier=0.7)
    k = "input_blocks.1.1.transformer_blocks.0.attn1.to_q.weight"
    want = (sd[k].float() + 0.7 * (2.0 / 4) * (up @ down)).half()
    assert torch.equal(out[k], want)
    k1 = "input_blocks.1.1.proj_in.weight"
    want1 = (sd[k1].float() + 0.7 * (up1.reshape(64, 8) @ down1.reshape(8, 64)).reshape(64, 64, 1, 1)).half()
    assert torch.equal(out[k1], want1)
    k3 = "input_blocks.1.0.in_layers.2.weight"
    want3 = (sd[k3].float() + 0.7 * (4.0 / 4) * (up3.reshape(64, 4) @ down3.reshape(4, -1)).reshape(64, 32, 3, 3)).half()
    assert torch.equal(out[k3], want3)
    assert out["out.2.weight"] is sd["out.2.weight"] and out[k3 .replace("weight", "bias")] is sd["input_blocks.1.0.in_layers.2.bias"]
    assert len(rep["merged"]) == 3 and len(rep["skipped_text_encoder"]) == 1 and len(rep["unmatched"]) == 1
    # two networks on the same weight: deltas add in fp32, one rounding
    out2, _ = X.merge_loras(sd, [(lo, 0.7), (lo, -0.7)])
    assert torch.equal(out2[k], sd[k])


def test_unsupported_module_types_raise():
    sd = {"out.2.weight": torch.zeros(4, 8, 3, 3)}
    with pytest.raises(L.SdxeError):  # LoHa
        X.merge_lora_into_state_dict(sd, {"lora_unet_conv_out.hada_w1_a": torch.zeros(4, 2), "lora_unet_conv_out.hada_w1_b": torch.zeros(2, 72)})
    with pytest.raises(L.SdxeError):  # DoRA
        X.merge_lora_into_state_dict(sd, {"lora_unet_conv_out.lora_up.weight": torch.zeros(4, 2, 1, 1),
                                          "lora_unet_conv_out.lora_down.weight": torch.zeros(2, 8, 3, 3),
                                          "lora_unet_conv_out.dora_scale": torch.zeros(1, 8, 1, 1)})
    with pytest.raises(L.SdxeError):  # wrong geometry
        X.merge_lora_into_state_dict(sd, {"lora_unet_conv_out.lora_up.weight": torch.zeros(5, 2, 1, 1),
                                          "lora_unet_conv_out.lora_down.weight": torch.zeros(2, 8, 3, 3)})
