"""Style LoRAs (reference `backends/cuda_worker.py:123-196`, `backends/styles.py`): file-format
parsing (kohya / peft / old diffusers), dW = (alpha/r) up @ down, and the linearity the merge
relies on — pack(W + w dW) == pack(W) + w pack(dW) for every packed tensor the kernels read."""
import torch

from backends.styles import StyleRequest, StyleDef, clamp_level, parse_style_request
from dreamlab_b200 import synthetic as S
from dreamlab_b200.lora import StyleManager, _leaves, lora_weight_deltas, pack_unet_f32
from oracle.unet import UNetConfig


def _tiny():
    cfg = UNetConfig.tiny()
    shapes = S.unet_shapes(cfg)
    return cfg, shapes, S.random_state_dict(shapes, 0)


def _kohya_lora(shapes, modules, r=4, alpha=2.0, seed=3):
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for mod in modules:
        shp = shapes[mod + ".weight"]
        name = "lora_unet_" + mod.replace(".", "_")
        if len(shp) == 4:
            sd[name + ".lora_down.weight"] = torch.randn(r, shp[1], shp[2], shp[3], generator=g) * 0.1
            sd[name + ".lora_up.weight"] = torch.randn(shp[0], r, 1, 1, generator=g) * 0.1
        else:
            sd[name + ".lora_down.weight"] = torch.randn(r, shp[1], generator=g) * 0.1
            sd[name + ".lora_up.weight"] = torch.randn(shp[0], r, generator=g) * 0.1
        sd[name + ".alpha"] = torch.tensor(alpha)
    sd["lora_te_text_model_encoder_layers_0_self_attn_q_proj.lora_down.weight"] = torch.zeros(r, 8)
    return sd


MODS = ["down_blocks.0.attentions.0.transformer_blocks.0.attn1.to_q",
        "down_blocks.0.attentions.0.transformer_blocks.0.attn1.to_v",
        "down_blocks.1.attentions.1.transformer_blocks.0.attn2.to_k",
        "mid_block.attentions.0.transformer_blocks.0.attn2.to_out.0",
        "up_blocks.1.attentions.2.transformer_blocks.0.ff.net.0.proj",
        "up_blocks.2.attentions.0.transformer_blocks.0.ff.net.2",
        "down_blocks.0.attentions.1.proj_in",
        "down_blocks.1.resnets.0.conv1", "up_blocks.0.upsamplers.0.conv",
        "mid_block.resnets.0.time_emb_proj"]


def test_kohya_deltas_and_alpha_scale():
    _, shapes, _ = _tiny()
    lora = _kohya_lora(shapes, MODS, r=4, alpha=2.0)
    deltas, skipped = lora_weight_deltas(shapes, lora)
    assert set(deltas) == {m + ".weight" for m in MODS}
    assert any(k.startswith("lora_te") for k in skipped)
    m = MODS[0]
    n = "lora_unet_" + m.replace(".", "_")
    want = (2.0 / 4) * lora[n + ".lora_up.weight"] @ lora[n + ".lora_down.weight"]
    assert torch.allclose(deltas[m + ".weight"], want, atol=1e-6)
    c = "down_blocks.1.resnets.0.conv1"
    n = "lora_unet_" + c.replace(".", "_")
    want = 0.5 * torch.einsum("or,rikl->oikl", lora[n + ".lora_up.weight"][:, :, 0, 0], lora[n + ".lora_down.weight"])
    assert torch.allclose(deltas[c + ".weight"], want, atol=1e-6)


def test_peft_and_old_diffusers_formats():
    _, shapes, _ = _tiny()
    m = MODS[0]
    o, i = shapes[m + ".weight"]
    A, B = torch.randn(4, i), torch.randn(o, 4)
    for down, up in ((f"unet.{m}.lora_A.weight", f"unet.{m}.lora_B.weight"),
                     (f"unet.{m}.lora.down.weight", f"unet.{m}.lora.up.weight"),
                     (f"{m}.lora_A.default.weight", f"{m}.lora_B.default.weight")):
        d, _ = lora_weight_deltas(shapes, {down: A, up: B})
        assert torch.allclose(d[m + ".weight"], B @ A, atol=1e-5), down
    attn = m.rsplit(".", 1)[0]
    d, _ = lora_weight_deltas(shapes, {f"unet.{attn}.processor.to_q_lora.down.weight": A,
                                       f"unet.{attn}.processor.to_q_lora.up.weight": B})
    assert torch.allclose(d[m + ".weight"], B @ A, atol=1e-5)


def test_packing_is_linear_in_the_weights():
    cfg, shapes, sd = _tiny()
    deltas, _ = lora_weight_deltas(shapes, _kohya_lora(shapes, MODS))
    w = 0.85
    merged = {k: (v + w * deltas[k] if k in deltas else v) for k, v in sd.items()}
    dsd = {k: (deltas[k] if k in deltas else torch.zeros(shp)) for k, shp in shapes.items()}
    P0, P1, PD = (pack_unet_f32(x, cfg, "cpu") for x in (sd, merged, dsd))
    a, b = [], []
    _leaves(P0, P1, (), a)
    _leaves(P0, PD, (), b)
    touched = 0
    for (path, p0, p1), (_, _, pd) in zip(a, b):
        assert torch.allclose(p1, p0 + w * pd, atol=1e-5), path
        touched += int(float(pd.abs().max()) > 0)
    assert touched == len(MODS) - 1      # every adapted module lands in a packed tensor (to_q, to_v share qkv_w)


def test_style_manager_switch_and_restore_cpu():
    """set_adapter / disable on CPU tensors standing in for the packed device weights."""
    cfg, shapes, sd = _tiny()

    class Eng:
        device = torch.device("cpu")
    Eng.cfg = cfg
    from dreamlab_b200.weights import pack_unet
    Eng.P = pack_unet(sd, cfg, "cpu")                 # bf16 GEMM operands, as on the device
    before = Eng.P["down"][0]["attns"][0]["blocks"][0]["qkv_w"].clone()
    mgr = StyleManager(Eng, shapes)
    ad = mgr.load("style_x", _kohya_lora(shapes, MODS))
    assert ad.num_tensors >= len(MODS) - 1
    mgr.set_adapter("style_x", 1.0)
    after = Eng.P["down"][0]["attns"][0]["blocks"][0]["qkv_w"]
    assert not torch.equal(after, before)
    # the ones column of V (a constructed bias) is never touched
    assert torch.equal(Eng.P["down"][0]["attns"][0]["blocks"][0]["qkv_b"],
                       pack_unet(sd, cfg, "cpu")["down"][0]["attns"][0]["blocks"][0]["qkv_b"])
    mgr.set_adapter("style_x", 0.5)
    mgr.disable()
    assert torch.equal(Eng.P["down"][0]["attns"][0]["blocks"][0]["qkv_w"], before)   # bit-exact restore


def test_style_request_contract():
    reg = {"papercut": StyleDef("papercut", "Papercut", "/x.safetensors", "style_papercut", [0.8, 0.9, 1.0, 1.15])}
    assert parse_style_request({"style_lora": {"style": "papercut", "level": 2}}) == StyleRequest("papercut", 2)
    assert parse_style_request({"style": "none"}).style_id is None
    assert StyleRequest("papercut", 9).weight(reg) == 1.15 and StyleRequest("papercut", 1).weight(reg) == 0.8
    assert StyleRequest("papercut", 0).weight(reg) is None and StyleRequest("nope", 2).weight(reg) is None
    assert clamp_level(0, 4) == 1 and clamp_level(7, 4) == 4
