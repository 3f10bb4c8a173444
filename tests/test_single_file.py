"""Single-file checkpoints (original CompVis / SGM key layout) -> diffusers names (`single_file.py`), the
conversion diffusers performs inside `from_single_file` (reference `backends/cuda_worker.py:79-85`, `:380`).
CPU: the rename tables are bijections onto exactly the parameter set (names AND shapes) of the SD1.5 / SDXL UNets
and the VAE decoder; the OpenCLIP -> transformers text-tower conversion reproduces a transformers model's state
dict tensor for tensor.  GPU: a worker loaded from a single file renders byte-identical PNGs to the worker loaded
from the equivalent diffusers directory."""
import os
from types import SimpleNamespace

import pytest
import torch

from dreamlab_b200 import single_file as sf
from dreamlab_b200 import synthetic as syn


def _invert(key: str, modules: dict, inner: dict, prefix: str) -> str:
    """diffusers name -> original-layout name (test helper: the inverse of single_file._rename)."""
    inv_mod = {v: k for k, v in modules.items()}
    best = max((m for m in inv_mod if key == m or key.startswith(m + ".")), key=len)
    rest = key[len(best):].lstrip(".")
    for a, b in inner.items():                       # a: original, b: diffusers
        if rest == b or rest.startswith(b + "."):
            rest = a + rest[len(b):]
            break
    return prefix + inv_mod[best] + ("." + rest if rest else "")


def _meta(shapes):
    return {k: torch.empty(s, device="meta") for k, s in shapes.items()}


def _ns(d):
    down = d["down_block_types"]
    return SimpleNamespace(**{**d, "down_attn": tuple("CrossAttn" in t for t in down),
                              "block_out_channels": tuple(d["block_out_channels"]),
                              "attention_head_dim": (tuple(d["attention_head_dim"])
                                                     if isinstance(d["attention_head_dim"], list) else d["attention_head_dim"]),
                              "transformer_layers_per_block": tuple(d.get("transformer_layers_per_block", ())),
                              "use_linear_projection": d.get("use_linear_projection", False),
                              "addition_embed_type": d.get("addition_embed_type")})


@pytest.mark.parametrize("name", ["sd15", "sd15_lcm", "sdxl"])
def test_unet_rename_is_a_bijection_onto_the_diffusers_parameter_set(name):
    cfg = dict(sf.SDXL_UNET if name == "sdxl" else sf.SD15_UNET)
    if name == "sd15_lcm":
        cfg["time_cond_proj_dim"] = 256
    shapes = syn.unet_shapes(_ns(cfg))
    modules = sf.unet_module_map(cfg)
    ldm = {}
    for k, s in shapes.items():
        lk = _invert(k, modules, sf._RESNET, sf.UNET_PREFIX)
        if cfg.get("use_linear_projection") is not True and (k.endswith("proj_in.weight") or k.endswith("proj_out.weight")):
            s = tuple(s) + (1, 1) if len(s) == 2 else s          # SD1.x keeps them as 1x1 convs
        ldm[lk] = s
    assert len(ldm) == len(shapes)                               # injective
    assert any(k.startswith("model.diffusion_model.input_blocks.1.1.transformer_blocks.0.attn2.to_k") for k in ldm) \
        or name == "sdxl"
    assert "model.diffusion_model.out.2.weight" in ldm and "model.diffusion_model.middle_block.1.norm.weight" in ldm
    back = sf.convert_unet(_meta(ldm), cfg)
    assert set(back) == set(shapes)
    for k in shapes:
        assert tuple(back[k].shape)[:2] == tuple(shapes[k])[:2] and back[k].numel() == torch.empty(shapes[k], device="meta").numel(), k


def test_unknown_unet_key_is_an_error_not_a_silent_drop():
    bad = {"model.diffusion_model.input_blocks.77.0.in_layers.0.weight": torch.empty(4, device="meta")}
    with pytest.raises(RuntimeError, match="no diffusers name"):
        sf.convert_unet(bad, dict(sf.SD15_UNET))


@pytest.mark.parametrize("cfg", [sf.SD_VAE, sf.SDXL_VAE])
def test_vae_decoder_rename_and_attention_reshape(cfg):
    shapes = syn.vae_decoder_shapes(SimpleNamespace(**{**cfg, "block_out_channels": tuple(cfg["block_out_channels"])}))
    modules, attn = sf.vae_decoder_module_map(cfg)
    ldm = {}
    for k, s in shapes.items():
        is_attn = "attentions" in k
        lk = _invert(k, modules, sf._VAE_ATTN if is_attn else sf._VAE_RES, sf.VAE_PREFIX)
        if is_attn and len(s) == 2:
            s = tuple(s) + (1, 1)                                # the original layout stores q/k/v/proj_out as 1x1 convs
        ldm[lk] = s
    ldm["first_stage_model.encoder.conv_in.weight"] = (128, 3, 3, 3)     # encoder / quant_conv are skipped
    ldm["first_stage_model.quant_conv.weight"] = (8, 8, 1, 1)
    assert "first_stage_model.decoder.up.3.block.0.norm1.weight" in ldm          # lowest-resolution level
    assert "first_stage_model.decoder.mid.attn_1.q.weight" in ldm
    back = sf.convert_vae_decoder(_meta(ldm), cfg)
    assert set(back) == set(shapes)
    for k in shapes:
        assert tuple(back[k].shape) == tuple(shapes[k]), k


def test_openclip_text_tower_conversion_matches_transformers():
    """Round trip through the OpenCLIP naming (fused in_proj, transposed text_projection) on a small
    CLIPTextModelWithProjection: the conversion gives back the transformers state dict exactly."""
    from transformers import CLIPTextConfig, CLIPTextModelWithProjection
    cfg = CLIPTextConfig(vocab_size=1000, hidden_size=64, intermediate_size=128, num_hidden_layers=3,
                         num_attention_heads=4, max_position_embeddings=77, hidden_act="gelu", projection_dim=32)
    torch.manual_seed(0)
    ref = CLIPTextModelWithProjection(cfg).state_dict()
    ref = {k: v for k, v in ref.items() if not k.endswith("position_ids")}
    p = "conditioner.embedders.1.model."
    oc = {p + "token_embedding.weight": ref["text_model.embeddings.token_embedding.weight"],
          p + "positional_embedding": ref["text_model.embeddings.position_embedding.weight"],
          p + "ln_final.weight": ref["text_model.final_layer_norm.weight"],
          p + "ln_final.bias": ref["text_model.final_layer_norm.bias"],
          p + "text_projection": ref["text_projection.weight"].t().contiguous(),
          p + "logit_scale": torch.tensor(1.0)}
    for i in range(3):
        b, r = f"text_model.encoder.layers.{i}.", p + f"transformer.resblocks.{i}."
        oc[r + "attn.in_proj_weight"] = torch.cat([ref[b + f"self_attn.{n}_proj.weight"] for n in "qkv"], 0)
        oc[r + "attn.in_proj_bias"] = torch.cat([ref[b + f"self_attn.{n}_proj.bias"] for n in "qkv"], 0)
        for a, c in (("attn.out_proj", "self_attn.out_proj"), ("ln_1", "layer_norm1"), ("ln_2", "layer_norm2"),
                     ("mlp.c_fc", "mlp.fc1"), ("mlp.c_proj", "mlp.fc2")):
            for w in ("weight", "bias"):
                oc[r + f"{a}.{w}"] = ref[b + f"{c}.{w}"]
    back = sf.convert_openclip(oc, p, dict(hidden_size=64))
    assert set(back) == set(ref)
    assert all(torch.equal(back[k], ref[k]) for k in ref)
    hf = {"cond_stage_model.transformer." + k: v for k, v in ref.items()}
    assert set(sf.convert_clip_hf(hf, "cond_stage_model.transformer.")) == set(ref)


def _write_single_file(path, lcm=True):
    """The SD1.5(-LCM) random-init fixture in the ORIGINAL key layout, from the same seeds as write_model_dir."""
    from safetensors.torch import save_file
    ucfg = dict(sf.SD15_UNET)
    if lcm:
        ucfg["time_cond_proj_dim"] = 256
    unet = syn.random_state_dict(syn.unet_shapes(_ns(ucfg)), 0, torch.float16)
    vae = syn.random_state_dict(syn.vae_decoder_shapes(syn.sd_vae_cfg()), 1, torch.float16)
    out = {}
    um = sf.unet_module_map(ucfg)
    for k, v in unet.items():
        if k.endswith("proj_in.weight") or k.endswith("proj_out.weight"):
            v = v.reshape(*v.shape[:2], 1, 1)
        out[_invert(k, um, sf._RESNET, sf.UNET_PREFIX)] = v.contiguous()
    vm, _ = sf.vae_decoder_module_map(sf.SD_VAE)
    for k, v in vae.items():
        is_attn = "attentions" in k
        if is_attn and v.dim() == 2:
            v = v.reshape(*v.shape, 1, 1)
        out[_invert(k, vm, sf._VAE_ATTN if is_attn else sf._VAE_RES, sf.VAE_PREFIX)] = v.contiguous()
    save_file(out, path)


@pytest.mark.gpu
def test_worker_from_single_file_equals_worker_from_diffusers_dir(tmp_path):
    from backends.worker_factory import create_cuda_worker, detect_worker_type
    root = tmp_path
    syn.write_model_dir(str(root / "sd15-lcm"))
    _write_single_file(str(root / "sd15-lcm.safetensors"))
    assert sf.sniff(str(root / "sd15-lcm.safetensors")) == {"variant": "sd15", "cross_attention_dim": 768}
    old = {k: os.environ.get(k) for k in ("MODEL_ROOT", "MODEL", "CUDA_DEVICE")}
    os.environ["MODEL_ROOT"] = str(root)
    os.environ.pop("CUDA_DEVICE", None)
    try:
        pngs = []
        for model in ("sd15-lcm", "sd15-lcm.safetensors"):
            os.environ["MODEL"] = model
            assert detect_worker_type() == "sd15"
            w = create_cuda_worker(worker_id=0)
            job = SimpleNamespace(req=SimpleNamespace(prompt="a lighthouse", size="128x128", num_inference_steps=2,
                                                      guidance_scale=1.0, seed=5))
            pngs.append(w.run_job(job)[0])
            del w
            torch.cuda.empty_cache()
        assert pngs[0][:8] == b"\x89PNG\r\n\x1a\n" and pngs[0] == pngs[1]
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
