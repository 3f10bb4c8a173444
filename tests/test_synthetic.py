"""Product-side shape tables (dreamlab_b200.synthetic) agree key-for-key, shape-for-shape with
the oracle modules (and hence with the published parameter counts)."""
import torch

from dreamlab_b200 import synthetic as S
from oracle.unet import OracleUNet, UNetConfig
from oracle.vae import OracleVAEDecoder, VAEConfig


def _shapes(mod):
    return {k: tuple(v.shape) for k, v in mod.state_dict().items()}


def test_unet_and_vae_shape_tables_match_oracle():
    with torch.device("meta"):
        u, v = OracleUNet(UNetConfig()), OracleVAEDecoder(VAEConfig())
        ut, vt = OracleUNet(UNetConfig.tiny()), OracleVAEDecoder(VAEConfig.tiny())
    assert S.unet_shapes(S.sd15_lcm_unet_cfg()) == _shapes(u)
    assert S.vae_decoder_shapes(S.sd_vae_cfg()) == _shapes(v)
    assert S.unet_shapes(ut.cfg) == _shapes(ut)
    assert S.vae_decoder_shapes(vt.cfg) == _shapes(vt)
    n = sum(torch.Size(s).numel() for s in S.unet_shapes(S.sd15_lcm_unet_cfg()).values())
    assert n == 859_602_884
    with torch.device("meta"):
        x, xt = OracleUNet(UNetConfig.sdxl_base()), OracleUNet(UNetConfig.tiny_sdxl())
    assert S.unet_shapes(S.sdxl_unet_cfg()) == _shapes(x)
    assert S.unet_shapes(xt.cfg) == _shapes(xt)
    assert sum(torch.Size(s).numel() for s in S.unet_shapes(S.sdxl_unet_cfg()).values()) == 2_567_463_684


def test_random_state_dict_is_deterministic_and_bounded():
    shp = S.vae_decoder_shapes(VAEConfig.tiny())
    a, b = S.random_state_dict(shp, 3), S.random_state_dict(shp, 3)
    assert all(torch.equal(a[k], b[k]) for k in a)
    w = a["decoder.conv_in.weight"]
    assert w.abs().max() <= 1 / (4 * 9) ** 0.5 + 1e-6
    assert torch.equal(a["decoder.conv_norm_out.weight"], torch.ones(64))


def test_synthetic_inputs_match_oracle_protocol():
    from oracle.pipeline import synthetic_inputs as oracle_inputs
    a = S.synthetic_inputs(3, 64, 64, 4)
    b = oracle_inputs(3, 64, 64, 4)
    assert all(torch.equal(x, y) for x, y in zip(a, b))


def test_write_model_dir_roundtrip(tmp_path):
    from backends.b200_worker import _load_component, unet_cfg_from_json, vae_cfg_from_json
    p = S.write_model_dir(str(tmp_path / "m"), UNetConfig.tiny(), VAEConfig.tiny())
    cj, sd = _load_component(p + "/unet")
    cfg = unet_cfg_from_json(cj)
    assert cfg.block_out_channels == (64, 128, 256, 256) and cfg.down_attn == (True, True, True, False)
    assert set(sd) == set(S.unet_shapes(cfg))
    vj, vsd = _load_component(p + "/vae")
    assert vae_cfg_from_json(vj).sample_size == 128 and set(vsd) == set(S.vae_decoder_shapes(VAEConfig.tiny()))


def test_synthetic_text_encoder_has_the_transformers_layout(tmp_path):
    """write_text_encoder (the CLIP tower bench.py ships in its model dir so that the e2e leg runs prompt -> token ids
    -> on-device text tower): every key and shape is what transformers' CLIPTextModel expects, embeddings at the
    N(0, 0.02) scale transformers initialises them with."""
    import json
    import os
    from safetensors.torch import load_file
    from transformers import CLIPTextConfig, CLIPTextModel
    from dreamlab_b200 import synthetic as syn
    syn.write_text_encoder(str(tmp_path), layers=2)
    te = os.path.join(str(tmp_path), "text_encoder")
    cj = json.load(open(os.path.join(te, "config.json")))
    cfg = CLIPTextConfig(**{k: v for k, v in cj.items() if not k.startswith("_") and k not in ("architectures", "model_type")})
    sd = load_file(os.path.join(te, "model.safetensors"))
    r = CLIPTextModel(cfg).load_state_dict({k: v.float() for k, v in sd.items()}, strict=False)
    assert r.missing_keys == [] and r.unexpected_keys == []
    assert abs(float(sd["text_model.embeddings.token_embedding.weight"].float().std()) - 0.02) < 2e-3


def test_hashed_stand_in_token_ids():
    """_hash_tokens (no tokenizer files in the model dir): BOS, one id per UTF-8 byte (at most 75), EOS padding; ids
    inside the vocabulary; a function of the prompt alone."""
    from backends.b200_worker import _hash_tokens
    ids = _hash_tokens(["", "ab", "x" * 200, "ü"])
    assert ids.shape == (4, 77) and ids.dtype == torch.long
    assert (ids[:, 0] == 49406).all() and (ids[0, 1:] == 49407).all()
    assert ids[1, 1] == (ord("a") * 193) % 49406 and ids[1, 2] == (ord("b") * 193 + 7919) % 49406 and ids[1, 3] == 49407
    assert (ids[2, 1:76] != 49407).all() and ids[2, 76] == 49407 and int(ids.max()) <= 49407 and int(ids.min()) >= 0
    assert ids[3, 1] == (0xC3 * 193) % 49406 and ids[3, 2] == (0xBC * 193 + 7919) % 49406
    assert torch.equal(_hash_tokens(["ab"])[0], ids[1])


def test_layernorm_fold_identity_on_the_host():
    """weights.fold_layernorm: LN(x) W^T + b == rstd * (x W'^T - mean * colsum(W')) + (b + W beta) with
    W' = W * gamma — the identity the GEMM epilogue applies (dl_igemm_desc.ln_*), checked in fp64 on the host with
    the bf16-rounded W' and its column sums exactly as the packer hands them to the kernel."""
    from dreamlab_b200.weights import fold_layernorm
    g = torch.Generator().manual_seed(3)
    M, K, N, eps = 37, 320, 96, 1e-5
    x = torch.randn(M, K, generator=g) * 2 + 0.5
    w = torch.randn(N, K, generator=g) * K ** -0.5
    bias, gamma, beta = torch.randn(N, generator=g), 1 + 0.2 * torch.randn(K, generator=g), 0.3 * torch.randn(K, generator=g)
    wf, colsum, bf = fold_layernorm(w, bias, gamma, beta, "cpu")
    assert wf.dtype == torch.bfloat16 and colsum.dtype == torch.float32 and bf.dtype == torch.float32
    xd = x.double()
    mean, var = xd.mean(1, keepdim=True), xd.var(1, unbiased=False, keepdim=True)
    rstd = (var + eps).rsqrt()
    folded = rstd * (xd @ wf.double().t() - mean * colsum.double()[None, :]) + bf.double()[None, :]
    # the same weights the kernel multiplies (bf16 W'), un-folded: LayerNorm without affine, then W' and the folded bias
    direct = ((xd - mean) * rstd) @ wf.double().t() + bf.double()[None, :]
    assert torch.allclose(folded, direct, rtol=0, atol=1e-5)        # colsum is kept in fp32
    # and against torch's LayerNorm -> Linear in fp32: only the bf16 rounding of W' (and of W in the beta term) apart
    ref = torch.nn.functional.linear(torch.nn.functional.layer_norm(x, (K,), gamma, beta, eps), w, bias)
    err = (folded.float() - ref).abs().max() / ref.abs().max()
    assert err < 1e-2, float(err)
