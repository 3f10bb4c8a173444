"""On-device CLIP text tower against the real `transformers` implementation (fp32, CPU) on the same
random-init weights and token ids: last_hidden_state, hidden_states[-2] (what SDXL conditions on),
pooler_output and the projected text_embeds.  This is the one component of the request path whose
third-party reference IS installable offline, so parity here is pinned against the library itself.
Bar: the bf16 tolerance of the hot path (max |d| / max |ref| <= 2e-2)."""
import pytest
import torch

from test_pipeline_gpu import NOISE_PRED_TOL, max_rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("hidden,heads,mlp,layers,act,proj", [
    (128, 2, 256, 2, "quick_gelu", None),        # tiny
    (768, 12, 3072, 3, "quick_gelu", None),      # CLIP-L geometry (SD1.5 / SDXL text_encoder), 3 layers
    (1280, 20, 5120, 2, "gelu", 1280),           # OpenCLIP bigG geometry (SDXL text_encoder_2), 2 layers
])
def test_clip_text_tower_vs_transformers(hidden, heads, mlp, layers, act, proj):
    from transformers import CLIPTextConfig, CLIPTextModel, CLIPTextModelWithProjection
    from dreamlab_b200.clip import CLIPTextB200, clip_cfg_from_json
    torch.manual_seed(0)
    cfg = CLIPTextConfig(vocab_size=2000, hidden_size=hidden, intermediate_size=mlp, num_hidden_layers=layers,
                         num_attention_heads=heads, max_position_embeddings=77, hidden_act=act,
                         projection_dim=proj or 512, eos_token_id=1999, bos_token_id=1998)
    ref = (CLIPTextModelWithProjection(cfg) if proj else CLIPTextModel(cfg)).eval()
    ids = torch.randint(0, 1990, (3, 77), generator=torch.Generator().manual_seed(1))
    ids[:, 0] = 1998
    for b, n in enumerate((5, 40, 76)):          # EOS position differs per prompt; padding after it
        ids[b, n:] = 1999
    with torch.no_grad():
        o = ref(ids, output_hidden_states=True)
    eng = CLIPTextB200(ref.state_dict(), clip_cfg_from_json(cfg.to_dict()), "cuda:0")
    got = eng.forward(ids, want_hidden=-2)
    torch.cuda.synchronize()
    ref_last = o.last_hidden_state if not proj else ref.text_model.final_layer_norm(o.hidden_states[-1])
    errs = {"last_hidden_state": max_rel_err(got["last_hidden_state"].float().cpu(), ref_last),
            "hidden_states[-2]": max_rel_err(got["hidden"].float().cpu(), o.hidden_states[-2])}
    if proj:
        errs["text_embeds"] = max_rel_err(got["text_embeds"].float().cpu(), o.text_embeds)
    else:
        errs["pooler_output"] = max_rel_err(got["pooler_output"].float().cpu(), o.pooler_output)
    print(f"CLIP {hidden}/{heads}h/{layers}L {act}: " + ", ".join(f"{k} {v:.2e}" for k, v in errs.items()))
    assert max(errs.values()) <= NOISE_PRED_TOL, errs


def test_causal_attention_matches_torch():
    """dl_attention(DL_ATTN_SIMT_CAUSAL): keys <= query only."""
    import math
    from dreamlab_b200 import lib
    torch.manual_seed(0)
    B, T, H, d = 2, 77, 4, 64
    hs = 80
    qkv = torch.randn(B * T, 3 * H * hs, device="cuda").bfloat16()
    out = torch.empty(B * T, H * d, device="cuda", dtype=torch.bfloat16)
    lib.attention(qkv, qkv[:, H * hs:], qkv[:, 2 * H * hs:], out, batch=B, sq=T, skv=T, heads=H, d=d, dh_stride=hs,
                  ldq=3 * H * hs, ldk=3 * H * hs, ldv=3 * H * hs, ldo=H * d, scale=1 / math.sqrt(d),
                  impl=lib.ATTN_SIMT_CAUSAL)
    x = qkv.float().view(B, T, 3, H, hs)[..., :d]
    q, k, v = (x[:, :, i].transpose(1, 2) for i in range(3))
    ref = torch.nn.functional.scaled_dot_product_attention(q, k, v, is_causal=True).transpose(1, 2).reshape(B * T, H * d)
    torch.cuda.synchronize()
    assert max_rel_err(out.float(), ref) <= 1e-2
