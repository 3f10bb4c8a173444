"""CLIP text tower on the device (the step right before the hot path).

Replaces `transformers.CLIPTextModel` / `CLIPTextModelWithProjection` as diffusers' `encode_prompt`
calls them behind `self.pipe(prompt=…)` (`backends/cuda_worker.py:222`, `:532`): token + position
embedding, N pre-LN transformer layers with causal self-attention and a (quick-)GELU MLP, final
LayerNorm; the pooled output is the final-LN state at the EOS position, optionally through
`text_projection`.  State-dict keys are transformers' (`text_model.…`), so `text_encoder/` and
`text_encoder_2/` of a diffusers model directory load unchanged.

Kernels: `dl_embed_tokens`, `dl_layernorm`, `dl_igemm` (fused QKV with bias, out-proj / fc2 with
the residual on the tensor core), `dl_attention` (causal CUDA-core flash kernel: 77 tokens),
`dl_act_bf16`.  bf16 storage, fp32 accumulation — parity against transformers in fp32:
`tests/test_clip_gpu.py`.
"""
from __future__ import annotations

import math
import os
from types import SimpleNamespace
from typing import Dict, Optional

import torch

from . import lib
from .weights import _bf, _f32, head_stride, pad_heads

BF16 = torch.bfloat16


def clip_cfg_from_json(c: dict):
    return SimpleNamespace(vocab_size=c.get("vocab_size", 49408), hidden_size=c.get("hidden_size", 768),
                           intermediate_size=c.get("intermediate_size", 3072),
                           num_hidden_layers=c.get("num_hidden_layers", 12),
                           num_attention_heads=c.get("num_attention_heads", 12),
                           max_position_embeddings=c.get("max_position_embeddings", 77),
                           hidden_act=c.get("hidden_act", "quick_gelu"),
                           layer_norm_eps=c.get("layer_norm_eps", 1e-5), eos_token_id=c.get("eos_token_id", 2),
                           projection_dim=c.get("projection_dim"))


class CLIPTextB200:
    def __init__(self, state_dict: Dict[str, torch.Tensor], cfg, device="cuda:0"):
        lib.require_cuda()
        lib.load()
        self.device = torch.device(device)
        self.cfg = cfg
        D, H = cfg.hidden_size, cfg.num_attention_heads
        if D % H or D % 64 or cfg.intermediate_size % 64:
            raise RuntimeError(f"CLIP text tower: hidden {D} / heads {H} / mlp {cfg.intermediate_size} unsupported")
        if cfg.hidden_act not in ("quick_gelu", "gelu"):
            raise RuntimeError(f"CLIP text tower: hidden_act={cfg.hidden_act!r} unsupported")
        self.d = D // H
        self.hs = head_stride(self.d)
        sd, dev, p = state_dict, self.device, "text_model."
        self.tok = _bf(sd[p + "embeddings.token_embedding.weight"], dev)
        self.pos = _bf(sd[p + "embeddings.position_embedding.weight"], dev)
        self.layers = []
        for i in range(cfg.num_hidden_layers):
            b = f"{p}encoder.layers.{i}."
            L = {}
            for n in ("layer_norm1", "layer_norm2"):
                L[n + "_w"], L[n + "_b"] = _f32(sd[b + n + ".weight"], dev), _f32(sd[b + n + ".bias"], dev)
            w = torch.cat([pad_heads(sd[b + f"self_attn.{n}_proj.weight"].float(), H) for n in "qkv"], 0)
            bias = torch.cat([pad_heads(sd[b + f"self_attn.{n}_proj.bias"].float()[:, None], H)[:, 0] for n in "qkv"], 0)
            L["qkv_w"], L["qkv_b"] = _bf(w, dev), _f32(bias, dev)
            L["o_w"], L["o_b"] = _bf(sd[b + "self_attn.out_proj.weight"], dev), _f32(sd[b + "self_attn.out_proj.bias"], dev)
            L["fc1_w"], L["fc1_b"] = _bf(sd[b + "mlp.fc1.weight"], dev), _f32(sd[b + "mlp.fc1.bias"], dev)
            L["fc2_w"], L["fc2_b"] = _bf(sd[b + "mlp.fc2.weight"], dev), _f32(sd[b + "mlp.fc2.bias"], dev)
            self.layers.append(L)
        self._graphs = {}
        self.fln_w, self.fln_b = _f32(sd[p + "final_layer_norm.weight"], dev), _f32(sd[p + "final_layer_norm.bias"], dev)
        self.proj = _bf(sd["text_projection.weight"], dev) if "text_projection.weight" in sd else None

    GRAPH_BUCKETS = (1, 2, 4, 8, 16, 32)

    @torch.no_grad()
    def forward_graphed(self, input_ids: torch.Tensor, want_hidden: Optional[int] = None):
        """forward() as ONE CUDA-graph replay per (batch bucket, T, want_hidden): the tower is ~150 small launches,
        5-6 ms of interpreter time per batch in front of a 118 ms UNet/VAE replay when issued eagerly.  The batch is
        padded to its bucket (the pad rows keep whatever ids the buffer held); batches beyond the largest bucket,
        or B200_CLIP_GRAPH=0, run eagerly.  The returned tensors are views of the graph's STATIC outputs: convert
        or copy them before the next call (the workers' `.float()` does)."""
        B, T = input_ids.shape
        bucket = next((b for b in self.GRAPH_BUCKETS if b >= B), None)
        if bucket is None or os.environ.get("B200_CLIP_GRAPH", "1") in ("0", "false"):
            return self.forward(input_ids, want_hidden)
        key = (bucket, T, want_hidden)
        g = self._graphs.get(key)
        with torch.cuda.device(self.device):
            if g is None:
                from .engine import CAPTURE_LOCK
                ids = torch.zeros(bucket, T, device=self.device, dtype=torch.int64)
                ids[:B].copy_(input_ids)
                with CAPTURE_LOCK:
                    s = torch.cuda.Stream(device=self.device)
                    s.wait_stream(torch.cuda.current_stream(self.device))
                    with torch.cuda.stream(s):
                        self.forward(ids, want_hidden)                     # warm-up: lazy per-device init
                    torch.cuda.current_stream(self.device).wait_stream(s)
                    torch.cuda.synchronize(self.device)
                    graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(graph, stream=s, capture_error_mode="thread_local"):
                        out = self.forward(ids, want_hidden)
                g = self._graphs[key] = (graph, ids, out)
            graph, ids, out = g
            # pinned staging + async copy: a pageable host-to-device copy would first wait for the stream to drain
            src = input_ids.pin_memory() if input_ids.device.type == "cpu" and not input_ids.is_pinned() else input_ids
            ids[:B].copy_(src, non_blocking=True)
            graph.replay()
            return {k: v[:B] for k, v in out.items()}

    def _lin(self, x, w, b, n, residual=None):
        M = x.shape[0]
        out = torch.empty(M, n, device=self.device, dtype=BF16)
        lib.igemm(x, w, out, nimg=1, h=1, w=M, taps=1, n=n, bias=b, residual=residual, ldo=n,
                  ldr=None if residual is None else n)
        return out

    @torch.no_grad()
    def forward(self, input_ids: torch.Tensor, want_hidden: Optional[int] = None):
        """input_ids int64 [B, T] -> dict(last_hidden_state [B,T,D] bf16, pooler_output [B,D],
        text_embeds [B,P] if the tower has a projection, hidden [B,T,D] = hidden_states[want_hidden]
        (transformers indexing: 0 = embeddings, -2 = input of the last layer) if asked)."""
        cfg = self.cfg
        B, T = input_ids.shape
        D, H, d, hs = cfg.hidden_size, cfg.num_attention_heads, self.d, self.hs
        eps = cfg.layer_norm_eps
        act = 0 if cfg.hidden_act == "quick_gelu" else 1
        with torch.cuda.device(self.device):
            ids = input_ids.to(self.device, torch.int64).contiguous().view(-1)
            x = torch.empty(B * T, D, device=self.device, dtype=BF16)
            lib.embed_tokens(ids, self.tok, self.pos, x, T)
            n_layers = len(self.layers)
            want = None if want_hidden is None else (want_hidden % (n_layers + 1))
            hidden = x if want == 0 else None
            for li, L in enumerate(self.layers):
                h = torch.empty_like(x)
                lib.layernorm(x, h, L["layer_norm1_w"], L["layer_norm1_b"], eps)
                qkv = self._lin(h, L["qkv_w"], L["qkv_b"], 3 * H * hs)
                a = torch.empty(B * T, D, device=self.device, dtype=BF16)
                lib.attention(qkv, qkv[:, H * hs:], qkv[:, 2 * H * hs:], a, batch=B, sq=T, skv=T, heads=H, d=d,
                              dh_stride=hs, ldq=3 * H * hs, ldk=3 * H * hs, ldv=3 * H * hs, ldo=D,
                              scale=1.0 / math.sqrt(d), impl=lib.ATTN_SIMT_CAUSAL)
                x = self._lin(a, L["o_w"], L["o_b"], D, residual=x)
                h = torch.empty_like(x)
                lib.layernorm(x, h, L["layer_norm2_w"], L["layer_norm2_b"], eps)
                m = self._lin(h, L["fc1_w"], L["fc1_b"], cfg.intermediate_size)
                lib.act_bf16(m, m, act)
                x = self._lin(m, L["fc2_w"], L["fc2_b"], D, residual=x)
                if want == li + 1:
                    hidden = x
            last = torch.empty_like(x)
            lib.layernorm(x, last, self.fln_w, self.fln_b, eps)
            last = last.view(B, T, D)
            # pooled: state at the EOS token (legacy configs with eos_token_id == 2: the highest id)
            ii = input_ids.to(self.device)
            pos = ii.argmax(-1) if cfg.eos_token_id == 2 else (ii == cfg.eos_token_id).int().argmax(-1)
            pooled = last[torch.arange(B, device=self.device), pos]
            out = {"last_hidden_state": last, "pooler_output": pooled}
            if hidden is not None:
                out["hidden"] = hidden.view(B, T, D)
            if self.proj is not None:
                out["text_embeds"] = self._lin(pooled.contiguous(), self.proj, None, self.proj.shape[0])
            return out
