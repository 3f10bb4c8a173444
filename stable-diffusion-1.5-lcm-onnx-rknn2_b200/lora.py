"""Style LoRAs as weight-space updates of the packed UNet weights.

What this replaces in the reference: `pipe.load_lora_weights(path, adapter_name=…)` at worker
construction and `pipe.set_adapters([name], adapter_weights=[w])` / `pipe.disable_lora()` around
every job (`backends/cuda_worker.py:123-196`; registry `backends/styles.py`).  diffusers keeps the
adapter un-merged and adds `w * (alpha/r) * up(down(x))` at run time; here the same linear map is
folded into the weights the kernels already read:

    W_eff = W + w * (alpha / r) * (up @ down)

Packing (head padding, Q/K/V fusion, GEGLU interleave, OHWI conv layout, the folded upsample,
the fused time_emb_proj matrix) is linear in the weights, so the packed-layout delta of a LoRA is
`pack(dW)`: it is computed once per style at load, kept on the device next to an fp32 master of
every tensor it touches, and switching a style on, off or to another level is an in-place
`target = bf16(master + w * delta)` on those tensors — pointers never change, captured CUDA
graphs stay valid.

File formats: kohya (`lora_unet_<path_with_underscores>.lora_down.weight / .lora_up.weight /
.alpha`), peft (`unet.<path>.lora_A.weight / lora_B.weight`), old diffusers
(`unet.<path>.lora.down.weight / lora.up.weight`, `…processor.to_q_lora.down.weight`).  Text-encoder
entries (`lora_te*`, `text_encoder.*`) are outside the hot path and are skipped (reported).
"""
from __future__ import annotations

import re
from typing import Dict, List, Tuple

import torch

# constructed tensors that are not linear images of checkpoint weights (the ones column of V)
_NONLINEAR = {"qkv_b", "kv2_b"}


def _module_table(unet_keys) -> Dict[str, str]:
    """kohya-style flattened module name -> dotted diffusers module path, for every module of
    the UNet that has a weight."""
    table = {}
    for k in unet_keys:
        if k.endswith(".weight"):
            mod = k[:-7]
            table["lora_unet_" + mod.replace(".", "_")] = mod
    return table


_PEFT = re.compile(r"^(?:unet\.)?(?P<mod>.+?)\.(?:lora_A|lora\.down|lora_down)(?:\.default[^.]*)?\.weight$")
_OLD_PROC = re.compile(r"^(?:unet\.)?(?P<attn>.+?)\.processor\.(?P<proj>to_q|to_k|to_v|to_out)_lora\.down\.weight$")


def lora_weight_deltas(unet_shapes: Dict[str, Tuple[int, ...]], lora_sd: Dict[str, torch.Tensor]):
    """-> ({diffusers weight key: dW fp32 with the base weight's shape}, skipped entry names).
    dW = (alpha / r) * up @ down for unit adapter weight."""
    table = _module_table(unet_shapes.keys())
    pairs = {}           # module path -> [down, up, alpha]

    def slot(mod):
        return pairs.setdefault(mod, [None, None, None])

    skipped = []
    for k, v in lora_sd.items():
        if k.startswith(("lora_te", "text_encoder", "lora_te1_", "lora_te2_")):
            skipped.append(k)
            continue
        if k.startswith("lora_unet_"):
            name, _, tail = k.partition(".")
            mod = table.get(name)
            if mod is None:
                skipped.append(k)
                continue
            if tail == "lora_down.weight":
                slot(mod)[0] = v
            elif tail == "lora_up.weight":
                slot(mod)[1] = v
            elif tail == "alpha":
                slot(mod)[2] = float(v)
            else:
                skipped.append(k)
            continue
        m = _OLD_PROC.match(k)
        if m:
            proj = "to_out.0" if m.group("proj") == "to_out" else m.group("proj")
            mod = f"{m.group('attn')}.{proj}"
            up_key = k.replace(".down.weight", ".up.weight")
            if mod + ".weight" in unet_shapes and up_key in lora_sd:
                s = slot(mod)
                s[0], s[1] = v, lora_sd[up_key]
            else:
                skipped.append(k)
            continue
        m = _PEFT.match(k)
        if m:
            mod = m.group("mod")
            up_key = (k.replace("lora_A", "lora_B").replace("lora.down", "lora.up")
                      .replace("lora_down", "lora_up"))
            if mod + ".weight" in unet_shapes and up_key in lora_sd:
                s = slot(mod)
                s[0], s[1] = v, lora_sd[up_key]
                a = lora_sd.get(f"{'unet.' if k.startswith('unet.') else ''}{mod}.alpha")
                if a is not None:
                    s[2] = float(a)
            else:
                skipped.append(k)
            continue
        if not (k.endswith(("lora_B.weight", "lora.up.weight", "lora_up.weight", ".alpha"))
                or ".up.weight" in k):
            skipped.append(k)
    deltas = {}
    for mod, (down, up, alpha) in pairs.items():
        if down is None or up is None:
            skipped.append(mod + " (incomplete pair)")
            continue
        shape = unet_shapes[mod + ".weight"]
        r = down.shape[0]
        scale = (alpha / r) if alpha is not None else 1.0
        d2 = down.float().reshape(r, -1)                      # [r, in (*kh*kw)]
        u2 = up.float().reshape(up.shape[0], r)               # [out, r]
        dw = (u2 @ d2) * scale
        if dw.numel() != int(torch.Size(shape).numel()):
            skipped.append(mod + f" (shape {tuple(dw.shape)} does not fit {shape})")
            continue
        deltas[mod + ".weight"] = dw.reshape(shape)
    return deltas, skipped


def _leaves(P, D, path, out):
    if isinstance(P, dict):
        for k in P:
            if k in ("cfg", "transformers") or k not in D:     # "transformers" aliases down/mid/up
                continue
            _leaves(P[k], D[k], path + (k,), out)
    elif isinstance(P, (list, tuple)):
        for i, (a, b) in enumerate(zip(P, D)):
            _leaves(a, b, path + (i,), out)
    elif torch.is_tensor(P) and torch.is_tensor(D) and P.shape == D.shape:
        if path[-1] not in _NONLINEAR:
            out.append((path, P, D))


class StyleAdapter:
    """One loaded style: the packed tensors it touches, their fp32 masters and packed deltas."""

    def __init__(self, name: str, entries: List[tuple], skipped: List[str]):
        self.name = name
        self.entries = entries            # [(path, target tensor, master fp32, delta fp32)]
        self.skipped = skipped

    @property
    def num_tensors(self) -> int:
        return len(self.entries)


def build_style_adapter(name: str, unet_engine, unet_shapes, lora_sd) -> StyleAdapter:
    """unet_engine: UNetB200 (its packed weights `P` are the update targets)."""
    deltas, skipped = lora_weight_deltas(unet_shapes, lora_sd)
    if not deltas:
        raise RuntimeError(f"LoRA '{name}': no UNet entries matched this model "
                           f"({len(lora_sd)} tensors in the file, {len(skipped)} skipped)")
    dev = unet_engine.device
    # pack a state dict that is zero everywhere except the deltas: packing is linear
    dsd = {k: (deltas[k] if k in deltas else torch.zeros(shp)) for k, shp in unet_shapes.items()}
    DP = pack_unet_f32(dsd, unet_engine.cfg, dev)
    leaves = []
    _leaves(unet_engine.P, DP, (), leaves)
    entries = []
    for path, target, delta in leaves:
        if float(delta.abs().max()) == 0.0:
            continue
        entries.append((path, target, None, delta.to(torch.float32)))
    del DP
    return StyleAdapter(name, entries, skipped)


def pack_unet_f32(sd, cfg, device):
    """pack_unet with fp32 leaves (deltas are far below bf16 resolution of the base weights)."""
    from . import weights as Wm
    with Wm.pack_dtype(torch.float32):
        return Wm.pack_unet(sd, cfg, device)


class StyleManager:
    """Exclusive style selection over the packed weights of one UNet (the reference's
    `_apply_style`: one adapter at a time, or none)."""

    def __init__(self, unet_engine, unet_shapes):
        self.unet = unet_engine
        self.shapes = unet_shapes
        self.adapters: Dict[str, StyleAdapter] = {}
        self._masters = {}                 # id(target) -> fp32 copy of the un-adapted tensor
        self._dirty = []                   # targets currently carrying an adapter
        self.active = (None, 0.0)

    def load(self, name: str, lora_sd) -> StyleAdapter:
        ad = build_style_adapter(name, self.unet, self.shapes, lora_sd)
        # the adapters update the plain packed projections in place; the gamma-folded copies used by the
        # LayerNorm fold (weights.fold_layernorm) would go stale, so a UNet that carries style adapters keeps
        # the standalone LayerNorm kernel
        self.unet.fold_ln = False
        for i, (path, target, _, delta) in enumerate(ad.entries):
            key = id(target)
            if key not in self._masters:
                self._masters[key] = target.detach().to(torch.float32).clone()
            ad.entries[i] = (path, target, self._masters[key], delta)
        self.adapters[name] = ad
        return ad

    def disable(self):
        for target, master in self._dirty:
            target.copy_(master)
        self._dirty = []
        self.active = (None, 0.0)

    def set_adapter(self, name: str, weight: float):
        if self.active == (name, float(weight)):
            return
        self.disable()
        ad = self.adapters[name]
        for _, target, master, delta in ad.entries:
            target.copy_(torch.add(master, delta, alpha=float(weight)))
            self._dirty.append((target, master))
        self.active = (name, float(weight))
