"""Style registry — the request-side contract of the reference's `backends/styles.py`:
a style id names exactly one LoRA (exclusive selection) and `level` (1..N) indexes a ladder of
adapter weights; level 0 / no style = off.  Same names (`StyleDef`, `StyleRequest`,
`STYLE_REGISTRY`, `parse_style_request`) so the server-side callers keep working.  The registry is
read from `B200_STYLES` (a JSON file: {id: {title, lora_path, adapter_name, levels,
required_cross_attention_dim}}) instead of being edited in source; empty by default.
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass
from typing import Dict, Optional, Sequence


@dataclass(frozen=True)
class StyleDef:
    id: str
    title: str
    lora_path: str
    adapter_name: str
    levels: Sequence[float]
    required_cross_attention_dim: Optional[int] = 768


@dataclass(frozen=True)
class StyleRequest:
    style_id: Optional[str] = None
    level: int = 0

    def is_enabled(self) -> bool:
        return bool(self.style_id) and self.level > 0

    def weight(self, registry: Dict[str, StyleDef]) -> Optional[float]:
        if not self.is_enabled():
            return None
        sd = registry.get(self.style_id or "")
        if not sd:
            return None
        return float(sd.levels[clamp_level(self.level, len(sd.levels)) - 1])


def clamp_level(level: int, n: int) -> int:
    """levels are 1-indexed and clamp to the ladder (reference `cuda_worker.py:181-182`)."""
    return max(1, min(int(level), n))


def parse_style_request(params: dict) -> StyleRequest:
    style = params.get("style")
    obj = params.get("style_lora")
    if isinstance(obj, dict):
        style = obj.get("style") or obj.get("id") or style
        level = obj.get("level", 0)
    else:
        level = params.get("style_level", 0)
    if style in (None, "", "none", "off", False):
        style = None
    try:
        level = int(level or 0)
    except Exception:
        level = 0
    return StyleRequest(style_id=style, level=level)


def load_registry(path: Optional[str] = None) -> Dict[str, StyleDef]:
    path = path or os.environ.get("B200_STYLES", "").strip()
    if not path:
        return {}
    with open(path) as f:
        raw = json.load(f)
    out = {}
    for sid, d in raw.items():
        out[sid] = StyleDef(id=sid, title=d.get("title", sid), lora_path=d["lora_path"],
                            adapter_name=d.get("adapter_name", f"style_{sid}"),
                            levels=tuple(float(x) for x in d["levels"]),
                            required_cross_attention_dim=d.get("required_cross_attention_dim", 768))
    return out


STYLE_REGISTRY: Dict[str, StyleDef] = load_registry()
