"""One shape of dl_attention (default: the SD1.5 64^2-level self-attention at batch 16), for A/B runs and ncu."""
import math
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dreamlab_b200 import lib

B, S, heads, d = (int(x) for x in (sys.argv[1:5] + ["16", "4096", "8", "40"][len(sys.argv) - 1:]))
reps = int(os.environ.get("REPS", "6"))
hs = (d + 1 + 15) // 16 * 16
qkv = torch.randn(B * S, 3 * heads * hs, device="cuda").bfloat16()
out = torch.empty(B * S, heads * d, device="cuda", dtype=torch.bfloat16)


def run():
    lib.attention(qkv, qkv[:, heads * hs:], qkv[:, 2 * heads * hs:], out, batch=B, sq=S, skv=S, heads=heads, d=d,
                  dh_stride=hs, ldq=3 * heads * hs, ldk=3 * heads * hs, ldv=3 * heads * hs, ldo=heads * d,
                  scale=1 / math.sqrt(d), v_ones=True)


for _ in range(2):
    run()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(reps):
    run()
b.record()
torch.cuda.synchronize()
t = a.elapsed_time(b) / reps
fl = 4.0 * B * heads * S * S * d
tag = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("DL_ATTN"))
print(f"B={B} S={S} h={heads} d={d} [{tag}]: {t * 1e3:8.1f} us  {fl / t / 1e9:7.1f} TFLOP/s", flush=True)
