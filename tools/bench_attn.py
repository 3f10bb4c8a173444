"""Micro-benchmark of dl_attention (tcgen05 flash kernel) on the UNet shapes."""
import math
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dreamlab_b200 import lib

dev = "cuda"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
for (S, Skv, heads, d) in [(4096, 4096, 8, 40), (1024, 1024, 8, 80), (256, 256, 8, 160), (4096, 77, 8, 40)]:
    hs = (d + 1 + 15) // 16 * 16
    qkv = torch.randn(B * S, 3 * heads * hs, device=dev).bfloat16()
    kv = torch.randn(B * Skv, 2 * heads * hs, device=dev).bfloat16()
    out = torch.empty(B * S, heads * d, device=dev, dtype=torch.bfloat16)

    def run():
        if S == Skv:
            lib.attention(qkv, qkv[:, heads * hs:], qkv[:, 2 * heads * hs:], out, batch=B, sq=S, skv=Skv,
                          heads=heads, d=d, dh_stride=hs, ldq=3 * heads * hs, ldk=3 * heads * hs,
                          ldv=3 * heads * hs, ldo=heads * d, scale=1 / math.sqrt(d), v_ones=True)
        else:
            lib.attention(qkv, kv, kv[:, heads * hs:], out, batch=B, sq=S, skv=Skv, heads=heads, d=d,
                          dh_stride=hs, ldq=3 * heads * hs, ldk=2 * heads * hs, ldv=2 * heads * hs,
                          ldo=heads * d, scale=1 / math.sqrt(d), v_ones=True)
    for _ in range(2):
        run()
    torch.cuda._sleep(int(2e6))
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(4):
        run()
    b.record()
    torch.cuda.synchronize()
    t = a.elapsed_time(b) / 4
    fl = 4.0 * B * heads * S * Skv * d
    print(f"B={B} S={S} Skv={Skv} h={heads} d={d}: {t * 1e3:8.1f} us  {fl / t / 1e9:7.1f} TFLOP/s", flush=True)
