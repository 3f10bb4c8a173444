"""Micro-benchmark of dl_groupnorm on UNet / VAE shapes (B=16)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dreamlab_b200 import lib

dev = "cuda"
only = sys.argv[1] if len(sys.argv) > 1 else None
B = 16
ws = torch.empty(lib.groupnorm_workspace_bytes(B), device=dev, dtype=torch.uint8)
l2 = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
for name, hw, c0, c1 in [("vae512", 262144, 128, 0), ("vae256", 65536, 256, 0), ("vae128", 16384, 512, 0),
                         ("vae64", 4096, 512, 0), ("L0", 4096, 320, 0), ("L0cat", 4096, 640, 320),
                         ("L1", 1024, 640, 0), ("L2", 256, 1280, 0), ("L3", 64, 1280, 0)]:
    if only and only != name:
        continue
    x0 = torch.randn(B, hw, c0, device=dev).bfloat16()
    x1 = torch.randn(B, hw, c1, device=dev).bfloat16() if c1 else None
    C = c0 + c1
    out = torch.empty(B, hw, C, device=dev, dtype=torch.bfloat16)
    g, b = torch.ones(C, device=dev), torch.zeros(C, device=dev)

    def run():
        lib.groupnorm(x0, out, g, b, ws, nimg=B, hw=hw, eps=1e-5, silu=True, x1=x1)
    for _ in range(2):
        run()
    l2.zero_()
    torch.cuda._sleep(int(2e6))
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(4):
        run()
    e.record()
    torch.cuda.synchronize()
    t = a.elapsed_time(e) / 4
    nbytes = 4.0 * B * hw * C
    print(f"{name:8s} hw={hw:7d} C={c0}+{c1}: {t * 1e3:8.1f} us  {nbytes / t / 1e6:7.0f} GB/s algorithmic", flush=True)
