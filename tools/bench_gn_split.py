"""fused cooperative GroupNorm vs the split (stats | apply) kernels on the big tensors (B=16)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dreamlab_b200 import lib

dev = "cuda"
B = 16
ws = torch.empty(lib.groupnorm_workspace_bytes(B), device=dev, dtype=torch.uint8)
ws2 = torch.zeros(lib.groupnorm_split_workspace_bytes(B, 32), device=dev, dtype=torch.uint8)
l2 = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
for name, hw, c in [("vae512", 262144, 128), ("vae512b", 262144, 256), ("vae256", 65536, 256), ("vae256b", 65536, 512),
                    ("vae128", 16384, 512), ("vae64", 4096, 512), ("L0", 4096, 320), ("L1", 1024, 640)]:
    x = torch.randn(B, hw, c, device=dev).bfloat16()
    out = torch.empty_like(x)
    g, b = torch.ones(c, device=dev), torch.zeros(c, device=dev)
    stats = torch.empty(B, 32, 2, device=dev)

    def fused():
        lib.groupnorm(x, out, g, b, ws, nimg=B, hw=hw, eps=1e-5, silu=True)

    def split():
        lib.groupnorm_stats(x, stats, ws2, nimg=B, hw=hw)
        lib.groupnorm_apply(x, out, g, b, stats.unsqueeze(0), nimg=B, hw=hw, eps=1e-5, silu=True)
    res = []
    for fn in (fused, split):
        for _ in range(2):
            fn()
        l2.zero_()
        torch.cuda._sleep(int(2e6))
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(4):
            fn()
        e.record()
        torch.cuda.synchronize()
        res.append(a.elapsed_time(e) / 4)
    nbytes = 4.0 * B * hw * c
    print(f"{name:8s} hw={hw:7d} C={c}: fused {res[0] * 1e3:8.1f} us {nbytes / res[0] / 1e6:6.0f} GB/s | "
          f"split {res[1] * 1e3:8.1f} us {nbytes / res[1] / 1e6:6.0f} GB/s", flush=True)
