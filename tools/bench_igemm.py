"""Micro-benchmark of dl_igemm over the hot shapes of the UNet/VAE (B=16), sweeping the N tile.
Prints time, TFLOP/s and effective HBM GB/s (A + out (+res) bytes) per (shape, bn)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dreamlab_b200 import lib

dev = "cuda"
SHAPES = [
    # (name, nimg, h, w, cin, taps, n, mode, residual)
    ("lin320 q2", 1, 1, 65536, 320, 1, 384, 0, False),
    ("lin320 proj", 1, 1, 65536, 320, 1, 320, 0, False),
    ("lin320 +res", 1, 1, 65536, 320, 1, 320, 0, True),
    ("lin320 qkv", 1, 1, 65536, 320, 1, 1152, 0, False),
    ("geglu320", 1, 1, 65536, 320, 1, 2560, 1, False),
    ("ff2 320", 1, 1, 65536, 1280, 1, 320, 0, True),
    ("lin640 +res", 1, 1, 16384, 640, 1, 640, 0, True),
    ("geglu640", 1, 1, 16384, 640, 1, 5120, 1, False),
    ("lin1280 +res", 1, 1, 4096, 1280, 1, 1280, 0, True),
    ("conv 64x64 320", 16, 64, 64, 320, 9, 320, 0, True),
    ("conv 32x32 640", 16, 32, 32, 640, 9, 640, 0, True),
    ("conv 16x16 1280", 16, 16, 16, 1280, 9, 1280, 0, True),
    ("conv 8x8 1280", 16, 8, 8, 1280, 9, 1280, 0, True),
    ("conv 8x8 2560", 16, 8, 8, 2560, 9, 1280, 0, False),
    ("vae 512 128", 4, 512, 512, 128, 9, 128, 0, True),
    ("vae 256 256", 4, 256, 256, 256, 9, 256, 0, True),
]
only = sys.argv[1] if len(sys.argv) > 1 else None
l2 = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
for name, nimg, h, w, cin, taps, n, mode, res in SHAPES:
    if only and only not in name:
        continue
    x = torch.randn(nimg, h, w, cin, device=dev).bfloat16()
    wgt = (torch.randn(n, taps * cin, device=dev) * (taps * cin) ** -0.5).bfloat16()
    ncols = n // 2 if mode == 1 else n
    out = torch.empty(nimg, h, w, ncols, device=dev, dtype=torch.bfloat16)
    bias = torch.randn(n, device=dev)
    r = torch.randn(nimg, h, w, n, device=dev).bfloat16() if res else None
    M = nimg * h * w
    flops = 2.0 * M * n * taps * cin
    nbytes = 2.0 * (M * cin + M * ncols + (M * n if res else 0))
    cands = [0] + [b for b in (256, 192, 160, 128, 96, 64) if n % b == 0 and (mode != 1 or b % 64 == 0)]
    line = f"{name:18s} M={M:8d} N={n:5d} K={taps}x{cin:4d}: "
    for bn in cands:
        def run():
            lib.igemm(x, wgt, out, nimg=nimg, h=h, w=w, taps=taps, n=n, bias=bias, residual=r,
                      mode=mode, bn=bn)
        for _ in range(2):
            run()
        ts = []
        for _ in range(3):
            l2.zero_()
            torch.cuda._sleep(int(2e6))          # keep the GPU busy while the launches queue up
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(8):
                run()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) / 8)
        t = sorted(ts)[len(ts) // 2]
        line += f" bn{bn if bn else 'A'}:{t * 1e3:7.1f}us {flops / t / 1e9:6.0f}TF {nbytes / t / 1e6:5.0f}GB/s |"
    print(line, flush=True)
