"""VAE mid-block attention: the wide flash kernel alone, and a whole eager VAE decode with it on / off.
  python tools/bench_vae_attn.py [batch] [size]"""
import math
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dreamlab_b200 import lib, synthetic as syn
from dreamlab_b200.engine import VAEDecoderB200

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
size = int(sys.argv[2]) if len(sys.argv) > 2 else 512
S, d = (size // 8) ** 2, 512
reps = int(os.environ.get("REPS", "5"))


def timed(fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


qk = torch.randn(B * S, 2 * d, device="cuda").bfloat16()
v = torch.randn(B * S, d, device="cuda").bfloat16()
out = torch.empty(B * S, d, device="cuda", dtype=torch.bfloat16)
t = timed(lambda: lib.attention_wide(qk, qk[:, d:], v, out, batch=B, sq=S, skv=S, d=d, ldq=2 * d, ldk=2 * d, ldv=d,
                                     ldo=d, scale=1 / math.sqrt(d)))
print(f"attention_wide B={B} S={S} d={d}: {t * 1e3:8.1f} us  {4.0 * B * S * S * d / t / 1e9:7.1f} TFLOP/s (useful)", flush=True)

nb = min(B, 2)
ref = torch.softmax(qk[:nb * S, :d].float().view(nb, S, d) @ qk[:nb * S, d:].float().view(nb, S, d).transpose(1, 2) / math.sqrt(d), -1) \
    @ v[:nb * S].float().view(nb, S, d)
got = out[:nb * S].float().view(nb, S, d)
print(f"  rel err vs fp32 softmax(QK^T)V: {float((got - ref).norm() / ref.norm()):.3e}", flush=True)
del ref, got

vcfg = syn.sd_vae_cfg()
vae = VAEDecoderB200(syn.random_state_dict(syn.vae_decoder_shapes(vcfg), 1), vcfg, "cuda:0")
lat = torch.randn(B, size // 8, size // 8, 4, device="cuda")
img = {}
for flash in (False, True, False, True):
    vae.flash_attention = flash
    t = timed(lambda: vae.decode(lat))
    img[flash] = vae.decode(lat).clone()
    print(f"vae.decode B={B} {size}x{size} flash={int(flash)}: {t:8.3f} ms", flush=True)
dd = (img[True].int() - img[False].int()).abs()
print(f"flash vs unfused u8: max {int(dd.max())}  frac != 0 {float((dd > 0).float().mean()):.3e}  frac > 1 {float((dd > 1).float().mean()):.3e}")
