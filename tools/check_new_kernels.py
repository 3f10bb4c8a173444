"""Small invocations of every kernel added in round 2, each checked against a torch reference: a one-minute sanity
run (`python tools/check_new_kernels.py`), written to go under `compute-sanitizer --tool memcheck | synccheck` —
which is closed on this GPU pool ("runs under it have left GPUs needing a reset"), so the bounded mbarrier waits
(a protocol bug traps after 4 s instead of hanging), the device-side smem carve-up check of attention_pp and the
per-kernel parity tests are what stands in for it."""
import math
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import torch.nn.functional as F
from dreamlab_b200 import lib

lib.load()
DEV = "cuda"
g = torch.Generator().manual_seed(0)
rand = lambda *s, scale=1.0: (torch.randn(*s, generator=g) * scale).to(DEV)
bf = lambda t: t.to(torch.bfloat16)
rel = lambda a, b: ((a.float() - b.float()).abs().max() / b.float().abs().max()).item()


def attn(B, Sq, Skv, heads, d):
    hs = (d + 1 + 15) // 16 * 16
    q = torch.zeros(B * Sq, heads * hs, device=DEV, dtype=torch.bfloat16)
    k = torch.zeros(B * Skv, heads * hs, device=DEV, dtype=torch.bfloat16)
    v = torch.zeros(B * Skv, heads * hs, device=DEV, dtype=torch.bfloat16)
    qr, kr, vr = bf(rand(B, Sq, heads, d)), bf(rand(B, Skv, heads, d)), bf(rand(B, Skv, heads, d))
    q.view(B, Sq, heads, hs)[..., :d] = qr
    k.view(B, Skv, heads, hs)[..., :d] = kr
    v.view(B, Skv, heads, hs)[..., :d] = vr
    v.view(B, Skv, heads, hs)[..., d] = 1.0
    out = torch.zeros(B * Sq, heads * d, device=DEV, dtype=torch.bfloat16)
    lib.attention(q, k, v, out, batch=B, sq=Sq, skv=Skv, heads=heads, d=d, dh_stride=hs, ldq=heads * hs, ldk=heads * hs,
                  ldv=heads * hs, ldo=heads * d, scale=1 / math.sqrt(d), v_ones=True)
    ref = F.scaled_dot_product_attention(qr.float().transpose(1, 2), kr.float().transpose(1, 2),
                                         vr.float().transpose(1, 2)).transpose(1, 2)
    torch.cuda.synchronize()
    return rel(out.view(B, Sq, heads, d), ref)


print("attention_x  (256 x 77, d 40):", attn(1, 256, 77, 2, 40))
print("attention_x  (300 x 100, d 80):", attn(1, 300, 100, 2, 80))
print("attention_pp (512 x 600, d 40):", attn(1, 512, 600, 2, 40))

# igemm with per-channel GroupNorm records + finalize + apply
B, H, W, C, N = 1, 32, 32, 64, 320
x, wt = bf(rand(B, H, W, C)), bf(rand(N, 9 * C, scale=(9 * C) ** -0.5))
out = torch.empty(B, H, W, N, device=DEV, dtype=torch.bfloat16)
part = torch.zeros(B, lib.igemm_tiles_per_image(H, W), N, 2, device=DEV)
lib.igemm(x, wt, out, nimg=B, h=H, w=W, taps=9, n=N, bias=rand(N), gn_partial=part, gn_cpg=1)
stats = torch.empty(B, 32, 2, device=DEV)
lib.groupnorm_finalize_channels(part, None, stats, 32, H * W * (N // 32))
y = torch.empty_like(out)
lib.groupnorm_apply(out, y, rand(N), rand(N), stats.unsqueeze(0), nimg=B, hw=H * W, groups=32, eps=1e-5, silu=True)
torch.cuda.synchronize()
o = out.float().view(B, H * W, 32, N // 32).permute(0, 2, 1, 3).reshape(B, 32, -1)
print("gn channel records: mean err", (stats[..., 0] - o.mean(-1)).abs().max().item())

# LayerNorm fold
from dreamlab_b200.weights import fold_layernorm
M, Cc, Nn = 256, 320, 384
xx, w1 = bf(rand(M, Cc)), bf(rand(Cc, Cc, scale=Cc ** -0.5))
h = torch.empty(M, Cc, device=DEV, dtype=torch.bfloat16)
rs = lib.igemm(xx, w1, h, nimg=1, h=1, w=M, taps=1, n=Cc, bias=rand(Cc), residual=bf(rand(M, Cc)), ldr=Cc, ldo=Cc, row_stats=True)
gam, bet, w2, b2 = rand(Cc) * 0.3 + 1, rand(Cc) * 0.3, rand(Nn, Cc, scale=Cc ** -0.5), rand(Nn)
wf, cs, bfold = fold_layernorm(w2, b2, gam, bet, DEV)
o2 = torch.empty(M, Nn, device=DEV, dtype=torch.bfloat16)
lib.igemm(h, wf, o2, nimg=1, h=1, w=M, taps=1, n=Nn, bias=bfold, ldo=Nn, ln=(rs, cs, 1e-5))
torch.cuda.synchronize()
print("layernorm fold:", rel(o2, F.layer_norm(h.float(), (Cc,), gam, bet, 1e-5) @ w2.t() + b2))

# conv_out as 1x1 GEMM + tap sum, device PNG, adaptive latent pooling
yv = rand(1, 24, 40, 32)
img = torch.empty(1, 24, 40, 3, device=DEV, dtype=torch.uint8)
lib.conv_tapsum(yv, rand(3), img)
png, size = lib.png_stored(img)
from oracle.png import png_stored
torch.cuda.synchronize()
print("png bytes equal oracle:", png[0, :size].cpu().numpy().tobytes() == png_stored(img[0].cpu().numpy()))
lat = rand(2, 75, 75, 4)
p8 = torch.empty(2, 4, 8, 8, device=DEV, dtype=torch.float16)
lib.latent_pool8(lat, p8)
torch.cuda.synchronize()
print("latent_pool8 adaptive:", (p8.float() - F.adaptive_avg_pool2d(lat.permute(0, 3, 1, 2), (8, 8))).abs().max().item())
print("done")
