"""Per-tile clock stamps of one attention CTA (debug aid): where does the time go?"""
import math, os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dreamlab_b200 import lib
l = lib.load()
l.dl_debug_attention_trace.argtypes = [C.c_void_p]
B, S, heads, d = 16, 4096, 8, 40
hs = 48
qkv = torch.randn(B * S, 3 * heads * hs, device="cuda").bfloat16()
out = torch.empty(B * S, heads * d, device="cuda", dtype=torch.bfloat16)
tr = torch.zeros(256, device="cuda", dtype=torch.int64)
def run():
    lib.attention(qkv, qkv[:, heads * hs:], qkv[:, 2 * heads * hs:], out, batch=B, sq=S, skv=S, heads=heads,
                  d=d, dh_stride=hs, ldq=3 * heads * hs, ldk=3 * heads * hs, ldv=3 * heads * hs, ldo=heads * d,
                  scale=1 / math.sqrt(d), v_ones=True)
run(); torch.cuda.synchronize()
l.dl_debug_attention_trace(tr.data_ptr())
run(); torch.cuda.synchronize()
l.dl_debug_attention_trace(None)
t = tr.cpu().view(16, 16)
t0 = int(t[0, 4])
names = ["mma:p_ready", "mma:pv_issued", "mma:s_taken", "mma:s_issued", "sm:s_full", "sm:ld_done", "sm:max_done",
         "sm:bar_done", "sm:exp_issued", "sm:st_done", "sm:arrived", "sm:looptop", "mma:k_full"]
print("tile " + " ".join(f"{n:>13s}" for n in names))
for j in range(12):
    print(f"{j:4d} " + " ".join(f"{int(t[j, k]) - t0:13d}" for k in range(13)))
