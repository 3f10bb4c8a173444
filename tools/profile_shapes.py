"""Per-shape efficiency table of one eager pass of the hot path (CUDA events around every
native launch).  Usage: python tools/profile_shapes.py [batch] [size] > profiles/....txt"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import json
import torch
from dreamlab_b200 import lib, synthetic as syn
from dreamlab_b200.engine import LCMPipelineB200

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
size = int(sys.argv[2]) if len(sys.argv) > 2 else 512
ucfg, vcfg = syn.sd15_lcm_unet_cfg(), syn.sd_vae_cfg()
pipe = LCMPipelineB200(syn.random_state_dict(syn.unet_shapes(ucfg), 0), ucfg,
                       syn.random_state_dict(syn.vae_decoder_shapes(vcfg), 1), vcfg, "cuda:0")
pe, lat, noise = syn.synthetic_inputs(B, size, size, 4)
pe, lat, noise = pe.cuda(), lat.cuda(), noise.cuda()
for _ in range(2):
    pipe.generate(pe, lat, noise, 4, 1.0)
torch.cuda.synchronize()
# park the GPU on a spin kernel while the CPU enqueues the whole pass, so the kernels then run
# back-to-back and the per-launch CUDA events measure device time, not CPU launch latency
torch.cuda._sleep(int(2.0e8))
lib.profile_begin(detail=True)
pipe.generate(pe, lat, noise, 4, 1.0)
prof = lib.profile_end()
peak = json.load(open("MEASURED_PEAKS.json")) if os.path.exists("MEASURED_PEAKS.json") else {}
tf, bw = peak.get("bf16_tflops_sustained", 1386.3), peak.get("hbm_gbs", 6535.7)
tot = sum(v["ms"] for v in prof.values())
print(f"batch {B} size {size}: total {tot:.2f} ms over {sum(v['n'] for v in prof.values())} launches")
print(f"{'kernel / shape':92s} {'n':>4s} {'ms':>8s} {'%':>5s} {'TFLOP/s':>8s} {'%pk':>5s} {'GB/s':>7s} {'%bw':>5s}")
for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
    t = v["flops"] / (v["ms"] / 1e3) / 1e12 if v["flops"] else 0
    g = v["bytes"] / (v["ms"] / 1e3) / 1e9 if v["bytes"] else 0
    print(f"{k[:92]:92s} {v['n']:4d} {v['ms']:8.3f} {100 * v['ms'] / tot:5.1f} {t:8.1f} {100 * t / tf:5.1f} {g:7.0f} {100 * g / bw:5.1f}")
