"""One eager pass of the bench workload between cudaProfilerStart/Stop, for ncu:
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/launches.csv python tools/one_pass.py [batch] [size] [lcm_steps]
Only the ~1550 launches of that pass are profiled (warm-up and setup are skipped)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dreamlab_b200 import synthetic as syn
from dreamlab_b200.engine import LCMPipelineB200

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
size = int(sys.argv[2]) if len(sys.argv) > 2 else 512
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
ucfg, vcfg = syn.sd15_lcm_unet_cfg(), syn.sd_vae_cfg()
pipe = LCMPipelineB200(syn.random_state_dict(syn.unet_shapes(ucfg), 0), ucfg,
                       syn.random_state_dict(syn.vae_decoder_shapes(vcfg), 1), vcfg, "cuda:0")
pe, lat, noise = syn.synthetic_inputs(B, size, size, steps)
pe, lat, noise = pe.cuda(), lat.cuda(), noise.cuda()
for _ in range(2):
    pipe.generate(pe, lat, noise, steps, 1.0)
torch.cuda.synchronize()
torch.cuda.profiler.start()
pipe.generate(pe, lat, noise, steps, 1.0)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("one pass done")
