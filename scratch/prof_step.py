"""Where does one eager step go?  usage: prof_step.py cfg5 [compact]  -> top CUDA kernels by total time (torch profiler)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from planedepth_b200.boundary import HotPath
from planedepth_b200.graph import make_step
from planedepth_b200.synthetic import make_batch, make_opt

name = sys.argv[1] if len(sys.argv) > 1 else "cfg5"
layout = sys.argv[2] if len(sys.argv) > 2 else "reference"
B, H, W, over, photometric, desc = bench.CONFIGS[name]
opt = make_opt(**over)
mnov = opt.self_distillation > 0
opt.self_distillation = 0.0
b = make_batch(B, H, W, opt, seed=1234, device="cuda", layout=layout, mask_novel=mnov)
hp = HotPath(opt, b.target_sides, pc_net=None, photometric=photometric, disp_rowwise=True)
step = make_step(hp, b.inputs, b.outputs, list(b.leaves.values()), b.attach)
for _ in range(3):
    step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(5):
        step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=70))
