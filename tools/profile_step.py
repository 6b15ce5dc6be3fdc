"""Developer tool: torch.profiler table of the KD training step (which kernels the step spends time in)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "lightweight-multi-modal-scene-understanding-via-knowledge-distillation_b200"))
import argparse, torch
import bench
ap = argparse.ArgumentParser(); ap.add_argument("--batch", type=int, default=32); ap.add_argument("--points", type=int, default=170000)
ap.add_argument("--fp32", action="store_true"); ap.add_argument("--out", default="gpurun_out/profile_step.txt"); a = ap.parse_args()
from src.data_loading.synthetic_frames import make_frames
dev = torch.device("cuda", 0)
tr = bench.build_models(dev, a.fp32)
b = make_frames(a.batch, a.points, seed=0, device=dev)
for _ in range(3):
    tr.training_step(b["image"], b["points"], b["segmentation"])
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    for _ in range(3):
        tr.training_step(b["image"], b["points"], b["segmentation"])
    torch.cuda.synchronize()
txt = prof.key_averages().table(sort_by="cuda_time_total", row_limit=140, max_name_column_width=90)
# the library's element-wise traffic by operator and input shape (what is left to fuse away)
small = [e for e in prof.key_averages(group_by_input_shape=True)
         if e.key in ("aten::copy_", "aten::add", "aten::add_", "aten::mul", "aten::fill_", "aten::zero_", "aten::mm", "aten::cat",
                      "aten::convolution_backward", "aten::cudnn_convolution", "aten::sum", "aten::div", "aten::sub")]
small.sort(key=lambda e: -e.device_time_total)
txt += "\n\nlibrary operators by input shape (3 steps): device us total, calls, key, shapes\n"
for e in small[:70]:
    txt += f"{e.device_time_total:10.1f} {e.count:5d} {e.key:28s} {str(e.input_shapes)[:150]}\n"
open(a.out, "w").write(txt)
print(txt[:200])
print("peak mem GB", torch.cuda.max_memory_allocated() / 2**30)
