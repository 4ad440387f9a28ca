"""One bench step (config 2) under torch.profiler: the kernels and CUDA runtime calls of a `PatchAttacker.train_step`
in issue order, with the launches outside the two CUDA-graph replays counted (how many eager
launches and host synchronisations the step has besides the victim).  Cheap (no ncu): python scripts/step_launch_list.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

from mladversarialobjectdetection_b200 import synth, victim
from mladversarialobjectdetection_b200.attacker import PatchAttacker
from mladversarialobjectdetection_b200.ragged import RaggedBoxes

B, H, P = 64, 512, 100
dev = torch.device("cuda", 0)
torch.backends.cudnn.benchmark = True
model = victim.get_victim_model("efficientdet-d0", device=dev, image_size=H)
att = PatchAttacker(model, patch_size=P, device=dev, seed=7, cuda_graphs=True, always_first_pass=True, box_capacity=B * 8)
att.compile(learning_rate=1e-2)
bt = synth.make_batch(B, H, H, max_boxes=8)
images = torch.from_numpy(bt.images).to(dev)
boxes = RaggedBoxes(torch.from_numpy(bt.boxes).to(dev), torch.from_numpy(bt.offsets).to(dev))
for _ in range(3):
    att.train_step(images, boxes=boxes)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    att.train_step(images, boxes=boxes)
    torch.cuda.synchronize()
ev = sorted(prof.events(), key=lambda e: e.time_range.start)
runtime = [e for e in ev if e.name.startswith("cuda") and e.device_type.name == "CPU"]
kernels = [e for e in ev if e.device_type.name == "CUDA"]
print("## CUDA runtime calls of one train_step, in issue order (name x consecutive count)")
last, n = None, 0
for e in runtime + [None]:
    name = e.name if e is not None else None
    if name == last:
        n += 1
        continue
    if last is not None:
        print(f"{n:5d} x {last}")
    last, n = name, 1
eager = sum(1 for e in runtime if e.name in ("cudaLaunchKernel", "cudaLaunchKernelExC", "cuLaunchKernel"))
graphs = sum(1 for e in runtime if e.name == "cudaGraphLaunch")
syncs = [e.name for e in runtime if "Synchronize" in e.name or e.name in ("cudaMemcpy",)]
print(f"\neager kernel launches: {eager}; graph replays: {graphs}; host synchronisations inside the step: {syncs[:-1] if syncs else []} "
      f"(the last one is the profiler's own torch.cuda.synchronize)")
print("\n## device kernels in execution order (libeotpatch kernels are eot::*)")
last, n, t = None, 0, 0.0
for e in kernels + [None]:
    name = e.name[:90] if e is not None else None
    if name == last:
        n += 1
        t += e.device_time if hasattr(e, "device_time") else e.cuda_time
        continue
    if last is not None:
        print(f"{n:5d} x {t:9.1f} us  {last}")
    if e is not None:
        last, n, t = name, 1, (e.device_time if hasattr(e, "device_time") else e.cuda_time)
