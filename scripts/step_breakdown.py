"""Where the attack step's time goes (config 2): CUDA-event timing of each stage + torch profiler top kernels."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mladversarialobjectdetection_b200 import ops, synth, victim
from mladversarialobjectdetection_b200.attacker import PatchAttacker
from mladversarialobjectdetection_b200.ragged import RaggedBoxes

dev = torch.device("cuda")
torch.backends.cudnn.benchmark = True
B, H, P = 64, 512, 100
model = victim.get_victim_model("efficientdet-d0", device=dev, image_size=H)
att = PatchAttacker(model, patch_size=P, device=dev, seed=7)
bt = synth.make_batch(B, H, H)
images = torch.from_numpy(bt.images).to(dev)
boxes = RaggedBoxes(torch.from_numpy(bt.boxes).to(dev), torch.from_numpy(bt.offsets).to(dev))


def timed(name, fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        r = fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name:28s} gpu {e0.elapsed_time(e1) / n:8.2f} ms   wall {(time.perf_counter() - t0) / n * 1e3:8.2f} ms")
    return r


def fwd_nograd():
    with torch.no_grad():
        return model(images)


def fwd_bwd():
    x = images.clone().requires_grad_(True)
    cls, box = model(x)
    torch.autograd.backward(cls, [torch.ones_like(c) for c in cls])
    return x.grad


timed("victim fwd (no grad)", fwd_nograd)
timed("victim fwd+bwd", fwd_bwd)
timed("first_pass (fwd+score+NMS)", lambda: att.first_pass(images))
timed("train_step", lambda: att.train_step(images, boxes=boxes))
att.always_first_pass = False
timed("train_step w/o first pass", lambda: att.train_step(images, boxes=boxes))
if "--prof" in sys.argv:
    from torch.profiler import profile, ProfilerActivity
    att.always_first_pass = True
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        att.train_step(images, boxes=boxes)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))
