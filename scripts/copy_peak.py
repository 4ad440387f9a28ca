"""What a plain device copy of the config-2 image batch (201 MB) achieves on this GPU: the practical ceiling of the image pass."""
import torch
for n_img in (64, 256, 680):
    a = torch.rand(n_img, 512, 512, 3, device="cuda")
    b = torch.empty_like(a)
    for _ in range(5):
        b.copy_(a)
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            b.copy_(a)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 20)
    nbytes = a.numel() * 4 * 2
    print(f"{n_img} images: {nbytes / 1e6:.0f} MB read+write in {best * 1e3:.1f} us = {nbytes / best / 1e6:.0f} GB/s")
