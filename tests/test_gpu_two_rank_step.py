"""Two ranks (torchrun-style launch, NCCL) through `PatchAttacker.train_step` itself: after 2 steps the patch, the scale
and the loss on both ranks equal those of ONE rank stepping on the concatenated batch (the loss is a sum over images
and the transform draws hash the global image index, so the sharded step is the single-GPU step up to float32
summation order).  Skipped with fewer than 2 GPUs."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, json
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["EOT_ROOT"])
from mladversarialobjectdetection_b200 import synth, victim
from mladversarialobjectdetection_b200.attacker import PatchAttacker
from mladversarialobjectdetection_b200.ragged import RaggedBoxes
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
B, H, P = 4, 256, 32                      # global batch
per = B // world
bt = synth.make_batch(B, H, H, seed=17, max_boxes=3, min_boxes=1)
full = RaggedBoxes(torch.from_numpy(bt.boxes).to(dev), torch.from_numpy(bt.offsets).to(dev))
mine = full.slice_rows(rank * per, (rank + 1) * per)
images = torch.from_numpy(bt.images[rank * per:(rank + 1) * per]).to(dev)
torch.manual_seed(0)
model = victim.get_victim_model("efficientdet-d0", device=dev, image_size=H)
att = PatchAttacker(model, patch_size=P, device=dev, seed=3)     # first_image is derived from the rank
att.compile(learning_rate=1e-2)
losses = []
for _ in range(2):
    m = att.train_step(images, boxes=mine)                         # global batch inferred: B_local * world
    losses.append(float(m["loss"]))
torch.cuda.synchronize()
np.savez(os.environ["EOT_OUT"] + f".w{world}.r{rank}.npz", patch=att._patch.cpu().numpy(), scale=float(att._scale_regressor),
         losses=np.asarray(losses), mean=float(m["mean_max_score"]))
if world > 1:
    dist.destroy_process_group()
'''


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_train_step_equals_single_rank(tmp_path):
    import numpy as np
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    out = str(tmp_path / "res")
    env = dict(os.environ, EOT_ROOT=ROOT, EOT_OUT=out)
    subprocess.run([sys.executable, str(script)], env=dict(env, RANK="0", WORLD_SIZE="1"), check=True, timeout=600)
    subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                    "127.0.0.1", "--master-port", str(_free_port()), str(script)], env=env, check=True, timeout=600)
    one = np.load(out + ".w1.r0.npz")
    r0, r1 = np.load(out + ".w2.r0.npz"), np.load(out + ".w2.r1.npz")
    np.testing.assert_array_equal(r0["patch"], r1["patch"])            # every rank applies the same update
    assert float(r0["scale"]) == float(r1["scale"])
    np.testing.assert_allclose(r0["patch"], one["patch"], rtol=0, atol=1e-6)
    assert abs(float(r0["scale"]) - float(one["scale"])) <= 1e-6
    np.testing.assert_allclose(r0["losses"], one["losses"], rtol=1e-5, atol=1e-6)
    assert abs(float(r0["mean"]) - float(one["mean"])) <= 1e-6
