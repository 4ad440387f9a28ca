"""world_size-2 gloo tests (CPU) of the host-side logic of the data-parallel path: shard-invariant transform
draws, CSR sharding, and the packed sum all-reduce that makes every rank apply the same update."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mladversarialobjectdetection_b200 import synth
from mladversarialobjectdetection_b200.ragged import RaggedBoxes
from mladversarialobjectdetection_b200.sampler import TransformSampler


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    B, per = 6, 3
    bt = synth.make_batch(B, 64, 64, seed=5, max_boxes=3)
    full = RaggedBoxes(torch.from_numpy(bt.boxes), torch.from_numpy(bt.offsets))
    mine = full.slice_rows(rank * per, (rank + 1) * per)
    from mladversarialobjectdetection_b200.attacker import resolve_first_image
    assert resolve_first_image(None, per) == rank * per and resolve_first_image(7, per) == 7   # the layers shard by rank on their own
    smp = TransformSampler(seed=3)
    params = smp.box_params(4, rank * per, mine.row_splits, mine.values.shape[0])
    wb = smp.print_wb(4, rank * per, per, "cpu")
    # gather the shards and compare with the single-rank draw
    all_wb = [torch.empty_like(wb) for _ in range(world)]
    dist.all_gather(all_wb, wb)
    counts = torch.tensor([params.shape[0]])
    all_counts = [torch.empty_like(counts) for _ in range(world)]
    dist.all_gather(all_counts, counts)
    whole_p = smp.box_params(4, 0, full.row_splits, full.values.shape[0])
    whole_wb = smp.print_wb(4, 0, B, "cpu")
    lo = sum(int(c) for c in all_counts[:rank])
    ok = torch.equal(torch.cat(all_wb), whole_wb) and torch.equal(params, whole_p[lo:lo + params.shape[0]])
    # packed gradient buffer: [dpatch | dscale | loss | sum M | sum M^2], SUM all-reduce (loss is a sum over images)
    g = torch.full((3 * 8 * 8 + 4,), float(rank + 1))
    dist.all_reduce(g)
    ok = ok and bool((g == 3.0).all())
    torch.save(dict(ok=ok, g=g), os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


def test_two_rank_gloo_sharding(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = torch.load(tmp_path / "r0.pt"), torch.load(tmp_path / "r1.pt")
    assert r0["ok"] and r1["ok"]
    assert torch.equal(r0["g"], r1["g"])


def test_slice_rows_csr():
    bt = synth.make_batch(5, 64, 64, seed=6, max_boxes=4)
    full = RaggedBoxes(torch.from_numpy(bt.boxes), torch.from_numpy(bt.offsets))
    part = full.slice_rows(2, 5)
    assert part.nrows() == 3 and int(part.row_splits[0]) == 0
    rows = full.to_rows()
    for a, b in zip(part.to_rows(), rows[2:5]):
        np.testing.assert_array_equal(a, b)
