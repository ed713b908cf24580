"""world_size=2 over gloo on CPU: the N>1 host path (shard plan + the one all-reduce)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import smooth_field


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, axis, out_dir):
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "tests"))
    from oracle import bitinfo_oracle as bo
    from xbitinfo_b200 import distributed

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(21)  # same global array on every rank
        x = smooth_field(rng, (6, 9, 16), "float32", axis=axis, scale=0.3)
        plan = distributed.plan_shard(x.shape, axis, world, rank)
        sl = [slice(None)] * x.ndim
        sl[plan.split_axis] = plan.read_slice
        slab = x[tuple(sl)]
        counts = bo.bitinformation(slab, axis)[0] if slab.shape[axis] >= 2 else np.zeros((32, 2, 2), dtype=np.int64)
        t = torch.from_numpy(counts.copy())
        distributed.allreduce_counts(t)  # the path's only collective
        assert distributed.world() == (rank, world)
        ref = bo.bitinformation(x, axis)[0]
        np.save(os.path.join(out_dir, f"ok_{axis}_{rank}.npy"), np.array([int(np.array_equal(t.numpy(), ref))]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("axis", [0, 2])
def test_two_ranks_gloo(tmp_path, axis):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), axis, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert np.load(tmp_path / f"ok_{axis}_{r}.npy")[0] == 1
