"""N>1 on real GPUs (skipped on a single-GPU box): per-rank K1 + one all-reduce -- NCCL, and the push over
NVLink peer memory inside the finish kernel -- must give exactly the 1-GPU / oracle counts."""
import os
import socket

import numpy as np
import pytest
import torch

from conftest import smooth_field

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import sys

    import torch.distributed as dist

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "tests"))
    from oracle import bitinfo_oracle as bo
    from xbitinfo_b200 import distributed

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    ok = 1
    try:
        rng = np.random.default_rng(33)
        x = smooth_field(rng, (8, 120, 256), "float32", axis=-1, scale=0.3)
        x[:, rng.random((120, 256)) < 0.2] = np.nan
        for axis, mv in ((2, None), (1, float("nan")), (0, None)):
            plan = distributed.plan_shard(x.shape, axis, world, rank)
            sl = [slice(None)] * x.ndim
            sl[plan.split_axis] = plan.read_slice
            slab = torch.from_numpy(np.ascontiguousarray(x[tuple(sl)])).cuda()
            mi, counts = distributed.sharded_bitinformation(slab, axis, masked_value=mv, return_counts=True)
            ref_counts, _, ref_mi = bo.bitinformation(x, axis, masked_value=mv)
            ok &= int(np.array_equal(counts, ref_counts))
            ok &= int(np.allclose(mi, ref_mi, rtol=1e-12, atol=1e-12 * np.nanmax(np.abs(ref_mi))))
            # host-buffer route of the same call
            mi2, counts2 = distributed.sharded_bitinformation(np.ascontiguousarray(x[tuple(sl)]), axis, masked_value=mv,
                                                              return_counts=True, device=rank)
            ok &= int(np.array_equal(counts2, ref_counts))
        # the all-reduce inside the finish kernel (NVLink peer memory, CUDA IPC): same global table on every rank
        from xbitinfo_b200.device import PairCounter

        peers = distributed.PeerExchange(torch.device("cuda", rank))
        try:
            for axis, mv in ((2, None), (1, float("nan")), (0, None), (2, float("nan"))):
                plan = distributed.plan_shard(x.shape, axis, world, rank)
                sl = [slice(None)] * x.ndim
                sl[plan.split_axis] = plan.read_slice
                slab = torch.from_numpy(np.ascontiguousarray(x[tuple(sl)])).cuda()
                pc = PairCounter(slab.dtype, slab.device)
                mi = pc.bitinformation(slab, axis, masked_value=mv, peers=peers).cpu().numpy()
                ref_counts, _, ref_mi = bo.bitinformation(x, axis, masked_value=mv)
                ok &= int(np.array_equal(pc.counts.cpu().numpy(), ref_counts))
                ok &= int(np.allclose(mi, ref_mi, rtol=1e-12, atol=1e-12 * np.nanmax(np.abs(ref_mi))))
            ok &= int(peers.failed_epoch() == 0)
        finally:
            peers.close()
        np.save(os.path.join(out_dir, f"ok_{rank}.npy"), np.array([ok]))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least 2 GPUs")
def test_two_ranks_nccl(tmp_path):
    import torch.multiprocessing as mp

    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert np.load(tmp_path / f"ok_{r}.npy")[0] == 1
