"""CPU, world_size 2 over gloo: the host side of the multi-GPU path (scene sharding + gather of packed results).
The fitting loop itself has no collective (scenes are independent); what is exercised here is exactly the code the
N>1 bench and ``fit_sharded`` run around it."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT


def test_shard_bounds():
    from scarlet_b200.distributed import shard, shard_bounds
    assert shard_bounds(10, 4) == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert shard_bounds(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]
    assert shard_bounds(4096, 8) == [(512 * r, 512 * (r + 1)) for r in range(8)]
    items = list(range(11))
    assert sum((shard(items, r, 3) for r in range(3)), []) == items


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from scarlet_b200 import distributed as sd
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        scene_ids = list(range(7))          # ragged: rank 0 gets 4 scenes, rank 1 gets 3
        mine = sd.shard(scene_ids, rank, world)
        # packed "fitted parameters" of this rank: 12 sources x 5 bands per scene, value encodes (scene, index)
        packed = torch.tensor(np.concatenate([1000.0 * s + np.arange(60) for s in mine]))
        parts = sd.all_gather_ragged(packed)
        assert [p.numel() for p in parts] == [240, 180]
        full = torch.cat(parts).numpy()
        expect = np.concatenate([1000.0 * s + np.arange(60) for s in scene_ids])
        assert np.array_equal(full, expect)
        recs = sd.gather_host_results([dict(scene_id=s, n_iter=10 + s) for s in mine])
        assert [r["scene_id"] for r in recs] == scene_ids and recs[5]["n_iter"] == 15
        # max-over-ranks timing reduction used by bench.py
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert t.item() == world
        open(os.path.join(out_dir, "ok%d" % rank), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_gather_world_size_2_gloo(tmp_path):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert sorted(os.listdir(tmp_path)) == ["ok0", "ok1"]
