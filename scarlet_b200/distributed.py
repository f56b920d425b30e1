"""Multi-GPU: independent scenes sharded across ranks, one process and one plan per GPU.

Every blend is a closed optimisation problem (scarlet/blend.py:57-83; the reference loops over blends at user level,
testing/api.py:216-226), so the fitting loop needs NO collective: each rank fits a contiguous block of scenes.  The
only exchange is the gather of the packed fitted parameters after the loop (``torch.distributed``: NCCL over NVLink
for device buffers, gloo for host arrays in the CPU tests).
"""
import numpy as np


def shard_bounds(n_items, world_size):
    """Contiguous, balanced blocks: -> list of (start, stop) per rank; the first ``n_items % world_size`` ranks get one more."""
    base, extra = divmod(int(n_items), int(world_size))
    bounds, start = [], 0
    for r in range(world_size):
        stop = start + base + (1 if r < extra else 0)
        bounds.append((start, stop))
        start = stop
    return bounds


def shard(items, rank, world_size):
    a, b = shard_bounds(len(items), world_size)[rank]
    return items[a:b]


def _world():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist, dist.get_rank(), dist.get_world_size()
    return None, 0, 1


def all_gather_ragged(tensor):
    """All-gather 1-D tensors whose lengths differ between ranks -> list of per-rank tensors (on every rank).
    Works for CUDA tensors under NCCL and CPU tensors under gloo."""
    import torch
    dist, rank, world = _world()
    if world == 1:
        return [tensor]
    n = torch.tensor([tensor.numel()], dtype=torch.int64, device=tensor.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(s.item()) for s in sizes]
    cap = max(sizes)
    padded = torch.zeros(cap, dtype=tensor.dtype, device=tensor.device)
    padded[:tensor.numel()] = tensor.reshape(-1)
    out = torch.empty(world * cap, dtype=tensor.dtype, device=tensor.device)
    dist.all_gather_into_tensor(out, padded)
    return [out[r * cap:r * cap + sizes[r]] for r in range(world)]


class _DevArray:
    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def gather_device_parameters(plan, device_index):
    """NCCL gather of the packed fitted spectra and morphologies straight from the plan's device buffers.
    -> (list of per-rank sed tensors, list of per-rank morph tensors, bytes received)."""
    import torch
    dp = plan.device_params()
    dev = "cuda:%d" % device_index
    sed = torch.as_tensor(_DevArray(dp["sed"], dp["n_sed"], "<f8"), device=dev)
    morph = torch.as_tensor(_DevArray(dp["morph"], dp["n_morph"], "<f4" if dp["elem_bytes"] == 4 else "<f8"), device=dev)
    seds, morphs = all_gather_ragged(sed), all_gather_ragged(morph)
    torch.cuda.synchronize()
    nbytes = sum(t.numel() * t.element_size() for t in seds + morphs)
    return seds, morphs, nbytes


def gather_host_results(results):
    """Gather per-scene result records (any picklable objects) from all ranks, in global scene order."""
    dist, rank, world = _world()
    if world == 1:
        return list(results)
    parts = [None] * world
    dist.all_gather_object(parts, list(results))
    return [r for part in parts for r in part]


def fit_sharded(make_blend, scene_ids, max_iter=200, e_rel=1e-3, precision=32, device=None, **fit_kwargs):
    """Fit ``scene_ids`` split across the ranks of the current process group: rank r builds (``make_blend(scene_id)``)
    and fits only its block as one ``BlendBatch``; returns, on every rank, one record per scene in global order:
    ``dict(scene_id, n_iter, logL, sed=[...], morph=[...])``."""
    from .blend import BlendBatch
    dist, rank, world = _world()
    mine = shard(list(scene_ids), rank, world)
    records = []
    if mine:
        blends = [make_blend(sid) for sid in mine]
        batch = BlendBatch(blends, precision=precision, device=device)
        res = batch.fit(max_iter=max_iter, e_rel=e_rel, **fit_kwargs)
        for sid, b, (n, logL) in zip(mine, blends, res):
            records.append(dict(scene_id=sid, n_iter=int(n), logL=float(logL),
                                sed=[np.array(s.parameters[0]) for s in b.sources],
                                morph=[np.array(s.parameters[1]) for s in b.sources]))
        batch.close()
    return gather_host_results(records)
