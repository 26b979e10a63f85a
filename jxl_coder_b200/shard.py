"""Multi-GPU batches: one process per GPU, images sharded across ranks, no data-path collective.

JPEG XL images of a batch are independent, so a batch partitions over the ranks of a torch.distributed
job and every rank decodes its shard on its own GPU through the C ABI.  The only communication is control-plane: the
(index -> rank) assignment is computed identically on every rank from the compressed sizes, and -- only when the caller
asks for the pixels in one place -- decoded images are gathered to a destination rank (NCCL for device tensors over
NVLink, gloo for host tensors in the CPU tests).
"""
import numpy as np


def partition(sizes, world_size):
    """Greedy longest-processing-time assignment of items (compressed byte counts) to ranks.  Deterministic: every rank
    computes the same answer from the same sizes.  Returns a list of index lists, one per rank, each in input order."""
    world_size = max(1, int(world_size))
    order = sorted(range(len(sizes)), key=lambda i: (-int(sizes[i]), i))
    load = [0] * world_size
    count = [0] * world_size
    out = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], count[k], k))
        out[r].append(i)
        load[r] += int(sizes[i])
        count[r] += 1
    return [sorted(v) for v in out]


def shard_for_rank(datas, world_size, rank):
    """(indices, shard) of this rank."""
    idx = partition([len(d) for d in datas], world_size)[rank]
    return idx, [datas[i] for i in idx]


def decode_batch_sharded(datas, decode_fn, world_size=None, rank=None, gather_to=None, group=None):
    """Decodes `datas` (the SAME list on every rank) with each rank taking its shard.

    decode_fn(list_of_bytes) -> list of numpy arrays (e.g. lambda d: [b.as_array() for b in decode_batch(d, ...)]).
    Returns {index: array} holding this rank's results; with gather_to = r, rank r receives every image (the other ranks
    keep only their own).  Uses torch.distributed only for the optional gather."""
    dist = None
    if world_size is None or rank is None or gather_to is not None:
        import torch.distributed as dist  # noqa: F811
    if world_size is None:
        world_size = dist.get_world_size(group) if dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank(group) if dist.is_initialized() else 0
    parts = partition([len(d) for d in datas], world_size)
    mine = parts[rank]
    arrays = decode_fn([datas[i] for i in mine]) if mine else []
    if len(arrays) != len(mine):
        raise RuntimeError("decode_fn returned %d results for %d inputs" % (len(arrays), len(mine)))
    result = {i: a for i, a in zip(mine, arrays)}
    if gather_to is None or world_size == 1:
        return result
    import torch
    # shapes / dtypes first (tiny), then one send/recv per image: images differ in size, so no equal-sized all_gather
    meta = [(i, tuple(result[i].shape), str(result[i].dtype)) for i in mine]
    metas = [None] * world_size
    dist.all_gather_object(metas, meta, group=group)
    if rank == gather_to:
        reqs = []
        for r, m in enumerate(metas):
            if r == rank:
                continue
            for (i, shape, dtype) in m:
                t = torch.empty(shape, dtype=getattr(torch, dtype))
                reqs.append((i, t, dist.irecv(t, src=r, group=group, tag=i)))
        for i, t, w in reqs:
            w.wait()
            result[i] = t.numpy()
    else:
        ws = [dist.isend(torch.from_numpy(np.ascontiguousarray(result[i])), dst=gather_to, group=group, tag=i) for i in mine]
        for w in ws:
            w.wait()
    return result
