"""Multi-GPU batches: one process per GPU, images sharded across ranks, no data-path collective.

JPEG XL images of a batch are independent, so a batch partitions over the ranks of a torch.distributed
job and every rank decodes its shard on its own GPU through the C ABI.  The only communication is control-plane (the (index -> rank) assignment is computed identically on every rank from the
compressed sizes) plus, only when the caller asks for the pixels in one place, a gather of the decoded images:
decode_batch_sharded_device leaves every image in HBM and moves it to the destination rank's GPU with NCCL point-to-point
transfers (device buffers over NVLink / NVSwitch); decode_batch_sharded is the host-array variant (gloo in the CPU tests).
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


class _DevView:
    """__cuda_array_interface__ over a decoded image left in HBM by the C ABI (jxlb_image.data on device >= 0)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


def device_tensor(bitmap):
    """uint8 CUDA tensor [height, stride] aliasing a Bitmap decoded with output_device >= 0 (zero-copy; keep the Bitmap alive)."""
    import torch
    if bitmap.device is None or bitmap.device < 0 or not bitmap.device_ptr:
        raise ValueError("the bitmap is not in device memory")
    n = bitmap.stride * bitmap.height
    with torch.cuda.device(bitmap.device):
        t = torch.as_tensor(_DevView(bitmap.device_ptr, n), device="cuda:%d" % bitmap.device)
    return t.view(bitmap.height, bitmap.stride)


def decode_batch_sharded_device(datas, device, gather_to=None, group=None, **decode_kwargs):
    """GPU data plane of a sharded batch: every rank decodes its shard with the pixels LEFT IN HBM (output_device = its
    GPU), and -- when gather_to is given -- the decoded images travel to that rank's GPU with NCCL point-to-point
    transfers (device buffers, NVLink / NVSwitch; images differ in size, so one send / recv per image batched with
    batch_isend_irecv rather than an equal-sized all_gather).  No image ever touches host memory.

    Returns {index: uint8 CUDA tensor [height, stride_bytes]} (this rank's shard; the whole batch on rank gather_to) and the
    list of (index, width, height, stride, config) of everything held.  Requires an initialised NCCL process group."""
    import torch
    import torch.distributed as dist
    from . import api
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    parts = partition([len(d) for d in datas], world)
    mine = parts[rank]
    bitmaps = api.decode_batch([datas[i] for i in mine], device=device, output_device=device, keep_native=True, **decode_kwargs) if mine else []
    held = {}
    meta = []
    for i, b in zip(mine, bitmaps):
        # own copy in torch's allocator, then hand the library's buffer back (the result buffers are cudaMalloc'ed per image)
        t = device_tensor(b).clone()
        b.free()
        held[i] = t
        meta.append((i, b.width, b.height, b.stride, b.config))
    if gather_to is None or world == 1:
        return held, meta
    metas = [None] * world
    dist.all_gather_object(metas, meta, group=group)
    ops = []
    if rank == gather_to:
        for r, m in enumerate(metas):
            if r == rank:
                continue
            for (i, w, h, stride, cfg) in m:
                t = torch.empty((h, stride), dtype=torch.uint8, device="cuda:%d" % device)
                held[i] = t
                ops.append(dist.P2POp(dist.irecv, t, r, group=group))
        meta = [x for m in metas for x in m]
    else:
        for i in mine:
            ops.append(dist.P2POp(dist.isend, held[i], gather_to, group=group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    torch.cuda.synchronize(device)
    return held, meta
