"""Multi-process path on CPU: 2 ranks over gloo shard a batch with jxl_coder_b200.shard and gather to rank 0.  The decode
function is a CPU stand-in (no GPU here); the sharding, ordering and gather logic is what is under test."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from jxl_coder_b200 import shard  # noqa: E402


def _fake_decode(datas):
    # deterministic "image" per input: shape depends on the length, content on the bytes
    out = []
    for d in datas:
        h = 2 + len(d) % 5
        a = np.frombuffer((d * (h * 3 * 4 // max(1, len(d)) + 1))[: h * 3 * 4], dtype=np.uint8).reshape(h, 3, 4).copy()
        out.append(a)
    return out


def _inputs():
    rng = np.random.default_rng(7)
    return [bytes(rng.integers(0, 256, size=int(n), dtype=np.uint8)) for n in (900, 10, 450, 451, 300, 20, 700)]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    datas = _inputs()
    res = shard.decode_batch_sharded(datas, _fake_decode, gather_to=0)
    idx, _ = shard.shard_for_rank(datas, world, rank)
    q.put((rank, sorted(res.keys()), idx, {k: v.tobytes() for k, v in res.items()}))
    dist.barrier()
    dist.destroy_process_group()


def test_partition_is_balanced_and_deterministic():
    sizes = [900, 10, 450, 451, 300, 20, 700]
    p2 = shard.partition(sizes, 2)
    assert sorted(i for part in p2 for i in part) == list(range(7))
    loads = [sum(sizes[i] for i in part) for part in p2]
    assert abs(loads[0] - loads[1]) <= 100
    assert shard.partition(sizes, 2) == p2
    assert shard.partition(sizes, 1) == [list(range(7))]
    assert shard.partition([], 4) == [[], [], [], []]
    p8 = shard.partition(sizes, 8)  # more ranks than images: some ranks idle, nothing lost
    assert sorted(i for part in p8 for i in part) == list(range(7))


def test_two_ranks_shard_and_gather_over_gloo():
    world = 2
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(world):
        rank, keys, idx, blobs = q.get(timeout=120)
        got[rank] = (keys, idx, blobs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    datas = _inputs()
    want = {i: a.tobytes() for i, a in enumerate(_fake_decode(datas))}
    # rank 0 holds every image after the gather, rank 1 only its own shard; together the shards cover the batch once
    assert got[0][0] == list(range(len(datas)))
    assert got[0][2] == want
    assert got[1][0] == got[1][1]
    assert sorted(got[0][1] + got[1][1]) == list(range(len(datas)))
    for i in got[1][0]:
        assert got[1][2][i] == want[i]
