"""2-rank NCCL test of the sharded GPU data plane (jxl_coder_b200.shard.decode_batch_sharded_device): each rank decodes its
shard with the pixels left in HBM and rank 0 gathers every image over NCCL; the gathered pictures equal a plain
single-process decode.  Needs 2 GPUs (skipped otherwise); the ranks are spawned here, rendezvous on 127.0.0.1."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import cases
    import jxl_coder_b200 as J
    from jxl_coder_b200 import shard
    datas = [cases.get(n) for n in cases.SMALL[:7]]
    held, meta = shard.decode_batch_sharded_device(datas, device=rank, gather_to=0, config=2)
    out = {}
    if rank == 0:
        for (i, w, h, stride, cfg) in meta:
            out[i] = held[i].cpu().numpy()[:, : w * 4].tobytes()
    q.put((rank, sorted(held.keys()), out))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_decode_in_hbm_and_gather_over_nccl():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    import cases
    import jxl_coder_b200 as J
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
        rank, keys, out = q.get(timeout=300)
        got[rank] = (keys, out)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    datas = [cases.get(n) for n in cases.SMALL[:7]]
    assert got[0][0] == list(range(len(datas)))          # rank 0 holds the whole batch
    assert 0 < len(got[1][0]) < len(datas)               # rank 1 only its shard
    want = J.decode_batch(datas, config=2)
    for i, b in enumerate(want):
        assert got[0][1][i] == b.pixels[:, : b.width * 4].tobytes(), i


def test_device_tensor_aliases_the_decoded_image():
    import torch
    import cases
    import jxl_coder_b200 as J
    from jxl_coder_b200 import shard
    data = cases.get(cases.SMALL[0])
    host = J.decode_batch([data], config=2)[0]
    dev = J.decode_batch([data], config=2, output_device=torch.cuda.current_device(), keep_native=True)[0]
    t = shard.device_tensor(dev)
    assert t.is_cuda and tuple(t.shape) == (dev.height, dev.stride)
    assert (t.cpu().numpy() == host.pixels).all()
    dev.free()
