"""world_size-2 CPU (gloo) test of the N>1 host logic: rendezvous, image-index sharding, the single weight-blob
broadcast, and the host-side gather. No data-path collective exists to test beyond these."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from sdwebui_b200 import parallel as P
    from sdwebui_b200.rng import ImageRNG

    r, w, _ = P.init_from_env("gloo")
    seeds = list(range(1000, 1008))
    mine = P.shard(seeds, r, w)
    blob = torch.arange(1 << 12, dtype=torch.uint8) if r == 0 else torch.zeros(1 << 12, dtype=torch.uint8)
    P.broadcast_weight_blob(blob, src=0)
    ok_blob = bool(torch.equal(blob, torch.arange(1 << 12, dtype=torch.uint8)))
    noise = ImageRNG((4, 8, 8), mine, source="NV", device="cpu").next()
    gathered = P.gather_images(noise, w)
    if r == 0:
        full = torch.cat(gathered)
        ref = ImageRNG((4, 8, 8), seeds, source="NV", device="cpu").next()
        q.put((ok_blob, bool(torch.equal(full, ref)), mine))
    else:
        q.put((ok_blob, True, mine))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_broadcast_gather():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[0] and r[1] for r in res)
    assert sorted(sum((r[2] for r in res), [])) == list(range(1000, 1008))  # sharding invisible in the output
