"""world_size=2 gloo tests (CPU) of the frame-parallel host logic used by bench.py --gpus N."""
import os
import socket
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nunif_b200 import parallel


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_frames, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = parallel.shard_frames(n_frames, rank, world)
        everyone = parallel.gather_frame_order(mine, world)
        # weight "blob": rank 0 holds the real bytes, the others garbage -> identical after the broadcast
        g = torch.Generator().manual_seed(7)
        blob = torch.randint(0, 255, (4099,), dtype=torch.uint8, generator=g) if rank == 0 else torch.zeros(4099, dtype=torch.uint8)
        parallel.broadcast_blob(blob, src=0)
        ref = torch.randint(0, 255, (4099,), dtype=torch.uint8, generator=torch.Generator().manual_seed(7))
        ms = parallel.max_over_ranks([10.0 + rank, 5.0 - rank])
        q.put((rank, mine, everyone, bool(torch.equal(blob, ref)), ms))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_broadcast_and_timing():
    world, n_frames = 2, 7
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_frames, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == [0, 2, 4, 6] and res[1][1] == [1, 3, 5]
    for _, _, everyone, same, ms in res:
        assert sorted(i for part in everyone for i in part) == list(range(n_frames))   # each frame exactly once
        assert same                                                                   # blob identical on all ranks
        assert ms == [11.0, 5.0]                                                      # max over ranks


def test_shard_edge_cases():
    assert parallel.shard_frames(0, 0, 2) == []
    assert parallel.shard_frames(1, 1, 2) == []
    assert parallel.shard_frames(5, 0, 1) == [0, 1, 2, 3, 4]
    assert [parallel.owner_of(i, 4) for i in range(6)] == [0, 1, 2, 3, 0, 1]
    import pytest
    with pytest.raises(ValueError):
        parallel.shard_frames(4, 2, 2)
    assert parallel.max_over_ranks([1.5]) == [1.5]
