"""CPU, world_size 2 over gloo: pair sharding, detection gather and gradient averaging (the N>1 host logic)."""
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


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from side_b200.engine import allreduce_gradients, gather_detections, shard_range
        lo, hi = shard_range(7, rank, world)
        B, K = hi - lo, 5
        dets = torch.full((B, K, 6), float(rank)); detr = dets + 0.5; info = torch.full((B, K, 10), 10.0 + rank)
        # all_gather needs equal shapes: pad the short shard like bench.py does for uneven splits
        Bmax = 4
        pad = lambda t: torch.cat([t, t.new_zeros((Bmax - t.shape[0],) + tuple(t.shape[1:]))], 0)
        a, b, c = gather_detections(pad(dets), pad(detr), pad(info))
        ok = len(a) == world and all(float(a[r][0, 0, 0]) == r and float(b[r][0, 0, 0]) == r + 0.5 and
                                     float(c[r][0, 0, 0]) == 10.0 + r and c[r].shape[-1] == 10 for r in range(world))
        lin = torch.nn.Linear(4, 3)
        big = torch.nn.Linear(300, 300)
        for p in list(lin.parameters()) + list(big.parameters()):
            p.grad = torch.full_like(p, float(rank + 1))
        nb = allreduce_gradients(list(lin.parameters()) + list(big.parameters()), bucket_bytes=100 * 1024)
        ok = ok and nb >= 2 and all(torch.allclose(p.grad, torch.full_like(p, 1.5)) for p in list(lin.parameters()) + list(big.parameters()))
        ret[rank] = (ok, (lo, hi))
    finally:
        dist.destroy_process_group()


def test_world2_gloo():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert ret[0][0] and ret[1][0]
    assert ret[0][1] == (0, 4) and ret[1][1] == (4, 7)


def test_shard_range_covers_everything():
    from side_b200.engine import shard_range
    for total in (1, 7, 32, 33):
        for world in (1, 2, 4, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
