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
        from side_b200.engine import GradientAllReducer, allreduce_gradients, gather_detections, shard_range
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
        # a rank whose batch skipped a branch has NO gradient for that branch's parameters (stereo_network.forward without
        # boxes): buckets must not depend on which grads exist, the missing ones count as zeros
        skip = torch.nn.Linear(50, 50)
        params = list(lin.parameters()) + list(skip.parameters()) + list(big.parameters())
        for p in params:
            p.grad = torch.full_like(p, float(rank + 1))
        if rank == 1:
            for p in skip.parameters():
                p.grad = None
        allreduce_gradients(params, bucket_bytes=100 * 1024)
        ok = ok and all(torch.allclose(p.grad, torch.full_like(p, 0.5)) for p in skip.parameters())
        ok = ok and all(torch.allclose(p.grad, torch.full_like(p, 1.5)) for p in lin.parameters())
        # overlapped reducer: hooks fire during backward, fixed bucket order, the unused branch on rank 1 still reduces
        torch.manual_seed(7)
        net = torch.nn.ModuleDict(dict(unused=torch.nn.Linear(64, 8), a=torch.nn.Linear(64, 64), b=torch.nn.Linear(64, 64)))
        red = GradientAllReducer(net.parameters(), bucket_bytes=8 * 1024)
        ok = ok and len(red.buckets) >= 3
        for step in range(2):
            red.zero_grad()
            x = torch.full((4, 64), 1.0 + rank)
            y = net["b"](torch.relu(net["a"](x))).sum()
            if rank == 0:
                y = y + net["unused"](x).sum()
            y.backward()
            early = red.finish()
            ok = ok and early >= 1                                  # at least one bucket went out while backward was running
            # reference: average of the two ranks' gradients computed locally
            ref = {}
            for r in range(world):
                net2 = torch.nn.ModuleDict(dict(unused=torch.nn.Linear(64, 8), a=torch.nn.Linear(64, 64), b=torch.nn.Linear(64, 64)))
                net2.load_state_dict(net.state_dict())
                x2 = torch.full((4, 64), 1.0 + r)
                y2 = net2["b"](torch.relu(net2["a"](x2))).sum()
                if r == 0:
                    y2 = y2 + net2["unused"](x2).sum()
                y2.backward()
                for n, p in net2.named_parameters():
                    ref[n] = ref.get(n, 0) + (p.grad if p.grad is not None else torch.zeros_like(p)) / world
            ok = ok and all(torch.allclose(p.grad, ref[n], atol=1e-5) for n, p in net.named_parameters())
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
