"""CPU suite, part 3: the N > 1 host logic over gloo with world_size 2 -- scene sharding (no collective on
the forward path), re-assembly of per-scene outputs, and the flat-bucket gradient all-reduce of the
training configuration."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from situation3d_b200.sharding import FlatGradAllReduce, gather_scene_outputs, scene_shard
        num_scenes = 7                                                  # ragged: 4 + 3
        mine = scene_shard(num_scenes, rank, world)
        assert mine == list(range(rank, num_scenes, world))
        # each rank "encodes" its own scenes (a stand-in for the backbone: per-scene function, no exchange)
        local = torch.stack([torch.full((3, 5), float(s)) + torch.arange(5.0) for s in mine])
        full = gather_scene_outputs(local, num_scenes, rank, world)
        want = torch.stack([torch.full((3, 5), float(s)) + torch.arange(5.0) for s in range(num_scenes)])
        assert torch.equal(full, want)

        # gradient all-reduce: one flat bucket, SUM then / world; parameters stay in sync
        torch.manual_seed(0)
        net = torch.nn.Sequential(torch.nn.Conv2d(4, 6, 1, bias=False), torch.nn.BatchNorm2d(6), torch.nn.ReLU(),
                                  torch.nn.Conv2d(6, 3, 1))
        bucket = FlatGradAllReduce(net)
        assert bucket.flat.numel() == sum(p.numel() for p in net.parameters())
        x = torch.randn(2, 4, 8, 2, generator=torch.Generator().manual_seed(100 + rank))
        bucket.zero()
        net(x).square().mean().backward()
        local_grads = [p.grad.clone() for p in net.parameters()]
        bucket.allreduce()
        gathered = [torch.zeros_like(bucket.flat) for _ in range(world)]
        flat_local = torch.cat([g.reshape(-1) for g in local_grads])
        dist.all_gather(gathered, flat_local)
        torch.testing.assert_close(bucket.flat, sum(gathered) / world)
        for p, v in zip(net.parameters(), bucket.views):
            assert p.grad.data_ptr() == v.data_ptr()                     # grads live in the bucket
        # chunked all-reduce started by hooks during backward gives the same averaged gradients
        net2 = torch.nn.Sequential(torch.nn.Conv2d(4, 6, 1, bias=False), torch.nn.BatchNorm2d(6), torch.nn.ReLU(),
                                   torch.nn.Conv2d(6, 3, 1))
        net2.load_state_dict(net.state_dict())
        b2 = FlatGradAllReduce(net2).enable_overlap(nchunks=2)
        assert len(b2.chunks) == 2 and b2.chunks[0][0] == 0 and b2.chunks[-1][1] == b2.flat.numel()
        b2.zero()
        net2(x).square().mean().backward()
        b2.finish()
        torch.testing.assert_close(b2.flat, bucket.flat)
        # BatchNorm statistics are NOT synchronised (the reference has no SyncBN): they differ across ranks
        stats = [torch.zeros(6) for _ in range(world)]
        dist.all_gather(stats, net[1].running_mean.clone())
        assert not torch.equal(stats[0], stats[1])
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_scene_sharding_and_grad_allreduce_world2():
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        assert dict(ret) == {0: "ok", 1: "ok"}


def test_scene_shard_covers_every_scene_once():
    from situation3d_b200.sharding import scene_shard
    for num, world in [(8, 8), (64, 8), (7, 2), (3, 4), (0, 2)]:
        got = sorted(s for r in range(world) for s in scene_shard(num, r, world))
        assert got == list(range(num))
