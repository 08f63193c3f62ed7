"""Host-side data-parallel logic on CPU with the gloo backend (world_size 2): gradient buckets tile the arena in
backward order and the bucketed all-reduce + 1/world scaling reproduces DDP's gradient averaging."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from argus_b200.engine import all_reduce_bucket, gradient_buckets
        from argus_b200.models import NCameraCNN

        torch.manual_seed(0)
        model = NCameraCNN()  # layout only: no CUDA call is made
        ranges = model.stage_ranges()
        n = model.flat_params.numel()
        g = torch.Generator().manual_seed(100 + rank)
        flat = torch.randn(n, generator=g)
        local = flat.clone()
        works = [all_reduce_bucket(b, async_op=True) for b in gradient_buckets(flat, ranges)]
        for w in works:
            w.wait()
        # reference: plain all-reduce of the whole arena
        whole = local.clone()
        dist.all_reduce(whole)
        ok = torch.equal(flat, whole)
        gathered = [torch.zeros(4) for _ in range(world)]
        dist.all_gather(gathered, flat[:4] / world)
        out[rank] = (ok, ranges, [g_.tolist() for g_ in gathered])
    finally:
        dist.destroy_process_group()


def test_bucketed_allreduce_world2():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert all(out[r][0] for r in range(world))
    assert out[0][1] == out[1][1]
    assert out[0][2] == out[1][2]  # both ranks hold identical averaged gradients


def test_stage_ranges_tile_the_arena():
    from argus_b200.models import NCameraCNN

    model = NCameraCNN()
    ranges = model.stage_ranges()
    n = model.flat_params.numel()
    # stage 0 is the tail of the network (head, fc, layer4): buckets are issued from the back of the arena
    assert ranges[0][1] == n and ranges[3][0] == 0
    for k in range(3):
        assert ranges[k][0] == ranges[k + 1][1]
    names = [name for name, *_ in model._param_infos]
    offs = {name: off for name, off, *_ in model._param_infos}
    assert ranges[0][0] == offs["resnet.layer4.0.conv1.weight"]
    assert ranges[1][0] == offs["resnet.layer3.0.conv1.weight"]
    assert ranges[2][0] == offs["resnet.layer2.0.conv1.weight"]
    assert names[0] == "resnet.conv1.weight" and names[-1] == "output_mlp.4.bias"
