"""CPU tests of the N > 1 host logic with the gloo backend (world_size 2): view sharding and the flat
gradient buffer all-reduce produce exactly the single-process sum."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from online_lang_splatting_b200.sharding import FlatGradBuffer, shard_views


def test_shard_views_partition():
    for n, w in ((12, 2), (12, 8), (5, 4), (64, 8), (3, 8)):
        seen = []
        for r in range(w):
            seen += shard_views(n, r, w)
        assert sorted(seen) == list(range(n))
        sizes = [len(shard_views(n, r, w)) for r in range(w)]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_views(4, 4, 4)


def test_flat_buffer_layout():
    fb = FlatGradBuffer(P=10, F=15, M=1, device="cpu")
    assert fb.floats_per_gaussian == 29 and fb.flat.numel() == 290  # 29 floats per Gaussian at deg 0, F=15
    for name, v in fb.views.items():
        v.fill_(1.0)
    assert float(fb.flat.sum()) == 290.0
    fb.views["language"][3, 7] = 5.0
    assert fb.flat[14 * 10 + 3 * 15 + 7] == 5.0
    assert all(v.is_contiguous() for v in fb.views.values())


def _view_grad(view: int, P: int):
    g = torch.Generator().manual_seed(100 + view)
    return {"means3D": torch.randn(P, 3, generator=g), "sh": torch.randn(P, 1, 3, generator=g),
            "opacity": torch.randn(P, 1, generator=g), "scales": torch.randn(P, 3, generator=g),
            "rotations": torch.randn(P, 4, generator=g), "language": torch.randn(P, 15, generator=g)}


def _worker(rank, world, port, n_views, P, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    fb = FlatGradBuffer(P=P, F=15, M=1, device="cpu")
    for v in shard_views(n_views, rank, world):       # each rank accumulates only its own views
        for name, t in _view_grad(v, P).items():
            fb.views[name] += t
    fb.all_reduce()
    q.put((rank, fb.flat.clone()))
    dist.barrier()
    dist.destroy_process_group()


def test_allreduce_equals_single_process_sum():
    world, n_views, P = 2, 5, 64
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_views, P, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = FlatGradBuffer(P=P, F=15, M=1, device="cpu")
    for v in range(n_views):
        for name, t in _view_grad(v, P).items():
            ref.views[name] += t
    for r in range(world):
        assert torch.allclose(got[r], ref.flat, atol=1e-5)
    assert torch.equal(got[0], got[1])  # identical reduced gradients on every rank -> identical optimiser steps


def test_balanced_views_evens_out_rank_loads():
    from online_lang_splatting_b200.sharding import balanced_views
    import random
    rnd = random.Random(0)
    costs = [1.0 + 0.7 * rnd.random() ** 3 for _ in range(64)]   # a few expensive views, like the synthetic keyframes
    blocks = [list(range(r * 8, r * 8 + 8)) for r in range(8)]
    a = balanced_views(costs, 8)
    assert sorted(v for x in a for v in x) == list(range(64)) and all(len(x) == 8 for x in a)
    load = lambda parts: [sum(costs[v] for v in x) for x in parts]
    assert max(load(a)) <= max(load(blocks)) and max(load(a)) - min(load(a)) < 0.05 * max(load(a))
    assert balanced_views([3.0, 1.0], 1) == [[0, 1]]
    with pytest.raises(ValueError):
        balanced_views([1.0] * 7, 2)
