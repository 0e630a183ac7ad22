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


# ---- densification side statistics: SUM of accum / denom, MAX of max_radii2D over the ranks --------------
def _view_stats(view: int, P: int):
    g = torch.Generator().manual_seed(500 + view)
    vis = torch.rand(P, generator=g) > 0.5
    radii = torch.where(vis, torch.randint(1, 40, (P,), generator=g), torch.zeros(P, dtype=torch.long)).float()
    gnorm = torch.where(vis, torch.rand(P, generator=g), torch.zeros(P))
    return vis, radii, gnorm


def _fold(st, view, P):
    vis, radii, gnorm = _view_stats(view, P)        # what densification.update_stats does on the GPU, in torch
    st.delta[:P] += gnorm
    st.delta[P:] += vis.float()
    torch.maximum(st.delta_max, radii, out=st.delta_max)


def _stats_worker(rank, world, port, n_views, P, q):
    from online_lang_splatting_b200.sharding import SideStats
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    st = SideStats(P, device="cpu")
    for step in range(2):
        st.begin_step()
        for v in shard_views(n_views, rank, world):
            _fold(st, v + 10 * step, P)
        st.all_reduce().apply()
    q.put((rank, st.xyz_gradient_accum.clone(), st.denom.clone(), st.max_radii2D.clone()))
    dist.barrier()
    dist.destroy_process_group()


def test_side_stats_reduce_equals_single_process():
    from online_lang_splatting_b200.sharding import SideStats
    world, n_views, P = 2, 6, 257
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_stats_worker, args=(r, world, port, n_views, P, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = {r: (a, d, m) for r, a, d, m in (q.get(timeout=120) for _ in range(world))}
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = SideStats(P, device="cpu")
    for step in range(2):
        ref.begin_step()
        for v in range(n_views):
            _fold(ref, v + 10 * step, P)
        ref.all_reduce().apply()
    for r in range(world):
        assert torch.allclose(got[r][0], ref.xyz_gradient_accum, atol=1e-5)
        assert torch.equal(got[r][1], ref.denom) and torch.equal(got[r][2], ref.max_radii2D)


def test_adam_groups_tile_the_flat_layout():
    fb = FlatGradBuffer(P=7, F=15, M=4, device="cpu")
    lr = {"xyz": 1e-4, "f_dc": 2.5e-3, "opacity": 0.05, "scaling": 1e-3, "rotation": 1e-3, "f_language": 2.5e-3}
    groups = fb.adam_groups(lr)
    assert sum(g[1] for g in groups) == fb.flat.numel()
    o = 0
    for (name, shape), g in zip(fb.groups, groups):
        assert fb.views[name].data_ptr() == fb.flat.data_ptr() + 4 * o
        o += g[1]
    sh = groups[2]
    assert sh[4:] == (12, 3, 2.5e-3 / 20.0)                   # f_rest = feature_lr / 20 (gaussian_model.py:409-413)
    assert groups[0][0] == "rotation" and groups[0][1] % 4 == 0   # rows of 4 aligned for every P
