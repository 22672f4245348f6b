"""World-size-2 CPU (gloo) test of the N>1 host logic: contiguous sharding, padding of uneven shards, all-gather of
per-walker results, and an ensemble step built on it.  The rendezvous uses 127.0.0.1 (the hostname may not resolve)."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from baysar_b200.parallel import ShardedEvaluator, shard_bounds, stretch_move_proposal


def test_shard_bounds_cover_every_row_once():
    for n in (0, 1, 7, 64, 65, 4096, 65537):
        for world in (1, 2, 3, 8):
            rows = [shard_bounds(n, world, r) for r in range(world)]
            assert rows[0][0] == 0 and rows[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(rows, rows[1:]))
            sizes = [hi - lo for lo, hi in rows]
            assert max(sizes) - min(sizes) <= 1


def _fake_logp(theta):
    return -(theta ** 2).sum(dim=1) - 1.0


def _worker(rank, world, port, n, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(1234)  # the same replicated batch on every rank
        theta = torch.randn(n, 5, generator=g, dtype=torch.float64)
        calls = []

        def eval_fn(rows):
            calls.append(rows.shape[0])
            return _fake_logp(rows)

        ev = ShardedEvaluator(eval_fn)
        full = ev(theta)
        assert torch.equal(full, _fake_logp(theta))
        lo, hi = shard_bounds(n, world, rank)
        assert calls == [hi - lo]  # each rank evaluated only its own rows
        gathered = ev.evaluate_local(theta[lo:lo + n // world])
        assert gathered.shape[0] == world * (n // world)
        # one red/blue ensemble half-step: proposals are identical on every rank (same generator), so is the result
        g2 = torch.Generator().manual_seed(7)
        prop, factor = stretch_move_proposal(theta[: n // 2], theta[n // 2:], generator=g2)
        new = ev(prop)
        accept = torch.log(torch.rand(n // 2, generator=g2, dtype=torch.float64)) < factor + new - full[: n // 2]
        out[rank] = (float(full.sum()), int(accept.sum()))
    finally:
        dist.destroy_process_group()


def test_sharded_all_gather_two_ranks():
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        out = mgr.dict()
        port = 29500 + os.getpid() % 2000
        procs = [ctx.Process(target=_worker, args=(r, 2, port, 101, out)) for r in range(2)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(120)
            assert p.exitcode == 0
        assert out[0] == out[1]


def test_stretch_move_is_affine_combination():
    g = torch.Generator().manual_seed(0)
    w = torch.randn(64, 4, generator=g, dtype=torch.float64)
    c = torch.randn(32, 4, generator=g, dtype=torch.float64)
    prop, factor = stretch_move_proposal(w, c, generator=g)
    z = torch.exp(factor / 3.0)
    assert ((z >= 0.5 - 1e-12) & (z <= 2.0 + 1e-12)).all()
    # proposal - partner is parallel to walker - partner with ratio z: recover the partner and check membership
    keep = (1 - z).abs() > 0.05  # the back-solve is ill-conditioned for z ~ 1
    partner = (prop[keep] - z[keep, None] * w[keep]) / (1 - z[keep, None])
    d = (partner[:, None, :] - c[None, :, :]).norm(dim=2).min(dim=1).values
    assert keep.sum() > 32 and np.all(d.numpy() < 1e-9)


def _complement_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from baysar_b200.ensemble import gather_complement, reference_draws

        half, n = 6, 3
        blue = torch.arange(half * n, dtype=torch.float64).reshape(half, n) + 1000.0 * rank  # this rank's complementary half
        comp = gather_complement(blue)
        # the global ensemble's half-step on the host: partner rows from every rank, draws counted by global walker index
        z, pick, _ = reference_draws(half, seed=5, step=2, n_complement=world * half, walker_offset=rank * half)
        out[rank] = (comp.numpy().copy(), pick.copy(), z.copy())
    finally:
        dist.destroy_process_group()


def test_global_ensemble_complement_two_ranks():
    """The N > 1 ensemble is ONE ensemble: the complementary half is every rank's rows in rank order, identical on all
    ranks, and two ranks that share the seed still draw different partners (global walker index in the counter)."""
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        out = mgr.dict()
        port = 31500 + os.getpid() % 2000
        procs = [ctx.Process(target=_complement_worker, args=(r, 2, port, out)) for r in range(2)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(120)
            assert p.exitcode == 0
        (c0, p0, z0), (c1, p1, z1) = out[0], out[1]
        np.testing.assert_array_equal(c0, c1)
        assert c0.shape == (12, 3) and c0[6, 0] == 1000.0 and c0[0, 0] == 0.0
        assert not np.array_equal(z0, z1) and p0.max() < 12 and p1.max() < 12
