"""Ensemble-sampler step (csrc/ensemble.cuh, baysar_b200/ensemble.py).

CPU: the numpy twin of the kernel's Philox4x32-10 generator against the Random123 known-answer vectors.
GPU (-m gpu): the proposal kernel against the numpy twin on the same counters; the accept kernel's bookkeeping;
a short run on a real posterior (walkers stay consistent with their stored log-posteriors, acceptance is sane).
"""
import ctypes

import numpy as np
import pytest

from baysar_b200.ensemble import philox4x32_10, reference_draws, uniform01


def test_philox_known_answers():
    one = lambda v: np.array([v], dtype=np.uint64)  # noqa: E731
    r = philox4x32_10((one(0), one(0), one(0), one(0)), (0, 0))
    assert [int(x[0]) for x in r] == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    f = 0xFFFFFFFF
    r = philox4x32_10((one(f), one(f), one(f), one(f)), (f, f))
    assert [int(x[0]) for x in r] == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    r = philox4x32_10((one(0x243F6A88), one(0x85A308D3), one(0x13198A2E), one(0x03707344)), (0xA4093822, 0x299F31D0))
    assert [int(x[0]) for x in r] == [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_reference_draws_are_well_formed():
    z, pick, lnu = reference_draws(10000, seed=1234567890123, step=7, n_complement=4096, a=2.0)
    assert z.min() >= 0.5 and z.max() <= 2.0  # z in [1/a, a]
    assert pick.min() >= 0 and pick.max() < 4096
    assert (lnu < 0).all()
    # g(z) ~ 1/sqrt(z) on [1/a, a]: E[z] = (a + 1 + 1/a) / 3
    assert abs(z.mean() - (2 + 1 + 0.5) / 3) < 0.02
    u = uniform01(np.array([0, 0xFFFFFFFF], dtype=np.uint64), np.array([0, 0xFFFFFFFF], dtype=np.uint64))
    assert 0.0 < u[0] < 1e-15 and 1.0 - 1e-15 < u[1] < 1.0


@pytest.mark.gpu
def test_propose_and_accept_kernels_match_numpy_twin():
    import torch

    from baysar_b200 import runtime

    lib = runtime.library()
    rng = np.random.default_rng(5)
    S, C, n, seed, step, a = 3000, 1777, 11, 987654321987, 13, 2.0
    walkers, comp = rng.normal(size=(S, n)), rng.normal(size=(C, n))
    w, c = torch.as_tensor(walkers, device="cuda"), torch.as_tensor(comp, device="cuda")
    prop = torch.empty_like(w)
    fac = torch.empty(S, dtype=torch.float64, device="cuda")
    off = (1 << 33) + 12345  # global index of the first walker: the counter's 64-bit walker word
    step = (5 << 32) + 13     # and a step past 2^32: its high word goes into the counter too
    assert lib.baysar_stretch_propose(w.data_ptr(), c.data_ptr(), n, S, C, a, seed, step, off, prop.data_ptr(),
                                      fac.data_ptr(), None) == 0
    z, pick, lnu = reference_draws(S, seed, step, C, a, walker_offset=off)
    expect = comp[pick] + z[:, None] * (walkers - comp[pick])
    np.testing.assert_allclose(prop.cpu().numpy(), expect, rtol=1e-14, atol=1e-14)
    np.testing.assert_allclose(fac.cpu().numpy(), (n - 1) * np.log(z), rtol=1e-13, atol=1e-13)
    # accept: synthetic log-posteriors, some failed evaluations
    logp_old, logp_new = rng.normal(-50, 3, S), rng.normal(-50, 3, S)
    status = (rng.random(S) < 0.1).astype(np.uint32) * 4
    logp_new[::97] = np.nan
    lo = torch.as_tensor(logp_old.copy(), device="cuda")
    ln = torch.as_tensor(logp_new, device="cuda")
    st = torch.as_tensor(status.astype(np.int32), device="cuda")
    count = torch.zeros(1, dtype=torch.int64, device="cuda")
    assert lib.baysar_stretch_accept(w.data_ptr(), lo.data_ptr(), prop.data_ptr(), ln.data_ptr(), fac.data_ptr(),
                                     st.data_ptr(), n, S, seed, step, off, count.data_ptr(), None) == 0
    take = (status == 0) & np.isfinite(logp_new) & (lnu < (n - 1) * np.log(z) + logp_new - logp_old)
    assert int(count.item()) == int(take.sum())
    np.testing.assert_array_equal(w.cpu().numpy(), np.where(take[:, None], prop.cpu().numpy(), walkers))
    np.testing.assert_array_equal(lo.cpu().numpy(), np.where(take, logp_new, logp_old))


@pytest.mark.gpu
def test_ensemble_on_a_real_posterior():
    import torch

    from baysar_b200.ensemble import StretchEnsemble

    from .test_gpu_parity import gpu_posterior

    post, wl = gpu_posterior("multi_species")
    np.random.seed(11)
    centre = wl.theta_from_tags(post.plasma.slices, post.random_start())
    tb = np.asarray(post.plasma.theta_bounds, dtype=float)
    start = centre + 1e-3 * (tb[:, 1] - tb[:, 0]) * np.random.default_rng(3).normal(size=(256, len(centre)))
    ens = StretchEnsemble(post, start, seed=42)
    first = ens.logp.clone()
    ens.run(12)
    # stored log-posteriors are those of the stored walkers (same kernels, same arithmetic: bit-identical)
    again = post.logp_batch(ens.walkers)
    assert torch.equal(again, ens.logp)
    assert 0.05 < ens.acceptance_fraction < 0.95
    assert float(ens.logp.mean()) > float(first.mean()) - 50.0  # the ensemble does not run away from the mode
    # reproducible from the seed alone
    ens2 = StretchEnsemble(post, start, seed=42)
    ens2.run(12)
    assert torch.equal(ens2.walkers, ens.walkers)


def _two_rank_half_step(rank, world, port, out):
    """One red half-step of a 2-rank GLOBAL ensemble: the partner of every walker is drawn from both ranks' blue halves."""
    import os

    import torch
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from baysar_b200 import runtime
    from baysar_b200.ensemble import gather_complement

    lib = runtime.library()
    rng = np.random.default_rng(100)
    half, n = 64, 5
    all_walkers = rng.normal(size=(world, 2 * half, n))  # [rank][red then blue][param]
    mine = torch.as_tensor(all_walkers[rank], device="cuda")
    comp = gather_complement(mine[half:])
    prop = torch.empty((half, n), dtype=torch.float64, device="cuda")
    fac = torch.empty(half, dtype=torch.float64, device="cuda")
    assert lib.baysar_stretch_propose(mine[:half].data_ptr(), comp.data_ptr(), n, half, comp.shape[0], 2.0, 77, 3,
                                      rank * half, prop.data_ptr(), fac.data_ptr(), None) == 0
    torch.cuda.synchronize()
    out[rank] = (comp.cpu().numpy(), prop.cpu().numpy())
    dist.destroy_process_group()


@pytest.mark.gpu
def test_global_complement_half_step_on_two_gpus():
    """BASELINE configs[3] semantics: ONE ensemble sharded over the ranks.  Each rank's proposals use partners from the
    all-gathered complementary half (both ranks' rows, rank order) and the counter-based draws of its global walker
    indices -- compared with the numpy twin."""
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    world, half, n = 2, 64, 5
    out = mp.Manager().dict()
    mp.spawn(_two_rank_half_step, args=(world, 29731, out), nprocs=world, join=True)
    rng = np.random.default_rng(100)
    all_walkers = rng.normal(size=(world, 2 * half, n))
    blue = np.concatenate([all_walkers[r, half:] for r in range(world)])
    for r in range(world):
        comp, prop = out[r]
        np.testing.assert_array_equal(comp, blue)
        z, pick, _ = reference_draws(half, 77, 3, world * half, 2.0, walker_offset=r * half)
        assert pick.max() >= half  # partners do come from the other rank's rows
        np.testing.assert_allclose(prop, blue[pick] + z[:, None] * (all_walkers[r, :half] - blue[pick]), rtol=1e-14, atol=1e-14)
