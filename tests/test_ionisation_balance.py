"""`baysar_b200.ionisation_balance` against the unmodified reference (`gcr/ionisation_balance.py`) run on the same
synthetic ADF11 tables (`oracle/make_gcr_golden.py` -> `tests/golden/gcr_balance.npz`)."""
import numpy as np
import pytest

from baysar_b200 import adas_files, ionisation_balance as ib
from baysar_b200.atomic_data import GridTable

from .conftest import GOLDEN


def _tables(z, prefix):
    n = z[f"{prefix}_scd"].shape[0]
    blocks = 1 + np.arange(n)
    return (GridTable(blocks, z[f"{prefix}_ne"], z[f"{prefix}_te"], z[f"{prefix}_scd"]),
            GridTable(blocks, z[f"{prefix}_ne"], z[f"{prefix}_te"], z[f"{prefix}_acd"]))


@pytest.mark.parametrize("prefix,tag,tau", [("z3", "none", None), ("z3", "tau1e-2", 1e-2), ("z3", "tau1e2", 1e2),
                                            ("h", "none", None)])
def test_rate_matrix_and_balance_match_reference(prefix, tag, tau):
    z = np.load(GOLDEN / "gcr_balance.npz")
    scd, acd = _tables(z, prefix)
    rate = ib.build_rates_matrix(scd, acd, tau)
    np.testing.assert_array_equal(rate, z[f"{prefix}_{tag}_rate"])  # same sums of the same table entries: bit-exact
    f = ib.fractional_abundance(scd, acd, tau)
    ref = z[f"{prefix}_{tag}_f"]
    # the reference's null vector comes from an SVD: absolute accuracy ~1e-16 x condition, so compare absolutely
    np.testing.assert_allclose(f, ref, rtol=1e-9, atol=1e-13)
    np.testing.assert_allclose(f.sum(axis=-1), 1.0, rtol=1e-14)
    # and it IS a null vector of the reference's matrix (residual relative to the largest rate of the row)
    resid = np.einsum("...ij,...j->...i", z[f"{prefix}_{tag}_rate"], f)
    scale = np.abs(z[f"{prefix}_{tag}_rate"]).max(axis=(-1, -2))[..., None]
    assert np.abs(resid / scale).max() < 1e-15


def test_transport_stack_and_mean_charge():
    z = np.load(GOLDEN / "gcr_balance.npz")
    scd, acd = _tables(z, "z3")
    np.testing.assert_array_equal(ib.TRANSPORT_TAUS, z["transport_taus"])
    stack = ib.fractional_abundance_transport(scd, acd)
    assert stack.shape == (11,) + z["z3_none_f"].shape
    np.testing.assert_allclose(stack[list(ib.TRANSPORT_TAUS).index(1e2)], z["z3_tau1e2_f"], rtol=1e-9, atol=1e-13)
    zbar = ib.mean_charge(stack)
    assert np.all(np.diff(zbar, axis=0) <= 1e-12)  # smaller tau (less ionisation) never raises the mean charge
    assert np.all((zbar >= 0.0) & (zbar <= 3.0))


def test_extreme_dynamic_range_stays_finite():
    """Cold plasma: ionisation rates ~1e-200 of recombination — the closed form keeps ratios the SVD would lose."""
    ne, te = np.array([1e13, 1e14]), np.array([0.05, 0.1])
    g_ne, g_te = np.meshgrid(ne, te, indexing="ij")
    scd = GridTable([1, 2], ne, te, np.array([1e-8 * np.exp(-13.6 / g_te), 1e-8 * np.exp(-54.4 / g_te)]))
    acd = GridTable([1, 2], ne, te, np.array([1e-12 + 0 * g_te, 4e-12 + 0 * g_te]))
    f = ib.fractional_abundance(scd, acd)
    assert np.all(np.isfinite(f)) and np.all(f >= 0.0)
    np.testing.assert_allclose(f[..., 0], 1.0, rtol=1e-14)
    np.testing.assert_allclose(f[..., 1] / f[..., 0], scd.data[0] / acd.data[0], rtol=1e-12)


def test_tables_from_adf11_text_round_trip():
    """ADF11 text -> parse_adf11 -> balance: the file reader and the balance compose (block numbering 1..Z)."""
    z = np.load(GOLDEN / "gcr_balance.npz")
    scd, acd = _tables(z, "z3")
    scd_f = adas_files.parse_adf11(adas_files.format_adf11(scd, element="LITHIUM", adf11type="scd"))
    acd_f = adas_files.parse_adf11(adas_files.format_adf11(acd, element="LITHIUM", adf11type="acd"))
    f = ib.fractional_abundance(scd_f, acd_f)
    # five decimals of log10 in the file: 1e-5 * ln(10) relative per rate, three rates deep
    np.testing.assert_allclose(f, z["z3_none_f"], rtol=2e-4, atol=1e-12)


def test_input_validation():
    z = np.load(GOLDEN / "gcr_balance.npz")
    scd, acd = _tables(z, "z3")
    bad = GridTable(acd.block, acd.ne * 2.0, acd.te, acd.data)
    with pytest.raises(ValueError):
        ib.build_rates_matrix(scd, bad)
    neg = GridTable(scd.block, scd.ne, scd.te, -scd.data)
    with pytest.raises(ValueError):
        ib.fractional_abundance(neg, acd)


def test_time_dependent_balance_limits_on_the_reference_balance():
    """`time_dependent_abundance` (the run_adas406 primitive, expm of the reference's rate matrix): its long-time limit is
    the steady state the REFERENCE computes (`gcr_balance.npz`, written by gcr/ionisation_balance.py), t -> 0 returns the
    start vector, and the solution is a probability vector that relaxes monotonically in mean charge."""
    z = np.load(GOLDEN / "gcr_balance.npz")
    scd, acd = _tables(z, "z3")
    meta = np.array([1.0, 0.0, 0.0, 0.0])
    late = ib.time_dependent_abundance(scd, acd, 1e9, meta)
    np.testing.assert_allclose(late, z["z3_none_f"], rtol=1e-8, atol=1e-12)
    bare = np.array([0.0, 0.0, 0.0, 1.0])
    np.testing.assert_allclose(ib.time_dependent_abundance(scd, acd, 1e9, bare), z["z3_none_f"], rtol=1e-8, atol=1e-12)
    early = ib.time_dependent_abundance(scd, acd, 1e-30, meta)
    np.testing.assert_allclose(early, np.broadcast_to(meta, early.shape), atol=1e-12)
    zbar = [ib.mean_charge(ib.time_dependent_abundance(scd, acd, t, meta)) for t in np.logspace(-9, 3, 13)]
    assert np.all(np.diff(np.array(zbar), axis=0) >= -1e-12)
    for t in (1e-7, 1e-4, 1e-1):
        f = ib.time_dependent_abundance(scd, acd, t, meta)
        assert np.all(f >= 0.0)
        np.testing.assert_allclose(f.sum(axis=-1), 1.0, rtol=1e-13)
        # it solves the rate equation: (f(t + dt) - f(t - dt)) / (2 dt) = ne R f(t)
        dt = 1e-4 * t
        lhs = (ib.time_dependent_abundance(scd, acd, t + dt, meta) - ib.time_dependent_abundance(scd, acd, t - dt, meta)) / (2 * dt)
        rhs = np.einsum("...ij,...j->...i", z["z3_none_rate"], f) * scd.ne[:, None, None]
        # measured against the gross rates (at a point already in steady state the net rate is pure cancellation)
        gross = np.einsum("...ij,...j->...i", np.abs(z["z3_none_rate"]), f) * scd.ne[:, None, None]
        assert np.abs((lhs - rhs) / (gross.max(axis=-1, keepdims=True) + 1e-300)).max() < 1e-5


def test_file_adas_needs_no_other_provider():
    """FileADAS over the synthetic ADF11 / ADF15 files: the three impurity primitives come from the files alone; values at
    file grid points are the file's, and a posterior's tables build (the GPU parity of that posterior against the
    reference run on the same files is `test_golden_fixtures[file_adas_multi]`)."""
    from baysar_b200 import configs
    from baysar_b200.input_functions import make_input_dict
    from baysar_b200.plasmas import PlasmaLine

    adf = GOLDEN / "adf"
    prov = configs.file_adas_provider(adf)
    t = adas_files.parse_adf11(adf / "plt96_n_synthetic.dat")
    got = prov.adf11_rates("N", "plt", 3, t.te[[2, 7]], t.ne[[1, 5]])
    np.testing.assert_allclose(got, t.select(3)[np.ix_([1, 5], [2, 7])].T, rtol=1e-12)
    p = adas_files.parse_adf15(adf / "c_iii_vsu.pass.synthetic")
    got = prov.adf15_pecs("/home/dgahle/adas/idl_spectra/use_adas208/c_iii_vsu.pass", 7, p.te[[0, 3]], p.ne[[2]])
    np.testing.assert_allclose(got, p.select(7)[np.ix_([2], [0, 3])].T, rtol=1e-12)
    f = prov.ionisation_balance("C", np.array([1.0, 10.0]), np.array([1e13, 1e14]), 1e-3, np.eye(7)[0])
    assert f.shape == (2, 2, 7) and np.all(f >= 0)
    np.testing.assert_allclose(f.sum(-1), 1.0, rtol=1e-13)
    with pytest.raises(NotImplementedError):
        prov.ionisation_balance("O", np.array([1.0]), np.array([1e13]), 1e-3, np.eye(9)[0])
    wl = configs.multi_species_file_adas(adf)
    plasma = PlasmaLine(make_input_dict(**wl.input_kwargs), atomic_data=wl.make_provider())
    assert set(plasma.impurity_ion_bal) == {"N", "C"} and len(plasma.impurity_tecs) > 0
