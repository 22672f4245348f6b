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
