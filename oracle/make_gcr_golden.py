"""TEST INFRASTRUCTURE — what the REFERENCE's ionisation-balance builder makes of synthetic ADF11 tables.

Runs the unmodified `/root/reference/gcr/ionisation_balance.py` (`build_rates_matrix`, `solve_rate_matrix`) under the
xarray stand-in of `oracle/ref_shim.py`, with its `get_adf11` / `load_adf11` bound to a synthetic three-block element
(and to the one-block hydrogen tables of `SyntheticADAS`), and stores inputs and outputs in
`tests/golden/gcr_balance.npz`.  The loop over grid points of the reference's `ionisation_balance` (`:178-199`: one
`null_space` per `(ne, Te)` in `itertools.product` order, reshaped to `[ne, Te, charge]`) is repeated here around the
reference's own `solve_rate_matrix`, because its `DataArray.sel(ne=, Te=)` indexing is outside the stand-in.
`tests/test_ionisation_balance.py` pins `baysar_b200.ionisation_balance` on these arrays.
Run in the build container only: `python -m oracle.make_gcr_golden`.
"""
import importlib
import pathlib
import sys

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from baysar_b200.atomic_data import GridTable, SyntheticADAS  # noqa: E402
from oracle import ref_shim  # noqa: E402


def synthetic_element(z=3, n_ne=5, n_te=9):
    """SCD / ACD tables of a Z = 3 element with a moderate dynamic range (the reference's SVD null space loses
    abundances below ~1e-16 of the dominant stage; the closed form does not, so the comparison stays where both hold)."""
    ne = np.logspace(12.0, 15.0, n_ne)
    te = np.logspace(0.0, 2.0, n_te)
    g_ne, g_te = np.meshgrid(ne, te, indexing="ij")
    scd, acd = [], []
    for k in range(z):
        e_ion = 5.0 * (k + 1) ** 1.5
        scd.append(3e-8 / (k + 1) * np.exp(-e_ion / g_te) * np.sqrt(g_te) * (1 + 0.04 * np.log10(g_ne / 1e13)))
        acd.append(4e-12 * (k + 1) ** 2 * g_te**-0.6 * (1 + (g_ne / 1e14) ** 0.25))
    blocks = 1 + np.arange(z)
    return GridTable(blocks, ne, te, np.array(scd)), GridTable(blocks, ne, te, np.array(acd))


def reference_balance(gcr, tables, tau):
    """`ionisation_balance` of the reference (`:178-211`) on bound tables: rate matrix and fractional abundance."""
    def _as_dataarray(tab):
        return ref_shim.DataArray(tab.data, dims=("block", "ne", "Te"), coords=dict(block=tab.block, ne=tab.ne, Te=tab.te))

    gcr.get_adf11 = lambda element, adf11type, **k: adf11type
    gcr.load_adf11 = lambda adf11, passed=False: _as_dataarray(tables[adf11])
    rate = gcr.build_rates_matrix("synthetic", tau=tau)
    r = np.asarray(rate.data)
    n_ne, n_te, z1 = r.shape[0], r.shape[1], r.shape[2]
    matrices = [r[i, j] for i in range(n_ne) for j in range(n_te)]
    f = gcr.solve_rate_matrix(rate_matrix=matrices, fractional_abundance=np.zeros((len(matrices), z1)))
    return r, f.reshape(n_ne, n_te, z1)


def main():
    ref_shim.install(SyntheticADAS())
    gcr = importlib.import_module("gcr.ionisation_balance")
    arrays = {}
    scd, acd = synthetic_element()
    arrays.update(z3_ne=scd.ne, z3_te=scd.te, z3_scd=scd.data, z3_acd=acd.data)
    for tag, tau in (("none", None), ("tau1e-2", 1e-2), ("tau1e2", 1e2)):
        r, f = reference_balance(gcr, {"scd": scd, "acd": acd}, tau)
        arrays[f"z3_{tag}_rate"], arrays[f"z3_{tag}_f"] = r, f
    prov = SyntheticADAS()
    h_scd, h_acd = prov.hydrogen_adf11("scd"), prov.hydrogen_adf11("acd")
    arrays.update(h_ne=h_scd.ne, h_te=h_scd.te, h_scd=h_scd.data, h_acd=h_acd.data)
    r, f = reference_balance(gcr, {"scd": h_scd, "acd": h_acd}, None)
    arrays["h_none_rate"], arrays["h_none_f"] = r, f
    arrays["transport_taus"] = np.logspace(4, -6, 11)  # `gcr/ionisation_balance.py:226`
    np.savez_compressed(ROOT / "tests" / "golden" / "gcr_balance.npz", **arrays)
    print("wrote gcr_balance.npz", {k: v.shape for k, v in arrays.items()})


if __name__ == "__main__":
    main()
