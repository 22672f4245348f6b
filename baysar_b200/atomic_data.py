"""Atomic-data providers for the posterior path.

The reference pulls its atomic data from three places at construction time:

* hydrogen ADF15 photon-emissivity coefficients, ``OpenADAS.load_adf15(get_adf15(...))`` — a
  ``(block, ne, Te)`` table (`/root/reference/baysar/plasmas.py:1277-1316`,
  `/root/reference/OpenADAS/read_adf15.py:68-110`);
* hydrogen ADF11 ``scd/acd/plt/prb`` rate tables, ``OpenADAS.load_adf11(get_adf11(element="h", ...))`` —
  ``(block, ne, Te)`` with ``10**`` applied to the file's log10 grids and rates
  (`plasmas.py:706-721`, `OpenADAS/read_adf11.py:20-69`);
* impurity data through the *proprietary* ADAS python API that the reference still calls at HEAD
  (`run_adas406`, old-signature `read_adf11` / `read_adf15`; `plasmas.py:983-1022,1215-1216`).

There is no network and no ADAS installation in the build or benchmark environment, so the data source is
an explicit *provider object* with exactly those five primitives.  :class:`SyntheticADAS` implements them
with fixed closed-form functions (reproducible without files); the oracle harness patches the reference's
own lookups with the same instance so that both sides see identical tables.  Array layouts follow what the
reference code assumes at the call sites cited on each method.
"""
import numpy as np

_ATOMIC_NUMBER = {"He": 2, "Li": 3, "Be": 4, "B": 5, "C": 6, "N": 7, "O": 8, "F": 9, "Ne": 10}


def number_of_ions(element):
    """Z + 1 charge states (reference `plasmas.py:161-175`)."""
    return _ATOMIC_NUMBER[element] + 1


class GridTable:
    """A ``(block, ne, Te)`` table with its coordinate axes (the reference's ``xarray.DataArray``)."""

    def __init__(self, block, ne, te, data):
        self.block = np.asarray(block, dtype=int)
        self.ne = np.asarray(ne, dtype=float)
        self.te = np.asarray(te, dtype=float)
        self.data = np.asarray(data, dtype=float)
        if self.data.shape != (len(self.block), len(self.ne), len(self.te)):
            raise ValueError("GridTable data must have shape (block, ne, Te)")

    def select(self, block):
        """2-D ``(ne, Te)`` slice for an ADAS block id (``DataArray.sel(block=...)``)."""
        hit = np.where(self.block == block)[0]
        if len(hit) == 0:
            raise KeyError(f"block {block} not in table (have {self.block.min()}..{self.block.max()})")
        return self.data[int(hit[0])]


class AtomicDataProvider:
    """Interface: the five data primitives the posterior's constructor needs."""

    def hydrogen_adf15(self, element, charge):
        """``(block, ne, Te)`` PEC table in cm3/s — `plasmas.py:1290-1291`."""
        raise NotImplementedError

    def hydrogen_adf11(self, adf11type):
        """``(1, ne, Te)`` table of ``scd|acd|plt|prb`` for neutral hydrogen — `plasmas.py:710-721`."""
        raise NotImplementedError

    def ionisation_balance(self, elem, te, dens, tint, meta):
        """Fractional abundances ``[len(dens), len(te), Z+1]`` after ``tint`` seconds, starting from the
        metastable vector ``meta`` (``run_adas406(...)[0]["ion"]``, `plasmas.py:983-996`)."""
        raise NotImplementedError

    def adf11_rates(self, elem, adf11type, is1, te, dens):
        """``[len(te), len(dens)]`` power coefficients (old ``read_adf11(..., all=True)``, `plasmas.py:999-1018`)."""
        raise NotImplementedError

    def adf15_pecs(self, file, block, te, dens):
        """``[len(te), len(dens)]`` PECs (old ``read_adf15(file, block, te, ne, all=True)[0]``, `plasmas.py:1215-1216`)."""
        raise NotImplementedError


class SyntheticADAS(AtomicDataProvider):
    """Closed-form, file-free stand-in tables (smooth, positive, physically shaped).

    These are NOT real atomic data; they exist so that the oracle (the reference run under import shims)
    and the CUDA path can be driven by bit-identical tables with no network and no ADAS licence.
    """

    H_NE = np.logspace(np.log10(5e7), np.log10(2e15), 24)
    H_TE = np.logspace(np.log10(0.2), 4.0, 29)
    H_BLOCKS = 30  # blocks 1..18 excitation-like (n = block + 2 -> 2), 19..30 recombination-like

    def __init__(self, scale=1.0):
        self.scale = float(scale)

    # ---- hydrogen ---------------------------------------------------------------------------------
    def _h_pec(self, block, ne, te):
        if block <= 18:
            n_up = block + 2.0
            de = 13.6 * (0.25 - 1.0 / n_up**2)  # Balmer-like excitation energy from n=2 scale
            return (self.scale * 2.4e-9 / n_up**3 * np.exp(-(10.2 + de) / te)
                    * (1.0 + 0.08 * np.log10(ne / 1e13)) / (1.0 + te / 400.0))
        n_up = block - 16.0
        return self.scale * 6.0e-13 / n_up**1.5 * te**-0.85 * (ne / 1e13) ** 0.12

    def hydrogen_adf15(self, element="h", charge=0):
        g_ne, g_te = np.meshgrid(self.H_NE, self.H_TE, indexing="ij")
        data = np.array([self._h_pec(b + 1, g_ne, g_te) for b in range(self.H_BLOCKS)])
        return GridTable(1 + np.arange(self.H_BLOCKS), self.H_NE, self.H_TE, data)

    def hydrogen_adf11(self, adf11type):
        g_ne, g_te = np.meshgrid(self.H_NE, self.H_TE, indexing="ij")
        if adf11type == "scd":
            data = 2.6e-8 * np.exp(-13.6 / g_te) * np.sqrt(g_te) / (1.0 + g_te / 60.0) * (1 + 0.05 * np.log10(g_ne / 1e13))
        elif adf11type == "acd":
            data = 2.2e-13 * g_te**-0.72 * (1.0 + (g_ne / 1e14) ** 0.28)
        elif adf11type == "plt":
            data = 1.3e-25 * np.exp(-10.2 / g_te) * (1.0 + 0.02 * np.log10(g_ne / 1e13))
        elif adf11type == "prb":
            data = 2.7e-27 * g_te**-0.32 * (1.0 + (g_ne / 1e15) ** 0.2)
        else:
            raise KeyError(adf11type)
        return GridTable([1], self.H_NE, self.H_TE, data[None])

    # ---- impurities -------------------------------------------------------------------------------
    def ionisation_balance(self, elem, te, dens, tint, meta):
        nion = number_of_ions(elem)
        z = np.arange(nion, dtype=float)
        lt = np.log10(np.asarray(te, dtype=float))[None, :, None]
        ln = np.log10(np.asarray(dens, dtype=float))[:, None, None]
        start = float(np.argmax(np.asarray(meta)))  # 0 -> ionising from neutral, Z -> recombining from bare
        relax = 1.0 / (1.0 + (1e-3 / float(tint)) ** 0.6)  # 0 (t -> 0) .. 1 (t -> inf)
        z_eq = np.clip(1.4 + 2.1 * lt + 0.06 * (ln - 14.0), 0.0, nion - 1.0)
        zbar = start + (z_eq - start) * relax
        width = 0.55 + 0.25 * relax
        w = np.exp(-0.5 * ((z[None, None, :] - zbar) / width) ** 2) + 1e-14
        return w / w.sum(-1, keepdims=True)

    def adf11_rates(self, elem, adf11type, is1, te, dens):
        te = np.asarray(te, dtype=float)
        dens = np.asarray(dens, dtype=float)
        if adf11type == "plt":
            return 1.1e-26 * is1 * np.exp(-4.5 * is1 / te)[:, None] * (dens[None, :] / 1e14) ** 0.08
        if adf11type == "prb":
            return 2.0e-27 * is1**2 * (te**-0.4)[:, None] * (dens[None, :] / 1e14) ** 0.05
        raise KeyError(adf11type)

    def adf15_pecs(self, file, block, te, dens):
        te = np.asarray(te, dtype=float)
        dens = np.asarray(dens, dtype=float)
        # a weak, deterministic dependence on the file label keeps different ions' lines distinct
        salt = 1.0 + 0.1 * ((sum(ord(c) for c in str(file)) % 7) - 3) / 3.0
        if block < 30:
            v = salt * 1.2e-9 / (1.0 + block) * np.exp(-(3.2 + 0.11 * block) / te)[:, None] * (dens[None, :] / 1e14) ** -0.1
        else:
            v = salt * 0.9e-12 / (block - 20.0) * (te**-0.9)[:, None] * (dens[None, :] / 1e14) ** 0.2
        return self.scale * v
