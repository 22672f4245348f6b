"""Steady-state ionisation balance of an element from its ADF11 ionisation (SCD) and recombination (ACD) tables.

The step before the posterior path: fractional abundances of the charge states on the
``(ne, Te)`` grid of the ADF11 files, optionally with the reference's transport proxy ``tau`` that scales every
ionisation rate.  Mirrors `/root/reference/gcr/ionisation_balance.py`:

* :func:`build_rates_matrix` — `gcr/ionisation_balance.py:16-104`: the ``(Z+1) x (Z+1)`` rate matrix ``R`` of
  ``d f / dt = R f`` at every grid point, ``[ne, Te, row, col]``; block ``k`` of SCD is ``k-1 -> k``, block ``k`` of ACD
  is ``k -> k-1``;
* :func:`fractional_abundance` — `:107-211` (``ionisation_balance``): the normalised null vector of ``R``.  The
  reference finds it with one SVD per grid point (`scipy.linalg.null_space`, `:122-127`); ``R`` is the generator of a
  birth-death chain, so the null vector has the closed form ``f[k+1] / f[k] = SCD[k -> k+1] / ACD[k+1 -> k]`` (row ``k``
  of ``R f = 0`` given rows ``0..k-1``), evaluated here for the whole grid at once, in log space so that abundances
  far below 1e-300 of the dominant stage neither overflow nor lose their ratio;
* :func:`fractional_abundance_transport` — `:214-232`: the stack over ``tau = logspace(4, -6, 11)``;
* :func:`time_dependent_abundance` — what the reference obtains from the proprietary ``run_adas406``
  (`plasmas.py:983-996`, ADAS406: the time-dependent ionisation balance): the solution of ``d f / dt = ne R f`` after
  ``tint`` seconds from the start vector ``meta``, ``f = expm(ne R tint) meta``, with the same rate matrix.  Its long-time
  limit is :func:`fractional_abundance`, which pins it on the reference's own balance.

Tables come from :func:`baysar_b200.adas_files.parse_adf11` (``GridTable(block, ne, Te)``).  Host code (numpy): it runs
once per element when tables are built, not per posterior evaluation.  Parity: `tests/test_ionisation_balance.py`
against arrays produced by the unmodified reference functions (`oracle/make_gcr_golden.py`).
"""
import numpy as np

TRANSPORT_TAUS = np.logspace(4, -6, 11)  # `gcr/ionisation_balance.py:226`


def _check(scd, acd):
    if scd.data.shape != acd.data.shape:
        raise ValueError(f"SCD and ACD tables differ in shape: {scd.data.shape} vs {acd.data.shape}")
    if not (np.array_equal(scd.ne, acd.ne) and np.array_equal(scd.te, acd.te)):
        raise ValueError("SCD and ACD tables are on different (ne, Te) grids")
    if list(scd.block) != list(range(1, len(scd.block) + 1)) or list(acd.block) != list(scd.block):
        raise ValueError("ADF11 blocks must be numbered 1..Z")


def build_rates_matrix(scd, acd, tau=None):
    """``[n_ne, n_Te, Z+1, Z+1]`` rate matrix; ``tau`` multiplies every ionisation rate (`:37`)."""
    _check(scd, acd)
    s = scd.data if tau is None else tau * scd.data  # [Z, ne, Te]: s[k] ionises charge k -> k + 1
    a = acd.data                                      # a[k] recombines charge k + 1 -> k
    z = s.shape[0]
    rate = np.zeros(s.shape[1:] + (z + 1, z + 1))
    for charge in range(z + 1):
        if charge > 0:
            rate[..., charge, charge - 1] = s[charge - 1]      # source from below
            rate[..., charge, charge] -= a[charge - 1]         # loss by recombination
        if charge < z:
            rate[..., charge, charge + 1] = a[charge]          # source from above
            rate[..., charge, charge] -= s[charge]             # loss by ionisation
    return rate


def fractional_abundance(scd, acd, tau=None):
    """``[n_ne, n_Te, Z+1]`` steady-state fractional abundances (each grid point sums to 1)."""
    _check(scd, acd)
    s = scd.data if tau is None else tau * scd.data
    if np.any(s < 0.0) or np.any(acd.data <= 0.0) or not (np.all(np.isfinite(s)) and np.all(np.isfinite(acd.data))):
        raise ValueError("ionisation rates must be >= 0 and recombination rates > 0 (and finite)")
    with np.errstate(divide="ignore"):  # a zero ionisation rate empties every stage above it: log -> -inf -> 0
        log_ratio = np.log(s) - np.log(acd.data)                               # log(f[k+1] / f[k])
    log_f = np.concatenate([np.zeros((1,) + s.shape[1:]), np.cumsum(log_ratio, axis=0)], axis=0)
    log_f -= log_f.max(axis=0, keepdims=True)
    f = np.exp(log_f)
    f /= f.sum(axis=0, keepdims=True)
    return np.moveaxis(f, 0, -1)


def fractional_abundance_transport(scd, acd, taus=TRANSPORT_TAUS):
    """``[n_tau, n_ne, n_Te, Z+1]``: the balance for every transport proxy value (`:214-232`)."""
    return np.stack([fractional_abundance(scd, acd, tau) for tau in np.asarray(taus, dtype=float)], axis=0)


def mean_charge(f_ion):
    """Mean charge ``sum_k k f[k]`` over the last axis (what `plasmas.py:891-905` interpolates per element)."""
    return (f_ion * np.arange(f_ion.shape[-1])).sum(axis=-1)


def time_dependent_abundance(scd, acd, tint, meta):
    """``[n_ne, n_Te, Z+1]`` fractional abundances ``tint`` seconds after the state ``meta`` (length Z+1, normalised
    here): ``expm(ne R tint) meta`` at every grid point, ``R`` from :func:`build_rates_matrix` (cm3/s) and ``ne`` the
    tables' density axis (cm-3).  Tiny negative values of the matrix exponential are clipped and the result is
    renormalised (the exact solution is a probability vector)."""
    from scipy.linalg import expm

    rate = build_rates_matrix(scd, acd)                       # [ne, Te, Z+1, Z+1]
    meta = np.asarray(meta, dtype=float)
    if meta.shape != (rate.shape[-1],) or not meta.sum() > 0:
        raise ValueError("meta must be a non-zero vector with one entry per charge state")
    gen = rate * (np.asarray(scd.ne, dtype=float)[:, None, None, None] * float(tint))
    f = np.clip(expm(gen) @ (meta / meta.sum()), 0.0, None)   # expm works on the stack of matrices
    return f / f.sum(axis=-1, keepdims=True)
