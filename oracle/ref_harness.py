"""TEST INFRASTRUCTURE — drive the shimmed reference and record golden vectors.

Build-container only (needs `/root/reference`).  See `oracle/ref_shim.py` for the shims.
"""
import numpy as np

from baysar_b200.atomic_data import SyntheticADAS

from . import ref_shim


def build_reference_posterior(workload, provider=None, seed=0, plasma_attrs=None):
    """Construct the reference's ``BaysarPosterior`` for a `baysar_b200.configs.Workload`."""
    provider = provider or SyntheticADAS()
    mods = ref_shim.install(provider)
    kw = dict(workload.input_kwargs)
    kw.setdefault("line_data", ref_shim.pristine_line_data())
    # PlasmaLine.__init__ and init_test draw from the global RNG; the reference's constructor evaluates the posterior at a
    # random start and lets a numerical failure there propagate (posterior.py:197-217), so a family whose random starts
    # can overflow (ADoubleExpDecay) is constructed with the first seed that survives
    for attempt in range(50):
        np.random.seed(seed + attempt)
        try:
            with ref_shim.quiet():
                input_dict = mods["input_functions"].make_input_dict(**dict(kw))
                profile = workload.make_profile(mods["lineshapes"])
                post = mods["posterior"].BaysarPosterior(input_dict=input_dict, profile_function=profile, skip_test=True,
                                                         curvature=getattr(workload, "curvature", None))
            break
        except (ValueError, TypeError):
            if attempt == 49:
                raise
    for k, v in (plasma_attrs or {}).items():
        setattr(post.plasma, k, v)
    # optional priors: the reference appends passed priors after the defaults (posterior.py:163-165); they need the
    # posterior's own plasma, so they are appended here the way a user script does
    if getattr(workload, "extra_priors", None):
        post(list(workload.theta_from_tags(post.plasma.slices, post.random_start())), True)  # fills plasma_state
        post.posterior_components.extend(workload.make_priors(mods["priors"], plasma=post.plasma,
                                                              plasma_state=post.plasma.plasma_state, posterior=post))
    return post


def describe_setup(post):
    """Constant (theta-independent) state of a reference posterior, as plain arrays."""
    pl = post.plasma
    out = {
        "tags": np.array(list(pl.slices.keys())),
        "slice_start": np.array([s.start for s in pl.slices.values()]),
        "slice_stop": np.array([s.stop for s in pl.slices.values()]),
        "theta_bounds": np.array(pl.theta_bounds, dtype=float),
        "n_params": np.array(pl.n_params),
        "los": np.array(pl.los, dtype=float),
        "species": np.array(pl.species),
        "components": np.array([type(p).__name__ for p in post.posterior_components]),
        "n_chords": np.array(pl.num_chords),
    }
    for c in range(pl.num_chords):
        ch = post.posterior_components[c]
        out[f"c{c}_x_data"] = np.array(ch.x_data, dtype=float)
        out[f"c{c}_y_data"] = np.array(ch.y_data, dtype=float)
        out[f"c{c}_x_data_fm"] = np.array(ch.x_data_fm, dtype=float)
        out[f"c{c}_error"] = np.array(ch.error, dtype=float)
        out[f"c{c}_instrument_function_fm"] = np.array(ch.instrument_function_fm, dtype=float)
        out[f"c{c}_instrument_function"] = np.array(ch.instrument_function, dtype=float)
        out[f"c{c}_line_types"] = np.array([type(l).__name__ for l in ch.lines])
        out[f"c{c}_line_names"] = np.array(
            [l.line_tag if type(l).__name__ == "XLine" else l.line for l in ch.lines]).astype(str)
        out[f"c{c}_doppler_taps"] = np.array(
            [len(l.lineshape.wavelengths_doppler) if hasattr(l, "lineshape") else 0 for l in ch.lines])
    return out


_BALMER_FIELDS = ["rec_sum", "exc_sum", "rec_ne", "rec_te", "exc_ne", "exc_te", "f_rec", "ems_ne", "ems_te",
                  "emission_fitted", "exc_ti"]
_ADAS_FIELDS = ["emission_fitted", "ems_ne", "ems_conc", "ems_te", "ti"]


def evaluate(post, thetas, full_fine_grid=0):
    """Run ``post(theta)`` for each theta; record logP, per-component values, spectra, line scalars.

    A raised exception is recorded as (``logp`` = NaN, ``error`` = "<Type>: <msg>"); the per-component
    values already computed before the raise are kept where available.
    """
    pl = post.plasma
    n = len(thetas)
    ncomp = len(post.posterior_components)
    nch = pl.num_chords
    rec = {
        "theta": np.array(thetas, dtype=float),
        "logp": np.full(n, np.nan),
        "components": np.full((n, ncomp), np.nan),
        "error": np.array([""] * n, dtype=object),
        "ne": np.full((n, len(pl.los)), np.nan),
        "te": np.full((n, len(pl.los)), np.nan),
        "main_ion_density": np.full((n, len(pl.los)), np.nan),
    }
    if pl.contains_hydrogen:
        rec["n0"] = np.full((n, len(pl.los)), np.nan)
    for c in range(nch):
        ch = post.posterior_components[c]
        rec[f"c{c}_spectra"] = np.full((n, len(ch.x_data)), np.nan)
        rec[f"c{c}_preconv_integral"] = np.full(n, np.nan)
        rec[f"c{c}_postconv_integral"] = np.full(n, np.nan)
        rec[f"c{c}_gauss_likelihood"] = np.full(n, np.nan)  # SpectrometerChord.likelihood(cauchy=False)
        rec[f"c{c}_line_scalars"] = np.full((n, len(ch.lines), len(_BALMER_FIELDS)), np.nan)
        if full_fine_grid:
            rec[f"c{c}_fine_lines_sum"] = np.full((min(n, full_fine_grid), len(ch.x_data_fm)), np.nan)
            rec[f"c{c}_fine_prewavecal"] = np.full((min(n, full_fine_grid), len(ch.x_data_fm)), np.nan)

    for i, theta in enumerate(thetas):
        try:
            post.check_call_input(list(theta), False)
            pl(list(theta))
            rec["ne"][i] = pl.plasma_state["electron_density"]
            rec["te"][i] = pl.plasma_state["electron_temperature"]
            rec["main_ion_density"][i] = pl.plasma_state["main_ion_density"]
            if pl.contains_hydrogen:
                rec["n0"][i] = pl.plasma_state[pl.hydrogen_species[0] + "_dens"]
            total = 0
            for j, comp in enumerate(post.posterior_components):
                val = comp()
                rec["components"][i, j] = val
                total = total + val
                if j < nch:
                    ch = comp
                    rec[f"c{j}_spectra"][i] = ch.forward_model()
                    rec[f"c{j}_preconv_integral"][i] = ch.preconv_integral
                    rec[f"c{j}_postconv_integral"][i] = ch.postconv_integral
                    rec[f"c{j}_gauss_likelihood"][i] = ch.likelihood(cauchy=False)
                    for k, l in enumerate(ch.lines):
                        tname = type(l).__name__
                        if tname == "BalmerHydrogenLine":
                            vals = [getattr(l, f) for f in _BALMER_FIELDS]
                        elif tname == "ADAS406Lines":
                            vals = [getattr(l, f) for f in _ADAS_FIELDS]
                        else:
                            vals = [l.emission_fitted]
                        rec[f"c{j}_line_scalars"][i, k, : len(vals)] = np.array(vals, dtype=float).ravel()
                    if full_fine_grid and i < full_fine_grid:
                        rec[f"c{j}_fine_lines_sum"][i] = sum(l().flatten() for l in ch.lines)
                        rec[f"c{j}_fine_prewavecal"][i] = ch.prewavecal_spectra
            post.check_call_output(total)
            rec["logp"][i] = total
        except Exception as e:  # the reference signals numerical failure by raising
            rec["error"][i] = f"{type(e).__name__}: {str(e)[:120]}"
    rec["error"] = rec["error"].astype(str)
    return rec


def draw_thetas(post, n, workload, seed):
    """``n`` thetas: the workload's reference theta first, then the reference's own ``random_start`` draws."""
    np.random.seed(seed)
    out = []
    base = post.random_start()
    out.append(workload.theta_from_tags(post.plasma.slices, base))
    jitter = getattr(workload, "theta_jitter", None)
    rng = np.random.default_rng(seed)
    while len(out) < n:
        if jitter:  # families whose uniform draws overflow: stay within `jitter` of the bound range around the reference theta
            tb = np.asarray(post.plasma.theta_bounds, dtype=float)
            out.append(np.clip(out[0] + jitter * (tb[:, 1] - tb[:, 0]) * rng.uniform(-1, 1, len(tb)), tb[:, 0], tb[:, 1]))
        else:
            out.append(post.random_start())
    return np.array(out)
