"""Host-side mirror of the reference interface: behaviours that need no GPU."""
import numpy as np
import pytest

from baysar_b200 import configs
from baysar_b200.atomic_data import SyntheticADAS
from baysar_b200.input_functions import make_input_dict
from baysar_b200.plasmas import PlasmaLine


@pytest.mark.parametrize("no_sample_neutrals,exc", [(True, KeyError), (False, ValueError)])
def test_second_hydrogen_isotope_raises_like_the_reference(no_sample_neutrals, exc):
    """The reference cannot evaluate two hydrogen isotopes (run under the shims in the build container: KeyError
    'H_0_dens' from total_power with unsampled neutrals, plasmas.py:1184-1185; ValueError 'shapes (100,) and (1,) not
    aligned' from BalmerHydrogenLine with sampled ones, linemodels.py:817), and its constructor evaluates once
    (posterior.py:197-217): the mirror raises the same types at construction."""
    wl = configs.multi_species(1, data="flat")
    kw = dict(wl.input_kwargs)
    kw.update(species=["D", "H"], ions=[["0"], ["0"]], no_sample_neutrals=no_sample_neutrals)
    with pytest.raises(exc):
        PlasmaLine(make_input_dict(**kw), atomic_data=SyntheticADAS())


def test_plasma_switch_assignment_notifies_the_posterior():
    """recombining / reverse_electroneutrality / ne_tau are compiled into the device model: assigning one after
    construction must invalidate it (PlasmaLine.__setattr__ -> BaysarPosterior._models.clear)."""
    wl = configs.multi_species(1, data="flat")
    plasma = PlasmaLine(make_input_dict(**wl.input_kwargs), atomic_data=SyntheticADAS())
    calls = []
    plasma._on_option_change = lambda: calls.append(1)
    plasma.recombining = True
    plasma.some_other_attribute = 3
    plasma.ne_tau = False
    assert len(calls) == 2 and plasma.recombining is True


def test_closed_form_profile_families_host_statements():
    """The host classes of the closed-form families carry the device function name and the reference's bounds."""
    import baysar_b200.lineshapes as ls

    x = np.linspace(0.0, 20.0, 100)
    slab = ls.DoubleSlabPlasma(x=x)
    ne = slab.electron_density([14.0, 13.5, 6.0])
    te = slab.electron_temperature([0.2, 1.0])
    assert slab.electron_temperature.centre == 6.0  # the density hands its centre to the temperature
    assert ne[x < 6.0].min() == 10 ** 14.0 and te[x > 6.0].max() == 10 ** 1.0
    assert ls.SimplePlasma(x=x).electron_density.bounds[0] == [12, 16]
    assert ls.SimpleGaussianPlasma(x=x).electron_temperature.number_of_variables == 2
    assert {c.fn for c in (ls.ExpDecay(x), ls.ADoubleExpDecay(x), ls.FlatProfile(x, [0, 2]))} == \
        {"PF_EXP_DECAY", "PF_DOUBLE_EXP_DECAY", "PF_FLAT"}


def test_stand_alone_profile_callables_match_the_reference():
    """Poisson (lineshapes.py:537-550) and the separatrix profiles (:723-767) against values the unmodified reference
    classes returned (`oracle/make_profile_golden.py`)."""
    from baysar_b200 import lineshapes as ls

    from .conftest import GOLDEN

    z = np.load(GOLDEN / "profile_callables.npz")
    x, chords = z["x"], z["chords"]
    cases = {
        "poisson": ls.Poisson(x),
        "linear_separatrix": ls.LinearSeparatrix(chords),
        "cauchy_separatrix": ls.CauchySeparatrix(chords),
        "cauchy_separatrix_centred": ls.CauchySeparatrix(chords, centre=2.5),
        "simple_separatrix_ne": ls.SimpleSeparatrix(chords).electron_density,
        "simple_separatrix_te": ls.SimpleSeparatrix(chords).electron_temperature,
    }
    for name, fn in cases.items():
        np.testing.assert_array_equal(np.array(fn.bounds, dtype=float), z[name + "_bounds"])
        for theta, want in zip(z[name + "_theta"], z[name + "_value"]):
            np.testing.assert_allclose(fn(list(theta)), want, rtol=1e-14, atol=0, err_msg=name)
