"""TEST INFRASTRUCTURE -- golden values of the reference's stand-alone profile callables (lineshapes.py:537-550 Poisson,
:723-767 separatrix profiles), evaluated by the unmodified reference classes in the build container.

    python -m oracle.make_profile_golden        ->  tests/golden/profile_callables.npz
"""
import pathlib

import numpy as np

from baysar_b200.atomic_data import SyntheticADAS

from . import ref_shim

GOLDEN = pathlib.Path(__file__).resolve().parent.parent / "tests" / "golden"


def main():
    ls = ref_shim.install(SyntheticADAS())["lineshapes"]
    rng = np.random.default_rng(20261020)
    x = np.linspace(0.0, 30.0, 40)
    chords = np.linspace(-3.0, 12.0, 9)
    out = {"x": x, "chords": chords}
    cases = {
        "poisson": ls.Poisson(x),
        "linear_separatrix": ls.LinearSeparatrix(chords),
        "cauchy_separatrix": ls.CauchySeparatrix(chords),
        "cauchy_separatrix_centred": ls.CauchySeparatrix(chords, centre=2.5),
        "simple_separatrix_ne": ls.SimpleSeparatrix(chords).electron_density,
        "simple_separatrix_te": ls.SimpleSeparatrix(chords).electron_temperature,
    }
    for name, fn in cases.items():
        b = np.array(fn.bounds, dtype=float)
        thetas = b[:, 0] + (b[:, 1] - b[:, 0]) * rng.uniform(size=(6, len(b)))
        out[name + "_bounds"] = b
        out[name + "_theta"] = thetas
        out[name + "_value"] = np.array([np.array(fn(list(t)), dtype=float) for t in thetas])
    np.savez_compressed(GOLDEN / "profile_callables.npz", **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
