"""Design check: the chord kernel's float32 algorithm (tests/kernel_model.py) holds the parity gates on CPU.

Spectra within 1e-5 relative per pixel and the chord log-likelihood within 1e-6 relative of the reference-generated
fixtures (BASELINE.json north_star tolerances)."""
import numpy as np
import pytest

from .conftest import GOLDEN
from .kernel_model import chord_forward_model
from .test_oracle_golden import build
from .workloads import WORKLOADS


@pytest.mark.parametrize("name", sorted(WORKLOADS))
def test_fp32_algorithm_holds_tolerance(name):
    z = np.load(GOLDEN / f"{name}.npz")
    post = build(name)
    worst_s, worst_l = 0.0, 0.0
    for i, theta in enumerate(z["theta"][:8]):
        r = post.evaluate(theta)
        for c, ch in enumerate(post.chords):
            fm, logl = chord_forward_model(post, ch, r["state"], r["line_scalars"][c])
            ref = z[f"c{c}_spectra"][i]
            worst_s = max(worst_s, np.max(np.abs(fm - ref) / np.abs(ref)))
            worst_l = max(worst_l, abs(logl - z["components"][i, c]) / abs(z["components"][i, c]))
    print(f"{name}: spectra rel {worst_s:.2e}  logL rel {worst_l:.2e}")
    assert worst_s < 1e-5
    assert worst_l < 1e-6
