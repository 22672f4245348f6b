"""CPU test of the host-side plan of the folded pixel contraction (csrc/pix_contract.cuh, baysar_abi.cu): every column
evaluated through the plan's tables with the kernel's index arithmetic equals the definition
(1 - a_p) conv[i_p] + a_p conv[i_p + 1] (spectrometers.py:269-310), the explicit edge columns equal the convolution,
and the area identity of pixel_fold_kernel equals the direct trapz of the clipped convolution.  No GPU needed."""
import pytest

from baysar_b200 import runtime
from tests.pxc_cases import CASES, make_case


@pytest.mark.parametrize("F,K,sigma,P", CASES)
@pytest.mark.parametrize("tile_n", [64, 128])
def test_plan_reproduces_definition(F, K, sigma, P, tile_n):
    runtime.build_library()
    s, taps, idx, frac = make_case(F, K, sigma, P, n=1)
    r = runtime.debug_pix_plan_check(s[0], taps, idx, frac, tile_n)
    assert r["err_pixels"] < 1e-13 and r["err_edges"] < 1e-13
    assert r["err_area"] < 1e-8          # float32 taps in the contraction vs float64 tap sum
    assert r["clip_violation"] <= 1e-12  # non-negative taps: the clip can only bite in the edge zones
    assert r["n_folded"] > 0.8 * P and r["ring"] >= 2
    assert 0 <= r["err_colsum"] < 1e-7   # bias-correction identity with the float32 bbar table


@pytest.mark.parametrize("tile_n,P", [(128, 8400), (64, 4200)])
def test_plan_refuses_more_folded_tiles_than_the_pixel_stage_holds(tile_n, P):
    """More than 64 folded column tiles (P > 8192 at 128 columns, P > 4096 at 64) do not fit pixel_fold_kernel's
    per-tile shared-memory tables: the planner must decline, so that the chord stays on the every-column contraction
    (round-1 advisor finding: the plan used to stay on and the pixel stage read past its tables)."""
    runtime.build_library()
    s, taps, idx, frac = make_case(3 * P + 200, 61, 4.0, P, n=1)
    with pytest.raises(RuntimeError, match="no folded-contraction plan"):
        runtime.debug_pix_plan_check(s[0], taps, idx, frac, tile_n)
