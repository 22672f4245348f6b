"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path, called through the C ABI, against
(1) the reference-generated golden fixtures, (2) the oracle port on fresh seeded thetas, (3) size-independent
properties at BASELINE batch sizes.  Tolerances are the north_star's: logP 1e-6 relative, spectra 1e-5 per pixel.
"""
import numpy as np
import pytest

from .conftest import GOLDEN
from .workloads import WORKLOADS

pytestmark = pytest.mark.gpu

LOGP_RTOL = 1e-6
SPECTRA_RTOL = 1e-5


def gpu_posterior(name):
    from baysar_b200.atomic_data import SyntheticADAS
    from baysar_b200.input_functions import make_input_dict
    import baysar_b200.lineshapes as host_ns
    from baysar_b200.posterior import BaysarPosterior

    wl = WORKLOADS[name]()
    np.random.seed(3)
    import baysar_b200.priors as host_priors

    post = BaysarPosterior(make_input_dict(**wl.input_kwargs), profile_function=wl.make_profile(host_ns),
                           atomic_data=wl.make_provider(), skip_test=True, curvature=wl.curvature,
                           priors=wl.make_priors(host_priors, late=False))
    # host-evaluated priors need the live plasma / posterior: appended afterwards, as scripts for the reference do
    post.posterior_components.extend(wl.make_priors(host_priors, plasma=post.plasma, plasma_state=post.plasma.plasma_state,
                                                    posterior=post, late=True))
    for k, v in wl.plasma_attrs.items():  # the reference's post-construction switches: assigned the way a user does
        setattr(post.plasma, k, v)
    return post, wl


def oracle_posterior(name):
    from .test_oracle_golden import build

    return build(name)


@pytest.fixture(scope="module")
def posteriors():
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = gpu_posterior(name)
        return cache[name]

    return get


@pytest.mark.parametrize("name", sorted(WORKLOADS))
def test_golden_fixtures(name, posteriors):
    post, _ = posteriors(name)
    z = np.load(GOLDEN / f"{name}.npz")
    theta = z["theta"]
    logp = post(theta)
    assert (post.last_status == 0).all(), post.last_status
    np.testing.assert_allclose(logp, z["logp"], rtol=LOGP_RTOL)
    comps, status = post.components_batch(theta)
    ncol = comps.shape[1]  # chords + device priors; host-evaluated priors follow in the reference's component list
    for i in range(len(theta)):  # 1e-6 of the posterior the components add up to (see tests/test_emulated_kernels.py)
        np.testing.assert_allclose(comps[i], z["components"][i, :ncol], rtol=LOGP_RTOL, atol=LOGP_RTOL * abs(z["logp"][i]))
    if z["components"].shape[1] > ncol:  # host priors, through the scalar protocol: values from the published state
        assert len(post.host_priors) == z["components"].shape[1] - ncol
        for i in range(min(4, len(theta))):
            np.testing.assert_allclose(post(theta[i]), z["logp"][i], rtol=LOGP_RTOL)
            np.testing.assert_allclose([p() for p in post.host_priors], z["components"][i, ncol:], rtol=1e-8, atol=1e-12)
    lines, _, prof = post.lines_batch(theta, want_profiles=True)
    # CUDA's pow(10, x) and numpy's differ by an ulp now and then; the mesh family interpolates between knot values,
    # whose difference amplifies that ulp (2e-12 seen) — every other family holds 1e-12
    prof_rtol = 2e-11 if "mesh" in name else 1e-12
    np.testing.assert_allclose(prof[:, 0], z["ne"], rtol=prof_rtol)
    np.testing.assert_allclose(prof[:, 1], z["te"], rtol=prof_rtol)
    np.testing.assert_allclose(prof[:, 2], z["main_ion_density"], rtol=10 * prof_rtol)
    if "n0" in z.files:
        np.testing.assert_allclose(prof[:, 3], z["n0"], rtol=1e-10)
    for c in range(len(post.chords)):
        spectra, st = post.spectra_batch(theta, chord=c)
        np.testing.assert_allclose(spectra, z[f"c{c}_spectra"], rtol=SPECTRA_RTOL)
        ref = z[f"c{c}_line_scalars"]
        for k, u in enumerate(post.model_info["chord_line_uline"][c]):
            kind = str(z[f"setup_c{c}_line_types"][k])
            rec = lines[:, u]
            if kind == "BalmerHydrogenLine":
                np.testing.assert_allclose(rec[:, [0, 1, 2, 3, 4, 5, 7, 8, 9, 10, 6]], ref[:, k, :11], rtol=1e-9)
            elif kind == "ADAS406Lines":
                np.testing.assert_allclose(rec[:, :5], ref[:, k, :5], rtol=1e-9)
            else:
                np.testing.assert_allclose(rec[:, :1], ref[:, k, :1], rtol=1e-12)


@pytest.mark.parametrize("name", ["balmer_series", "multi_species", "multi_hot", "multi_calibrated", "aug_demo"])
def test_against_oracle_on_fresh_thetas(name, posteriors):
    post, wl = posteriors(name)
    oracle = oracle_posterior(name)
    rng = np.random.default_rng(99)
    lo, hi = post.plasma.theta_bounds[:, 0], post.plasma.theta_bounds[:, 1]
    theta = lo + rng.random((48, post.plasma.n_params)) * (hi - lo)
    logp = post(theta)
    status = post.last_status.copy()
    spectra, _ = post.spectra_batch(theta, chord=0)
    from oracle.numpy_port import OracleError

    checked = 0
    for i, th in enumerate(theta):
        try:
            ref = oracle.evaluate(th)
        except OracleError:
            assert status[i] != 0, f"oracle raised but status is clean for theta {i}"
            continue
        assert status[i] == 0, (i, status[i])
        np.testing.assert_allclose(logp[i], ref["logp"], rtol=LOGP_RTOL)
        np.testing.assert_allclose(spectra[i], ref["spectra"][0], rtol=SPECTRA_RTOL)
        checked += 1
    assert checked >= 40


def test_scalar_protocol_and_exceptions(posteriors):
    post, wl = posteriors("multi_species")
    z = np.load(GOLDEN / "multi_species.npz")
    th = z["theta"][0]
    val = post(th)
    assert isinstance(val, float)
    np.testing.assert_allclose(val, z["logp"][0], rtol=LOGP_RTOL)
    np.testing.assert_allclose(post.cost(th), -z["logp"][0], rtol=LOGP_RTOL)
    # reference-style introspection after a scalar call
    np.testing.assert_allclose(post.posterior_components[0].forward_model(), z["c0_spectra"][0], rtol=SPECTRA_RTOL)
    np.testing.assert_allclose(post.posterior_components[0](), z["components"][0, 0], rtol=LOGP_RTOL)
    np.testing.assert_allclose(post.plasma.plasma_state["electron_density"], z["ne"][0], rtol=1e-12)
    line = post.posterior_components[0].lines[0]
    np.testing.assert_allclose([line.rec_sum, line.exc_ne, line.f_rec], z["c0_line_scalars"][0, 0, [0, 4, 6]], rtol=1e-9)
    bad = th.copy()
    bad[2] = np.nan
    with pytest.raises(TypeError):
        post(bad)
    bad = th.copy()
    bad[post.plasma.slices["N_1_tau"].start] = -7.5  # below the ADAS tau grid: the reference raises ValueError
    with pytest.raises(ValueError):
        post(bad)
    val = post(bad, True)  # skip_error=True never raises; the failure is in the status word instead
    assert isinstance(val, float) and post.last_status[0] & (1 << 2)


def test_batch_properties_at_baseline_size(posteriors):
    """configs[1] at its full batch (4096 thetas): determinism, batch-splitting and permutation invariance,
    host-buffer path == device path."""
    import torch

    post, wl = posteriors("balmer_series")
    rng = np.random.default_rng(5)
    lo, hi = post.plasma.theta_bounds[:, 0], post.plasma.theta_bounds[:, 1]
    theta = lo + rng.random((4096, post.plasma.n_params)) * (hi - lo)
    dev = torch.as_tensor(theta, device="cuda")
    a = post.logp_batch(dev).cpu().numpy()
    b = post.logp_batch(dev).cpu().numpy()
    np.testing.assert_array_equal(a, b)
    parts = np.concatenate([post.logp_batch(dev[:1000]).cpu().numpy(), post.logp_batch(dev[1000:]).cpu().numpy()])
    np.testing.assert_array_equal(a, parts)
    perm = rng.permutation(4096)
    np.testing.assert_array_equal(a[perm], post.logp_batch(dev[torch.as_tensor(perm, device="cuda")]).cpu().numpy())
    host = post(theta)
    np.testing.assert_array_equal(a, host)
    ok = post.last_status == 0
    assert ok.mean() > 0.9
    assert np.all(np.isfinite(a[ok])) and np.all(a[ok] < 0)
    # a sample of the big batch against the oracle
    oracle = oracle_posterior("balmer_series")
    for i in rng.choice(np.where(ok)[0], 16, replace=False):
        np.testing.assert_allclose(a[i], oracle(theta[i]), rtol=LOGP_RTOL)


def test_background_linearity_property(posteriors):
    """forward_model = clip(lines + background): raising the background by d raises every pixel that is not clipped
    by exactly d (spectrometers.py:299-310) — a size-independent property of the fused epilogue."""
    post, wl = posteriors("multi_species")
    z = np.load(GOLDEN / "multi_species.npz")
    th = np.repeat(z["theta"][:1], 2, axis=0)
    b = post.plasma.slices["background_0"].start
    th[1, b] = np.log10(10 ** th[0, b] * 1.5)
    spectra, _ = post.spectra_batch(th, chord=0)
    d = 10 ** th[1, b] - 10 ** th[0, b]
    np.testing.assert_allclose(spectra[1] - spectra[0], d, rtol=1e-6)


def _fresh_posterior(workload, options):
    """A posterior created with explicit kernel-selection options (baysar_options_t)."""
    from baysar_b200.atomic_data import SyntheticADAS
    from baysar_b200.input_functions import make_input_dict
    from baysar_b200.posterior import BaysarPosterior

    np.random.seed(3)
    return BaysarPosterior(make_input_dict(**workload.input_kwargs), atomic_data=SyntheticADAS(), skip_test=True,
                           options=options)


def test_tensor_core_path_agrees_with_fused_path():
    """The tcgen05 contraction (spectral -> contraction -> pixel stages) and the fused FP32 kernel are two
    implementations of the same chord evaluation: compare them on 640 thetas, and both against the oracle."""
    from baysar_b200 import configs

    wl = configs.balmer_series()
    tensor = _fresh_posterior(wl, {})
    fused = _fresh_posterior(wl, {"tensor_cores": -1})
    assert tensor.model.uses_tensor_cores and not fused.model.uses_tensor_cores
    rng = np.random.default_rng(11)
    lo, hi = tensor.plasma.theta_bounds[:, 0], tensor.plasma.theta_bounds[:, 1]
    theta = lo + rng.random((640, tensor.plasma.n_params)) * (hi - lo)
    a, b = tensor(theta), fused(theta)
    ok = (tensor.last_status == 0) & (fused.last_status == 0)
    assert ok.mean() > 0.9
    np.testing.assert_allclose(a[ok], b[ok], rtol=LOGP_RTOL)
    sa, _ = tensor.spectra_batch(theta[:64], 0)
    sb, _ = fused.spectra_batch(theta[:64], 0)
    np.testing.assert_allclose(sa, sb, rtol=SPECTRA_RTOL)
    oracle = oracle_posterior("balmer_series")
    for i in np.where(ok)[0][:24]:
        ref = oracle.evaluate(theta[i])
        np.testing.assert_allclose(a[i], ref["logp"], rtol=LOGP_RTOL)
        if i < 64:
            np.testing.assert_allclose(sa[i], ref["spectra"][0], rtol=SPECTRA_RTOL)


def test_contraction_variants_agree():
    """The folded pixel contraction (default, 128-column tiles), its 64-column variant and the every-column Toeplitz
    contraction (the path a sampled wavelength map takes) evaluate the same chord: logP within the north-star
    tolerance of each other and of the oracle, spectra within 1e-5 per pixel."""
    from baysar_b200 import configs

    wl = configs.balmer_series()
    folded = _fresh_posterior(wl, {})
    folded64 = _fresh_posterior(wl, {"folded_tile": 64})
    toeplitz = _fresh_posterior(wl, {"folded_tile": -1})
    assert folded.model.folded_contraction[0][1] == 128 and folded64.model.folded_contraction[0][1] == 64
    assert toeplitz.model.folded_contraction[0][0] == 0 and toeplitz.model.uses_tensor_cores
    rng = np.random.default_rng(21)
    lo, hi = folded.plasma.theta_bounds[:, 0], folded.plasma.theta_bounds[:, 1]
    theta = lo + rng.random((512, folded.plasma.n_params)) * (hi - lo)
    a, b, c = folded(theta), folded64(theta), toeplitz(theta)
    # the B operand from the precomputed image (default) and from the builder warps: the same bytes reach the tensor core
    for opts, want in (({"contraction_b_image": -1}, a), ({"contraction_b_image": -1, "folded_tile": 64}, b)):
        built = _fresh_posterior(wl, opts)
        np.testing.assert_array_equal(built(theta), want)
    ok = (folded.last_status == 0) & (folded64.last_status == 0) & (toeplitz.last_status == 0)
    assert ok.mean() > 0.9
    np.testing.assert_allclose(a[ok], c[ok], rtol=LOGP_RTOL)
    np.testing.assert_allclose(b[ok], c[ok], rtol=LOGP_RTOL)
    sa, _ = folded.spectra_batch(theta[:32], 0)
    sc, _ = toeplitz.spectra_batch(theta[:32], 0)
    np.testing.assert_allclose(sa, sc, rtol=SPECTRA_RTOL)
    oracle = oracle_posterior("balmer_series")
    for i in np.where(ok)[0][:8]:
        np.testing.assert_allclose(b[i], oracle.evaluate(theta[i])["logp"], rtol=LOGP_RTOL)


@pytest.mark.parametrize("n_los", [500, 1000])
def test_long_line_of_sight_team_kernels(n_los):
    """Long lines of sight run K1 with one CTA per theta (four warps at L = 500, eight at L = 1000): log-posterior and
    spectra against the oracle on a shape-sweep chord (BASELINE configs[4])."""
    import baysar_b200.lineshapes as host_ns
    import oracle.numpy_port as oracle_ns
    from baysar_b200 import configs
    from baysar_b200.atomic_data import SyntheticADAS
    from baysar_b200.input_functions import make_input_dict
    from baysar_b200.posterior import BaysarPosterior
    from oracle.numpy_port import OracleError, OraclePosterior

    wl = configs.synthetic_sweep(n_los, 10, 512, balmer=True)
    np.random.seed(0)
    post = BaysarPosterior(make_input_dict(**wl.input_kwargs), profile_function=wl.make_profile(host_ns),
                           atomic_data=SyntheticADAS(), skip_test=True)
    oracle = OraclePosterior(make_input_dict(**wl.input_kwargs), SyntheticADAS(), wl.make_profile(oracle_ns))
    rng = np.random.default_rng(31)
    lo, hi = post.plasma.theta_bounds[:, 0], post.plasma.theta_bounds[:, 1]
    theta = lo + rng.random((6, post.plasma.n_params)) * (hi - lo)
    logp = post(theta)
    spectra, _ = post.spectra_batch(theta, chord=0)
    checked = 0
    for i, th in enumerate(theta):
        try:
            ref = oracle.evaluate(th)
        except OracleError:
            assert post.last_status[i] != 0
            continue
        assert post.last_status[i] == 0
        np.testing.assert_allclose(logp[i], ref["logp"], rtol=LOGP_RTOL)
        np.testing.assert_allclose(spectra[i], ref["spectra"][0], rtol=SPECTRA_RTOL)
        checked += 1
    assert checked >= 3


def test_tensor_core_path_multi_chord_and_row_chunks():
    """Two chords through the tensor path with a workspace budget that forces several row chunks per chord."""
    from baysar_b200 import configs
    from baysar_b200.atomic_data import SyntheticADAS
    from baysar_b200.input_functions import make_input_dict
    from oracle.numpy_port import OraclePosterior

    wl = configs.balmer_two_chords()
    small = _fresh_posterior(wl, {"ifc_budget_mb": 40})   # 384 rows per pass
    fused = _fresh_posterior(wl, {"tensor_cores": -1})
    assert small.model.uses_tensor_cores and len(small.chords) == 2
    rng = np.random.default_rng(12)
    lo, hi = small.plasma.theta_bounds[:, 0], small.plasma.theta_bounds[:, 1]
    theta = lo + rng.random((1000, small.plasma.n_params)) * (hi - lo)
    a, b = small(theta), fused(theta)
    ok = (small.last_status == 0) & (fused.last_status == 0)
    assert ok.mean() > 0.9
    np.testing.assert_allclose(a[ok], b[ok], rtol=LOGP_RTOL)
    assert small.model.launches_last_eval > 3 + 3 * 2  # more than one chunk per chord
    oracle = OraclePosterior(make_input_dict(**wl.input_kwargs), SyntheticADAS())
    for i in np.where(ok)[0][:12]:
        ref = oracle.evaluate(theta[i])
        np.testing.assert_allclose(a[i], ref["logp"], rtol=LOGP_RTOL)
        for c in range(2):
            spec, _ = small.spectra_batch(theta[i:i + 1], c)
            np.testing.assert_allclose(spec[0], ref["spectra"][c], rtol=SPECTRA_RTOL)


def test_batched_finite_difference_gradient(posteriors):
    """posterior.gradient: one batched launch of 2 n evaluations reproduces the oracle's central differences."""
    post, wl = posteriors("multi_species")
    oracle = oracle_posterior("multi_species")
    z = np.load(GOLDEN / "multi_species.npz")
    theta = z["theta"][0]
    grad = post.gradient(theta)
    tb = np.asarray(oracle.theta_bounds, dtype=float)
    h = 1e-3 * (tb[:, 1] - tb[:, 0])
    ref = np.zeros(len(theta))
    for k in range(len(theta)):
        up, dn = theta.copy(), theta.copy()
        up[k] += h[k]
        dn[k] -= h[k]
        ref[k] = (oracle.evaluate(up)["logp"] - oracle.evaluate(dn)["logp"]) / (2 * h[k])
    scale = np.abs(ref).max()
    np.testing.assert_allclose(grad, ref, rtol=2e-3, atol=2e-4 * scale)
    np.testing.assert_allclose(post.cost_gradient(theta), -grad)


def _compare_with_oracle(post, oracle, theta, chords=(0,), min_checked=3):
    """logP, status agreement and per-chord spectra of `post` against the oracle port on the rows of `theta`."""
    from oracle.numpy_port import OracleError

    logp = post(theta)
    status = post.last_status.copy()
    spectra = {c: post.spectra_batch(theta, chord=c)[0] for c in chords}
    checked = 0
    for i, th in enumerate(theta):
        try:
            ref = oracle.evaluate(th)
        except OracleError:
            assert status[i] != 0, f"oracle raised but status is clean for theta {i}"
            continue
        assert status[i] == 0, (i, status[i])
        np.testing.assert_allclose(logp[i], ref["logp"], rtol=LOGP_RTOL)
        for c in chords:
            np.testing.assert_allclose(spectra[c][i], ref["spectra"][c], rtol=SPECTRA_RTOL)
        checked += 1
    assert checked >= min_checked
    return logp, status


def test_config3_32_chords_against_oracle(posteriors):
    """BASELINE configs[3]: one posterior with 32 spectrometer chords sharing one plasma (48 parameters): log-posterior,
    every component and the spectra of four chords against the oracle on fresh thetas near the reference theta
    (the regime the ensemble workload samples) and across the whole prior box."""
    post, wl = posteriors("multi_species_32chords")
    oracle = oracle_posterior("multi_species_32chords")
    assert len(post.chords) == 32 and post.plasma.n_params == 48
    rng = np.random.default_rng(41)
    tb = np.asarray(post.plasma.theta_bounds, dtype=float)
    np.random.seed(5)
    centre = wl.theta_from_tags(post.plasma.slices, post.random_start())
    near = centre + 1e-3 * (tb[:, 1] - tb[:, 0]) * rng.normal(size=(4, len(centre)))
    wide = tb[:, 0] + rng.random((4, len(centre))) * (tb[:, 1] - tb[:, 0])
    theta = np.concatenate([near, wide])
    _compare_with_oracle(post, oracle, theta, chords=(0, 7, 19, 31), min_checked=5)
    comps, _ = post.components_batch(theta[:2])
    ref = oracle.evaluate(theta[0])
    np.testing.assert_allclose(comps[0], ref["components"], rtol=LOGP_RTOL, atol=1e-9)


def test_config2_full_batch_sample_against_oracle(posteriors):
    """BASELINE configs[2] at its full batch (65 536 thetas of the multi-species chord): a seeded sample of 64 thetas
    against the oracle, plus the size-independent properties (determinism, batch splitting) at that size."""
    import torch

    post, wl = posteriors("multi_species")
    oracle = oracle_posterior("multi_species")
    rng = np.random.default_rng(65536)
    lo, hi = post.plasma.theta_bounds[:, 0], post.plasma.theta_bounds[:, 1]
    theta = lo + rng.random((65536, post.plasma.n_params)) * (hi - lo)
    dev = torch.as_tensor(theta, device="cuda")
    a, st = post.logp_batch(dev, return_status=True)
    b = post.logp_batch(dev)
    assert torch.equal(a, b)
    parts = torch.cat([post.logp_batch(dev[:30000]), post.logp_batch(dev[30000:])])
    assert torch.equal(a, parts)
    a, st = a.cpu().numpy(), st.cpu().numpy()
    assert (st == 0).mean() > 0.9
    # the reference-facing call with a HOST array: batches from 65 536 thetas cross the bus in chunks of 32 768 while the
    # earlier chunks are evaluated (baysar_eval_logp_host) -- same bits as the device-resident call; with a smaller chunk
    # (options.host_chunk_thetas) also with a ragged last chunk
    np.testing.assert_array_equal(post(theta), a)
    np.testing.assert_array_equal(post.last_status, st)
    from baysar_b200 import configs

    small_chunks = _fresh_posterior(configs.multi_species(), {"host_chunk_thetas": 12000})
    lo2, hi2 = small_chunks.plasma.theta_bounds[:, 0], small_chunks.plasma.theta_bounds[:, 1]
    theta2 = lo2 + rng.random((40001, small_chunks.plasma.n_params)) * (hi2 - lo2)
    np.testing.assert_array_equal(small_chunks(theta2), small_chunks.logp_batch(torch.as_tensor(theta2, device="cuda")).cpu().numpy())
    from oracle.numpy_port import OracleError

    checked = 0
    for i in rng.choice(65536, 64, replace=False):
        try:
            ref = oracle(theta[i])
        except OracleError:
            assert st[i] != 0
            continue
        assert st[i] == 0
        np.testing.assert_allclose(a[i], ref, rtol=LOGP_RTOL)
        checked += 1
    assert checked >= 56


def test_hot_balmer_chord_against_oracle(posteriors):
    """Thermalised neutrals on the Balmer-series chord: every line takes the multi-tap Stark (x) Doppler path
    (30-200 taps of the 0.1 A grid), the case SURVEY hard part (e) calls the FLOP sink."""
    post, wl = posteriors("balmer_hot")
    oracle = oracle_posterior("balmer_hot")
    rng = np.random.default_rng(77)
    lo, hi = post.plasma.theta_bounds[:, 0], post.plasma.theta_bounds[:, 1]
    theta = lo + rng.random((32, post.plasma.n_params)) * (hi - lo)
    _compare_with_oracle(post, oracle, theta, min_checked=24)


@pytest.mark.parametrize("n_los,n_lines,n_pix,n_theta", [(200, 200, 4096, 4), (1000, 200, 4096, 3), (50, 10, 4096, 4),
                                                       (100, 200, 512, 4)])
def test_shape_sweep_corners_against_oracle(n_los, n_lines, n_pix, n_theta):
    """BASELINE configs[4] corners: the largest pixel count (P = 4096, F = 12 287), the most Gaussian components
    (n_G = 200) and the longest line of sight (L = 1000) — log-posterior and spectra against the oracle."""
    import baysar_b200.lineshapes as host_ns
    import oracle.numpy_port as oracle_ns
    from baysar_b200 import configs
    from baysar_b200.atomic_data import SyntheticADAS
    from baysar_b200.input_functions import make_input_dict
    from baysar_b200.posterior import BaysarPosterior
    from oracle.numpy_port import OraclePosterior

    wl = configs.synthetic_sweep(n_los, n_lines, n_pix, balmer=True)
    np.random.seed(0)
    post = BaysarPosterior(make_input_dict(**wl.input_kwargs), profile_function=wl.make_profile(host_ns),
                           atomic_data=SyntheticADAS(), skip_test=True)
    oracle = OraclePosterior(make_input_dict(**wl.input_kwargs), SyntheticADAS(), wl.make_profile(oracle_ns))
    rng = np.random.default_rng(n_los + n_lines + n_pix)
    lo, hi = post.plasma.theta_bounds[:, 0], post.plasma.theta_bounds[:, 1]
    theta = lo + rng.random((n_theta, post.plasma.n_params)) * (hi - lo)
    _compare_with_oracle(post, oracle, theta, min_checked=max(2, n_theta - 2))


def test_torch_operator_entry_point(posteriors):
    """torch.ops.baysar.eval_logp is the same call as the C ABI's baysar_eval_logp."""
    import torch

    import baysar_b200.torch_ops  # noqa: F401  (registers the operator)

    post, wl = posteriors("balmer_series")
    z = np.load(GOLDEN / "balmer_series.npz")
    theta = torch.as_tensor(z["theta"], device="cuda")
    logp = torch.ops.baysar.eval_logp(post.op_handle, theta)
    np.testing.assert_allclose(logp.cpu().numpy(), z["logp"], rtol=LOGP_RTOL)
    both, status = torch.ops.baysar.eval_logp_status(post.op_handle, theta)
    assert torch.equal(both, logp) and int(status.abs().sum()) == 0
    with pytest.raises(NotImplementedError):
        torch.ops.baysar.eval_logp(post.op_handle, theta.cpu())  # no CPU implementation: the dispatcher refuses


def test_gaussian_likelihood_matches_reference(posteriors):
    """SpectrometerChord.likelihood(cauchy=False) (spectrometers.py:187-188): fixture values written by the reference."""
    post, _ = posteriors("multi_recombining")
    z = np.load(GOLDEN / "multi_recombining.npz")
    comps, status = post.components_batch(z["theta"], gaussian=True)
    assert (status == 0).all()
    np.testing.assert_allclose(comps[:, 0], z["c0_gauss_likelihood"], rtol=1e-5)
    np.testing.assert_allclose(comps[:, 1:], z["components"][:, 1:], rtol=LOGP_RTOL, atol=1e-9)
    post(z["theta"][0])  # scalar protocol: chord.likelihood(cauchy=False) for the posterior's last theta
    np.testing.assert_allclose(post.posterior_components[0].likelihood(cauchy=False), z["c0_gauss_likelihood"][0], rtol=1e-5)
    np.testing.assert_allclose(post.posterior_components[0].likelihood(), z["components"][0, 0], rtol=LOGP_RTOL)


def test_plasma_switches_recompile_and_ne_tau_raises():
    """Assigning plasma.recombining after construction changes the next evaluation (the model is recompiled), and
    plasma.ne_tau = True raises ValueError like the reference does on every evaluation (fixture multi_ne_tau)."""
    post, _ = gpu_posterior("multi_species")
    z0, z1 = np.load(GOLDEN / "multi_species.npz"), np.load(GOLDEN / "multi_recombining.npz")
    np.testing.assert_allclose(post(z0["theta"][:4]), z0["logp"][:4], rtol=LOGP_RTOL)
    post.plasma.recombining = True
    np.testing.assert_allclose(post(z1["theta"][:4]), z1["logp"][:4], rtol=LOGP_RTOL)
    post.plasma.recombining = False
    np.testing.assert_allclose(post(z0["theta"][:4]), z0["logp"][:4], rtol=LOGP_RTOL)
    post.plasma.ne_tau = True
    with pytest.raises(ValueError, match="truth value of an array"):
        post(z0["theta"][0])
    with pytest.raises(ValueError, match="truth value of an array"):
        post(z0["theta"][:4])


def test_user_prior_callables():
    """posterior.py:163-165, :228: any zero-argument callable in priors=[...] joins the sum; it reads the state the
    posterior publishes (plasma.plasma_state, line attributes).  Scalar and batched calls; NaN from a callable is
    reported like the reference reports a NaN posterior."""
    from baysar_b200.atomic_data import SyntheticADAS
    from baysar_b200.input_functions import make_input_dict
    import baysar_b200.lineshapes as host_ns
    from baysar_b200.posterior import BaysarPosterior

    wl = WORKLOADS["multi_species"]()
    z = np.load(GOLDEN / "multi_species.npz")
    holder = {}

    def te_peak_prior():  # what a user script writes: reads the published plasma state
        if "post" not in holder:  # (the constructor's init_test calls the components before the script has the object)
            return 0.0
        te = holder["post"].plasma.plasma_state["electron_temperature"]
        return -0.5 * ((te.max() - 5.0) / 2.0) ** 2

    def line_ratio_prior():  # reads line attributes of the first chord
        lines = holder["post"].posterior_components[0].lines
        return -1e-3 * abs(np.log10(lines[0].emission_fitted / lines[1].emission_fitted))

    np.random.seed(3)
    post = BaysarPosterior(make_input_dict(**wl.input_kwargs), profile_function=wl.make_profile(host_ns),
                           atomic_data=SyntheticADAS(), skip_test=True, priors=[te_peak_prior])
    holder["post"] = post
    post.posterior_components.append(line_ratio_prior)  # ... or appended afterwards, as scripts for the reference do
    assert post.posterior_components[-2:] == [te_peak_prior, line_ratio_prior]
    theta = z["theta"][:6]
    batched = post(theta)
    for i in range(len(theta)):
        scalar = post(theta[i])
        te = z["te"][i]
        expect = z["logp"][i] - 0.5 * ((te.max() - 5.0) / 2.0) ** 2 + line_ratio_prior()
        np.testing.assert_allclose(scalar, expect, rtol=LOGP_RTOL)
        np.testing.assert_allclose(batched[i], scalar, rtol=1e-12)
    post.posterior_components.append(lambda: float("nan"))
    with pytest.raises(ValueError, match="NaN"):
        post(theta[0])
    assert post(theta)[0] != post(theta)[0] and (post.last_status & (1 << 13)).all()
    with pytest.raises(TypeError):
        BaysarPosterior(make_input_dict(**wl.input_kwargs), profile_function=wl.make_profile(host_ns),
                        atomic_data=SyntheticADAS(), skip_test=True, priors=[3.0])


def test_synthetic_measurement_noise_model(posteriors):
    """Forward model + the reference's synthetic-spectrometer noise model (solps_spectrometer.py:464-471)."""
    post, _ = posteriors("multi_species")
    z = np.load(GOLDEN / "multi_species.npz")
    c = 1e9
    spectra, sigmas = post.synthetic_measurement(np.repeat(z["theta"][:1], 400, axis=0), c, np.random.default_rng(1))
    fm = z["c0_spectra"][0]
    np.testing.assert_allclose(sigmas[0][0], np.sqrt((1e2 * c) ** 2 + c * fm), rtol=2e-5)
    assert spectra[0].shape == (400, len(fm)) and (spectra[0] >= c).all()
    resid = (spectra[0] - fm) / sigmas[0]
    assert abs(resid.mean()) < 0.01 and abs(resid.std() - 1.0) < 0.01  # 400 x 1024 standard-normal draws
    again, _ = post.synthetic_measurement(np.repeat(z["theta"][:1], 400, axis=0), c, np.random.default_rng(1))
    np.testing.assert_array_equal(again[0], spectra[0])


def test_neutral_power_prior_raises_key_error_like_the_reference(posteriors):
    """priors.py:576-588 reads plasma_state["power"]["D_ADAS_0_exc"], a key total_power never writes (plasmas.py:1187
    stores "<species>_exc"): the reference raises KeyError on every call, and so does the mirror."""
    from baysar_b200.priors import NeutralPowerPrior

    post, _ = posteriors("multi_species")
    z = np.load(GOLDEN / "multi_species.npz")
    post(z["theta"][0])
    with pytest.raises(KeyError):
        NeutralPowerPrior(post.plasma)()
