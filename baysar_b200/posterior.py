"""``BaysarPosterior`` — the drop-in boundary of the posterior-evaluation path.

Same constructor and call protocol as the reference (`/root/reference/baysar/posterior.py:57-258`):

    posterior = BaysarPosterior(input_dict, profile_function=None, priors=[], check_bounds=False, temper=1.0,
                                curvature=None, print_errors=False, skip_test=False)
    logp = posterior(theta)          # 1-D theta -> float, raising TypeError / ValueError like the reference
    cost = posterior.cost(theta)

plus the batched form the reference lacks: ``posterior(theta)`` with a 2-D ``(N, n_params)`` array (numpy, host) or
CUDA tensor returns N log-posteriors evaluated by three CUDA kernel launches; per-sample failures come back as status
flags (``posterior.last_status``) instead of exceptions.  Extra keyword arguments: ``atomic_data`` (provider, see
``baysar_b200.atomic_data``), ``device`` and ``options`` (dict of ``baysar_options_t`` fields, include/baysar_b200.h:
kernel-selection switches such as ``tensor_cores=-1``; every variant computes the same quantities).

There is no CPU path: construction needs a CUDA device and the in-tree ``libbaysar_b200.so``.
"""
import numpy as np

from . import runtime
from .model import compile_model
from .plasmas import PlasmaLine
from .priors import (AntiprotonCost, CalibrationPrior, ChargeStateOrderPrior, CurvatureCost, GaussianInstrumentFunctionPrior,
                     MainIonFractionCost, NeutralFractionCost, StarkToPeakPrior, StaticElectronPressureCost,
                     WavelengthCalibrationPrior, _DevicePrior)
from .spectrometers import BalmerHydrogenLine, SpectrometerChord, XLine

# status bit -> (exception type, message) in the order the reference would hit them (include/baysar_b200.h)
_STATUS = [
    (0, TypeError, "Theta contains NaNs"), (1, TypeError, "Theta contains infinities"),
    (2, ValueError, "impurity tau outside the ADAS tau grid"),
    (3, ValueError, "Negative numbers in neutral density profile!"), (4, ValueError, "NaNs in the PECs!"),
    (5, ValueError, "fwhm must be a positive scalar"), (6, TypeError, "NaNs/infs in line peaks"),
    (7, ValueError, "Instrument function is not pixel normalised!"), (8, TypeError, "spectra contains NaNs/infs"),
    (9, ValueError, "Area not conserved in convolution!"), (10, TypeError, "spectra contains NaNs/infs"),
    (11, ValueError, "Some points of square_normalised_residuals are zero"),
    (12, ValueError, "Some points of square_normalised_residuals are inf/nan"),
    (13, ValueError, "logP is NaNs"), (14, ValueError, "logP is inf"), (15, ValueError, "lopP is positive"),
]


def build_components(plasma, input_dict, curvature=None, passed_priors=()):
    """Chords, then the default priors in the reference's order (posterior.py:263-269, :108-165)."""
    chords = [SpectrometerChord(plasma=plasma, refine=refine, chord_number=c)
              for c, refine in enumerate(input_dict["refine"])]
    priors = []
    if curvature is not None:  # posterior.py:109-110
        priors.append(CurvatureCost(plasma, curvature))
    if plasma.calibrate_wavelength:
        for num in range(plasma.num_chords):
            inmean = chords[num].x_data.mean()
            indisp = np.diff(chords[0].x_data).mean()
            priors.append(WavelengthCalibrationPrior([inmean, indisp], [2 * indisp, indisp * 0.05]))
    if plasma.calibrate_intensity:
        priors.append(CalibrationPrior())
    if plasma.calibrate_instrument_function:
        priors.append(GaussianInstrumentFunctionPrior())
    priors.append(StaticElectronPressureCost())
    if plasma.contains_hydrogen:
        priors.append(NeutralFractionCost(species=plasma.hydrogen_species[0]))
        balmer = [l for ch in chords for l in ch.lines if type(l) == BalmerHydrogenLine]
        if balmer:
            priors.append(StarkToPeakPrior(line=balmer[0]))
    if len(plasma.impurities) > 0:
        priors.append(AntiprotonCost())
        priors.append(MainIonFractionCost())
        priors.append(ChargeStateOrderPrior(chords, plasma))
    for p in passed_priors:
        # posterior.py:163-165: `got_p(p)` compares types with the INSTANCE p, so a passed prior is always appended.
        # Priors of this package's classes are rows of the device prior table; any other zero-argument callable is a
        # HOST prior, called like the reference calls it (posterior.py:228) after the plasma state and the line
        # attributes of the theta at hand have been published (see BaysarPosterior._host_prior_sum).
        if not isinstance(p, _DevicePrior) and not callable(p):
            raise TypeError("priors must be callables: `sum(p() for p in posterior_components)` (posterior.py:228)")
        priors.append(p)
    return chords, priors


def set_sensible_bounds(plasma, chords, dcwl=2, ddisp=0.005):
    """Data-driven theta bounds (reference posterior.py:271-356)."""
    tb, sl = plasma.theta_bounds, plasma.slices
    for c, chord in enumerate(chords):
        if f"cal_{c}" in sl:
            tb[sl[f"cal_{c}"]] = [-0.3, 0.3] if chord.y_data.max() > 1e6 else [8, 15]
        if f"calwave_{c}" in sl:
            cwl, disp = chord.x_data.mean(), np.diff(chord.x_data).mean()
            tb[sl[f"calwave_{c}"]][0] = [cwl - dcwl, cwl + dcwl]
            tb[sl[f"calwave_{c}"]][1] = [disp - ddisp, disp + ddisp]
        level, spread = chord.y_data_continuum.mean(), chord.y_data_continuum.std()
        lo, hi = max(level - spread, level / 10), min(level + spread, level * 10)
        tb[sl[f"background_{c}"]] = [np.log10(lo), np.log10(hi)]
    for tag in sl:
        if tag.endswith("_dens"):
            charge = float(tag.split("_")[-2])
            top = tb[sl["electron_density"]][0, 1]
            if charge != 0:
                top, decades = top - np.log10(charge), 2
            else:
                decades = 3
            tb[sl[tag]][0] = [top - decades, top]
        if not plasma.thermalised and tag.endswith("_Ti"):
            tb[sl[tag]][0] = [0, 1.7]
    for chord in chords:
        for line in chord.lines:
            if type(line) == XLine:
                line.estimate_ems_and_bounds(chord.x_data, chord.y_data)
                tb[sl[line.line_tag]] = line.bounds
    if np.isnan(tb).any():
        raise ValueError("NaNs in theta bounds!")


class BaysarPosterior:
    def __init__(self, input_dict, profile_function=None, priors=[], check_bounds=False, temper=1.0, curvature=None,
                 print_errors=False, skip_test=False, atomic_data=None, device=None, options=None):
        self.check_init_inputs(priors, check_bounds, temper, curvature, print_errors)
        self.input_dict = input_dict
        self.plasma = PlasmaLine(input_dict=input_dict, profile_function=profile_function, atomic_data=atomic_data)
        self.check_bounds, self.print_errors, self.temper = check_bounds, print_errors, temper
        self.passed_priors, self.curvature = priors, curvature
        self.nan_thetas, self.inf_thetas, self.positive_thetas, self.runtimes = [], [], [], []
        self.last_proposal = None
        self.last_status = None

        self.chords, self.priors = build_components(self.plasma, input_dict, curvature, priors)
        self.posterior_components = list(self.chords) + list(self.priors)
        self.device_priors = [p for p in self.priors if isinstance(p, _DevicePrior)]
        self.host_priors = [p for p in self.priors if not isinstance(p, _DevicePrior)]
        set_sensible_bounds(self.plasma, self.chords)

        if device is None:
            import torch

            device = torch.cuda.current_device() if torch.cuda.is_available() else 0
        self.device = int(device)
        self._options = options
        self._models = {}  # gaussian_likelihood flag -> (blob, info, DeviceModel)
        self._compile(False)  # raises without a GPU / library: no CPU fallback
        for c in self.chords:
            c._posterior = self
        for k, p in enumerate(self.device_priors):
            p._posterior, p._column = self, len(self.chords) + k
        self.plasma._evaluator = self._evaluate_los
        self.plasma._on_option_change = self._models.clear  # recombining / reverse_electroneutrality / ne_tau: recompile
        self.init_test(skip_test)

    # ------------------------------------------------------------------------------------------ compiled device models
    def _compile(self, gaussian):
        """Compile (once per state of the plasma's switches) the descriptor blob and create its device model."""
        if gaussian not in self._models:
            if getattr(self.plasma, "ne_tau", False) and len(self.plasma.impurity_species) > 0 and len(self.plasma.los) > 1:
                # linemodels.py:398-399 turns tau into an array over the line of sight and get_tec (:454) then takes the
                # truth value of an array: the reference raises this on every evaluation of an impurity line
                raise ValueError("The truth value of an array with more than one element is ambiguous. "
                                 "Use a.any() or a.all()")
            blob, info = compile_model(self.plasma, self.chords, self.device_priors, gaussian_likelihood=gaussian)
            self._models[gaussian] = (blob, info, runtime.DeviceModel(blob, self.device, self._options))
        return self._models[gaussian]

    @property
    def model(self):
        return self._compile(False)[2]

    @property
    def blob(self):
        return self._compile(False)[0]

    @property
    def model_info(self):
        return self._compile(False)[1]

    # ------------------------------------------------------------------------------------------ reference protocol
    def check_init_inputs(self, priors, check_bounds, temper, curvature, print_errors):
        """posterior.py:167-195."""
        if len(priors) > 0 and type(priors) is not list:
            raise TypeError("If priors are being passed then it must be a list of functions")
        if type(check_bounds) is not bool:
            raise TypeError("check_bounds must be True or False and is False by default")
        if not np.isreal(temper):
            raise TypeError("temper must be a real number")
        if temper < 0:
            raise ValueError("temper must be positive (should, but doesn't have to, be greater than 1)")
        if curvature is not None:
            if type(curvature) not in (int, float, np.int64, np.float64):
                raise TypeError("curvature must be an int or float")
            if curvature < 1:
                raise ValueError("curvature must be greater than 1")
        if type(print_errors) is not bool:
            raise TypeError("print_errors must be True or False and is False by default")

    def got_p(self, comp):
        return any(type(p) is comp for p in self.posterior_components)

    def init_test(self, skip_test):
        """One evaluation at a random start must give a finite negative logP (posterior.py:197-217)."""
        start = self.random_start()
        run_check = self(start, skip_test)
        if not (np.isreal(run_check) and -1e50 < run_check and run_check < 0):
            err = ValueError("Posterior is not evalauating correctly. logP =", run_check, "from input=", start)
            if skip_test:
                print(err)
            else:
                raise err

    def __call__(self, theta, skip_error=False):
        import torch

        self._sync_components()
        if isinstance(theta, torch.Tensor):
            if theta.dim() == 2 and not self.host_priors:
                return self.logp_batch(theta)
            device = theta.device
            theta = theta.detach().cpu().numpy()
            if theta.ndim == 2:  # user prior callables run on the host (see _add_host_priors)
                return torch.as_tensor(self(theta), device=device)
        arr = np.asarray(theta, dtype=float)
        if arr.ndim == 2:
            logp, status = self.model.eval_logp_host(arr)
            if self.host_priors:
                logp, status = self._add_host_priors(arr, logp, status)
            self.last_status = status
            return logp
        theta = list(arr)
        self.last_proposal = theta
        if not skip_error:
            self.check_call_input(theta, skip_error)
        logp, status = self.model.eval_logp_host(arr[None, :])
        prob = float(logp[0])
        self._publish_line_attributes(arr)
        if self.host_priors:
            prob, status = self._finish_host(prob + self._host_prior_sum(), status)
        self.last_status = status
        if not skip_error:
            self._raise_from_status(int(status[0]), theta=theta)
        self.last_proposal_logp = prob
        return prob

    # ------------------------------------------------------------------------------------------ user prior callables
    def _sync_components(self):
        """``posterior_components`` is the live list the reference sums on every call (posterior.py:228), and user scripts
        append priors to it after construction.  Re-derive the device prior table (recompiling when it changed) and
        the host callables from it."""
        tail = self.posterior_components[len(self.chords):]
        dev = [p for p in tail if isinstance(p, _DevicePrior)]
        if len(dev) != len(self.device_priors) or any(a is not b for a, b in zip(dev, self.device_priors)):
            self.device_priors = dev
            for k, p in enumerate(dev):
                p._posterior, p._column = self, len(self.chords) + k
            self._models.clear()
        self.host_priors = [p for p in tail if not isinstance(p, _DevicePrior)]
        self.priors = list(tail)

    def _host_prior_sum(self):
        """Sum of the user-supplied prior callables for the state just published (reference posterior.py:228)."""
        return sum(p() for p in self.host_priors)

    @staticmethod
    def _finish_host(prob, status):
        """The output checks of posterior.py:249-258 on a total that includes host priors (bits 13-15 of the status)."""
        status = np.array(status, dtype=np.uint32, copy=True)
        status[0] &= ~np.uint32(0xE000)
        if np.isnan(prob):
            status[0] |= 1 << 13
        elif np.isinf(prob):
            status[0] |= 1 << 14
        elif prob >= 0:
            status[0] |= 1 << 15
        return prob, status

    def _add_host_priors(self, thetas, logp, status):
        """Batched call with user prior callables.  The device evaluates chords and device priors for the whole batch in
        one launch; the callables are Python and read ``plasma.plasma_state`` / line attributes, so they are called
        once per theta on the host after that theta's state (one batched LOS launch for all of them) is published.
        A batch with host priors is therefore host-bound: use the package's prior classes where throughput matters."""
        logp = np.array(logp, dtype=float, copy=True)
        status = np.array(status, dtype=np.uint32, copy=True)
        lines, _, prof = self.lines_batch(thetas, want_profiles=True)
        for n, theta in enumerate(thetas):
            out = dict(ne=prof[n, 0], te=prof[n, 1], main_ion_density=prof[n, 2], n0=prof[n, 3], lines=lines[n])
            self._publish_line_attributes(theta, out)
            logp[n], status[n:n + 1] = self._finish_host(logp[n] + self._host_prior_sum(), status[n:n + 1])
        return logp, status

    def check_call_input(self, theta, skip_error):
        if any(np.isnan(theta)):
            raise TypeError("Theta contains NaNs")
        if any(np.isinf(theta)):
            raise TypeError("Theta contains infinities")
        if skip_error not in (True, False):
            raise TypeError("skip_error not in (True, False)")

    def _raise_from_status(self, status, theta=None, mask=0xFFFF):
        status &= mask
        if not status:
            return
        for bit, exc, msg in _STATUS:
            if status & (1 << bit):
                if theta is not None and bit in (13, 14, 15):
                    {13: self.nan_thetas, 14: self.inf_thetas, 15: self.positive_thetas}[bit].append(theta)
                raise exc(msg)

    def cost(self, theta):
        return -self.__call__(theta)

    def random_start(self, order=1, flat=False):
        """Uniform draw inside theta_bounds, background re-drawn near the measured continuum (posterior.py:358-395).
        Consumes the global numpy RNG in the same order as the reference."""
        start = [np.mean(np.random.uniform(b[0], b[1], size=order)) for b in self.plasma.theta_bounds]
        if flat:
            for param in ("electron_density", "electron_temperature"):
                s = self.plasma.slices[param]
                for index in np.arange(s.start, s.stop):
                    start[index] = start[s.start]
        for chord in self.chords:
            mean = np.log10(np.mean(chord.y_data_continuum))
            start[self.plasma.slices["background_" + str(chord.chord_number)].start] = \
                np.random.uniform(mean - 0.1, mean + 0.1)
        return np.array(start)

    def random_sample(self, number=1, order=1, flat=False):
        """N independent evaluations sorted by logP (posterior.py:463-472) — one batched launch here."""
        thetas = np.array([self.random_start(order, flat) for _ in range(number)])
        logp = self(thetas)
        keep = np.isfinite(logp) & (self.last_status == 0)
        order_ = np.argsort(-logp[keep])
        return thetas[keep][order_], logp[keep][order_]

    # ------------------------------------------------------------------------------------------ batched device API
    def _as_device(self, theta):
        import torch

        if isinstance(theta, torch.Tensor):
            return theta.to(device=f"cuda:{self.device}", dtype=torch.float64)
        return torch.as_tensor(np.asarray(theta, dtype=float), device=f"cuda:{self.device}")

    @property
    def op_handle(self):
        """Integer handle of this posterior's compiled model for ``torch.ops.baysar.eval_logp(handle, theta)``."""
        from . import torch_ops

        return torch_ops.register_model(self.model)

    def logp_batch(self, theta, return_status=False):
        """theta: CUDA float64 tensor (N, n_params) -> CUDA tensor of N log-posteriors (no host round trip)."""
        logp, status, _ = self.model.eval_logp(self._as_device(theta))
        self.last_status_device = status
        return (logp, status) if return_status else logp

    def gradient(self, theta, step=None):
        """Finite-difference gradient of logP at one theta from ONE batched launch of 2 n_params evaluations.

        The reference's optimisers obtain gradients from scipy's ``approx_grad=True`` (`posterior.py:492-498`,
        `optimisation.py:84-101`): n + 1 sequential posterior calls per gradient.  Here the shifted thetas form one
        batch.  Central differences; ``step`` is per parameter (scalar or vector), default 1e-3 of each parameter's
        bound range — the spectral stage computes in float32, so scipy's forward step of 1e-8 would only measure
        rounding noise.  Samples whose evaluation raises in the reference (status != 0) give NaN components.
        """
        theta = np.asarray(theta, dtype=float)
        n = self.plasma.n_params
        if theta.shape != (n,):
            raise ValueError("len(theta)!=self.n_params")
        tb = np.asarray(self.plasma.theta_bounds, dtype=float)
        h = 1e-3 * (tb[:, 1] - tb[:, 0]) if step is None else np.broadcast_to(np.asarray(step, dtype=float), (n,)).copy()
        batch = np.repeat(theta[None], 2 * n, axis=0)
        batch[np.arange(n), np.arange(n)] += h
        batch[n + np.arange(n), np.arange(n)] -= h
        logp, status = self.logp_batch(batch, return_status=True)
        logp = logp.cpu().numpy()
        bad = status.cpu().numpy() != 0
        logp[bad] = np.nan
        return (logp[:n] - logp[n:]) / (2 * h)

    def cost_gradient(self, theta, step=None):
        """``-gradient``: what `baysar/optimisation.py:84-92` passes to L-BFGS-B as ``fprime`` when it exists."""
        return -self.gradient(theta, step)

    def components_batch(self, theta, gaussian=False):
        """-> (components [N, n_chords + n_device_priors], status) as numpy arrays: the chords, then the device priors in
        posterior_components order (host prior callables have no column).  ``gaussian``: the chord columns hold
        ``likelihood(cauchy=False)`` (a second device model, compiled on first use)."""
        _, status, comps = self._compile(bool(gaussian))[2].eval_logp(self._as_device(theta), want_components=True)
        return comps.cpu().numpy(), status.cpu().numpy().astype(np.uint32)

    def spectra_batch(self, theta, chord=0, want_fine=False):
        spectra, status, pre, post = self.model.eval_spectra(self._as_device(theta), chord, want_fine)
        if want_fine:
            return spectra.cpu().numpy(), status.cpu().numpy().astype(np.uint32), pre.cpu().numpy(), post.cpu().numpy()
        return spectra.cpu().numpy(), status.cpu().numpy().astype(np.uint32)

    def synthetic_measurement(self, theta, calibration_constant=1e9, rng=None):
        """Synthetic measured spectra for one theta (or a batch): the forward model of every chord plus the noise model of
        the reference's synthetic spectrometer (`solps_spectrometer.py:464-471`): sigma^2 = (1e2 c)^2 + c * spectrum,
        normal noise, floored at the calibration constant c.  -> (list of [N, P_chord] spectra, list of their sigmas).
        (The reference builds its synthetic spectra from SOLPS grids, which are outside this path; this applies the same
        measurement model to the posterior's own forward model, e.g. to make self-consistent test data on the GPU.)"""
        rng = np.random.default_rng() if rng is None else rng
        thetas = np.atleast_2d(np.asarray(theta, dtype=float))
        spectra, sigmas = [], []
        for c in range(len(self.chords)):
            fm, status = self.spectra_batch(thetas, c)
            if (status & 0x7FF).any():
                raise ValueError("the forward model failed for at least one theta (status flags set)")
            sigma = np.sqrt(np.square(1e2 * calibration_constant) + calibration_constant * fm)
            spectra.append((fm + rng.normal(0.0, sigma)).clip(calibration_constant))
            sigmas.append(sigma)
        return spectra, sigmas

    def lines_batch(self, theta, want_profiles=False):
        lines, status, prof = self.model.eval_lines(self._as_device(theta), want_profiles)
        out = lines.cpu().numpy(), status.cpu().numpy().astype(np.uint32)
        return out + (prof.cpu().numpy(),) if want_profiles else out

    def _evaluate_los(self, theta):
        lines, _, prof = self.lines_batch(np.asarray(theta, dtype=float)[None], want_profiles=True)
        return dict(ne=prof[0, 0], te=prof[0, 1], main_ion_density=prof[0, 2], n0=prof[0, 3], lines=lines[0])

    def _publish_line_attributes(self, theta, out=None):
        """After a scalar call, expose the reference's line attributes and plasma_state (priors/demos read them)."""
        out = self.plasma.update_plasma_state(theta, out)
        for c, chord in enumerate(self.chords):
            for k, line in enumerate(chord.lines):
                line._publish(out["lines"][self.model_info["chord_line_uline"][c][k]])
