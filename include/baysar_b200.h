/*
 * baysar_b200 — C ABI of the B200-native BaySAR posterior-evaluation path.
 *
 * The reference (dgahle/baysar) is pure Python and has no FFI: its boundary for this path is the duck-typed call
 * `BaysarPosterior.__call__(theta) -> float` (reference baysar/posterior.py:219-235) plus the per-chord
 * `SpectrometerChord.forward_model()` (baysar/spectrometers.py:197-335) and the per-line diagnostic attributes
 * read by priors and demos (baysar/linemodels.py:800-836, :407-426).  Each entry point below replaces one of those
 * calls for a whole batch of thetas; the Python host side (`baysar_b200/posterior.py`) binds them with ctypes.
 *
 * Conventions
 *   - plain pointers and sizes only; every `theta`, output and status pointer is a DEVICE pointer unless the
 *     function name ends in `_host`;
 *   - the caller owns all buffers; the model owns its device constants and a workspace that grows on demand;
 *   - return value 0 = ok, non-zero = error (message via baysar_last_error(), thread-local); nothing throws;
 *   - per-sample numerical failures (the reference's `raise` sites) are reported as bit flags in `status[N]`
 *     (BAYSAR_ST_*), never as a return code; logp of a failed sample is whatever arithmetic produced (NaN/inf);
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); calls are asynchronous w.r.t.
 *     the host except the `_host` variants;
 *   - a model's constants are immutable after creation: evaluation is re-entrant across models; one model owns one
 *     workspace, so calls on the same model must be ordered on one stream (or separated by events).
 */
#ifndef BAYSAR_B200_H
#define BAYSAR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct baysar_model baysar_model_t;

/* status bits: one per reference raise site (file:line of the reference in the comment) */
#define BAYSAR_ST_THETA_NAN        (1u << 0)  /* posterior.py:239-240  TypeError  */
#define BAYSAR_ST_THETA_INF        (1u << 1)  /* posterior.py:242-243  TypeError  */
#define BAYSAR_ST_TAU_RANGE        (1u << 2)  /* plasmas.py:1410-1429, linemodels.py:454-467  ValueError/IndexError */
#define BAYSAR_ST_N0_NAN           (1u << 3)  /* linemodels.py:768-769  ValueError */
#define BAYSAR_ST_PEC_NAN          (1u << 4)  /* linemodels.py:790-796  ValueError */
#define BAYSAR_ST_FWHM_NONPOS      (1u << 5)  /* lineshapes.py:97-103   ValueError */
#define BAYSAR_ST_LINE_NAN         (1u << 6)  /* linemodels.py:430-441, :860-863  TypeError/ValueError */
#define BAYSAR_ST_IF_NORM          (1u << 7)  /* spectrometers.py:246-250  ValueError */
#define BAYSAR_ST_CONV_NANINF      (1u << 8)  /* spectrometers.py:279-282  TypeError  */
#define BAYSAR_ST_AREA             (1u << 9)  /* spectrometers.py:292-297  ValueError */
#define BAYSAR_ST_SPECTRA_NANINF   (1u << 10) /* spectrometers.py:330-333  TypeError  */
#define BAYSAR_ST_RES_ZERO         (1u << 11) /* spectrometers.py:169-170  ValueError */
#define BAYSAR_ST_RES_NANINF       (1u << 12) /* spectrometers.py:172-183  ValueError */
#define BAYSAR_ST_LOGP_NAN         (1u << 13) /* posterior.py:250-252  ValueError */
#define BAYSAR_ST_LOGP_INF         (1u << 14) /* posterior.py:253-255  ValueError */
#define BAYSAR_ST_LOGP_POSITIVE    (1u << 15) /* posterior.py:256-258  ValueError */
#define BAYSAR_ST_WINDOW_CLAMPED   (1u << 16) /* not a reference error: a theta-dependent kernel was wider than
                                                 the shared-memory halo sized from the parameter bounds */

/* number of doubles per (theta, line) record returned by baysar_eval_lines */
#define BAYSAR_LINE_STRIDE 16
/* record layout, by line kind:
 *   Balmer (kind 0): rec_sum, exc_sum, rec_ne, rec_te, exc_ne, exc_te, ti, f_rec, ems_ne, ems_te, emission_fitted,
 *                    gamma_rec^2.5, gamma_exc^2.5 (Stark half-widths), Doppler-shifted centre, Doppler sigma, 0
 *   ADAS406 (kind 1): emission_fitted, ems_ne, ems_conc, ems_te, ti, 0...
 *   XLine (kind 2):   emission_fitted, 0...                                                                       */

/* Build a model from the descriptor blob produced by baysar_b200.model.compile_model (format: csrc/layout.h);
 * copies every constant table to HBM on `device` once.   Replaces BaysarPosterior.__init__'s object graph. */
int baysar_model_create(const void* desc_blob, size_t nbytes, int device, baysar_model_t** out);

/* Tuning options of a model (every field: 0 = the library's default).  They select between kernels that compute the same
 * quantities (parity tests compare the variants); none is read from the environment. */
typedef struct baysar_options {
  int tensor_cores;      /* 0 auto (tcgen05 contraction for long static instrument functions), 1 force on, -1 off:
                            every chord stays on the fused FP32 kernel */
  int ifc_min_taps;      /* shortest significant instrument-function window that takes the tensor-core path (160) */
  int ifc_budget_mb;     /* HBM budget of the tensor-core path's spectra workspaces in MiB (24576) */
  int folded_tile;       /* folded pixel contraction: 0 auto (128-column tiles), 64 / 128 tile width, -1 off (the
                            every-column Toeplitz contraction) */
  int los_min_blocks;    /* resident CTAs per SM the LOS kernel is compiled for (0 auto by batch size, 4..8) */
  int los_team_points;   /* LOS points from which one CTA works on one theta (0 auto: when 4 thetas no longer fit) */
  int chord_min_blocks;  /* upper limit of resident CTAs per SM of the chord kernel (0 auto = 5; 3..5) */
  int chord_threads;     /* threads per CTA (= per theta and chord) of the chord kernel: 0 auto, 128 or 256 */
  int los_team_warps;    /* warps that share one theta in the LOS kernel: 0 auto by line-of-sight length, or 2, 4, 8 */
  int serial_launches;   /* 1: plain kernel launches instead of programmatic dependent launch (A/B timing) */
  int contraction_b_image; /* folded contraction's B operand: 0 auto (a precomputed image in HBM / L2, copied with
                            cp.async.bulk), -1 generated by builder warps from the tap table in shared memory */
  int host_chunk_thetas; /* baysar_eval_logp_host: thetas per chunk of the pipelined host-to-device copy (0 auto = 32768;
                            batches below two chunks are copied in one piece; -1: never chunk) */
  int reserved[4];
} baysar_options_t;
int baysar_model_create_ex(const void* desc_blob, size_t nbytes, int device, const baysar_options_t* options,
                           baysar_model_t** out);
void baysar_model_destroy(baysar_model_t* model);

/* Sizes a caller needs to allocate outputs. what: 0 n_params, 1 n_chords, 2 n_unique_lines, 3 n_priors,
 * 4 max pixels per chord, 5 max fine-grid points per chord, 6 pixels of chord `arg`, 7 fine points of chord `arg`,
 * 8 LOS points, 9 kernels launched per baysar_eval_logp call, 11 any chord on the tensor-core path, 12 k-blocks per
 * theta block of chord `arg`'s folded pixel contraction (0: not folded), 13 its column tile width. */
int64_t baysar_model_query(const baysar_model_t* model, int what, int arg);

/* logp[n] = BaysarPosterior.__call__(theta[n])  (posterior.py:219-235): theta [N][n_params] row-major doubles.
 * Optional outputs (may be NULL): status[N]; components [N][n_chords + n_priors] in the reference's
 * posterior_components order (chord log-likelihoods, then priors). */
int baysar_eval_logp(baysar_model_t* model, const double* theta, int64_t n, double* logp, uint32_t* status,
                     double* components, void* stream);

/* spectra[n][p] = SpectrometerChord.forward_model() of chord `chord` (spectrometers.py:197-335), P doubles per
 * theta.  fine_pre / fine_post (may be NULL): the fine-grid spectrum before the instrument convolution (after the
 * intensity calibration) and after convolution + renormalisation + background, F floats per theta. */
int baysar_eval_spectra(baysar_model_t* model, const double* theta, int64_t n, int chord, double* spectra,
                        float* fine_pre, float* fine_post, uint32_t* status, void* stream);

/* lines [N][n_unique_lines][BAYSAR_LINE_STRIDE]: the line-of-sight stage scalars the reference exposes as line
 * attributes (linemodels.py:800-836, :407-426); profiles (may be NULL) [N][4][L]: ne, Te, main-ion density, n0. */
int baysar_eval_lines(baysar_model_t* model, const double* theta, int64_t n, double* lines, double* profiles,
                      uint32_t* status, void* stream);

/* Convenience for callers with HOST buffers (what a scalar `posterior(theta)` call or an emcee vectorised
 * call hands over): copies theta H2D, evaluates, copies logp/status D2H, synchronises. */
int baysar_eval_logp_host(baysar_model_t* model, const double* theta_host, int64_t n, double* logp_host,
                          uint32_t* status_host);

/* Measurement hooks (no reference counterpart; used by bench.py for the live roofline).  While profiling is on,
 * every baysar_eval_logp call records CUDA events around each of its kernels on the caller's stream (the events come
 * from a pool that grows during the first profiled calls and is reused afterwards: no event is created inside a
 * steady-state call); baysar_model_get_profile synchronises, returns the summed per-stage milliseconds
 * (LOS, chord, contraction, pixel, finalize) and the number of calls, and resets the log. */
int baysar_model_set_profiling(baysar_model_t* model, int on);
int baysar_model_get_profile(baysar_model_t* model, double* stage_ms, int64_t* calls);

const char* baysar_last_error(void);
const char* baysar_version(void);

/*
 * Affine-invariant ensemble step on the device (stretch move, red/blue split) — the caller of the batched posterior
 * for large walker counts; replaces what emcee's EnsembleSampler does on the host between two posterior calls
 * (reference caller: tulasa/fitting.py:1275-1304).  Philox4x32-10, counter (walker, step, draw), key = seed.
 *   propose: proposal[s] = c + z (x_s - c), z = ((a - 1) u + 1)^2 / a, c uniform in `complement`;
 *            log_factor[s] = (n_params - 1) ln z
 *   accept:  x_s, logp[s] <- proposal, logp_new[s] if ln u' < log_factor + logp_new - logp and the proposal evaluated
 *            cleanly (status_new[s] == 0, finite logp_new); n_accepted (device counter, may be NULL) += acceptances
 * All pointers are device pointers.
 */
int baysar_stretch_propose(const double* walkers, const double* complement, int n_params, int64_t n_walkers,
                           int64_t n_complement, double a, uint64_t seed, uint64_t step, int64_t walker_offset,
                           double* proposal, double* log_factor, void* stream);
int baysar_stretch_accept(double* walkers, double* logp, const double* proposal, const double* logp_new,
                          const double* log_factor, const uint32_t* status_new, int n_params, int64_t n_walkers,
                          uint64_t seed, uint64_t step, int64_t walker_offset, unsigned long long* n_accepted, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BAYSAR_B200_H */
