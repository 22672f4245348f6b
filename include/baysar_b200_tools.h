/*
 * baysar_b200 tools - measurement and unit-test entry points of libbaysar_b200_tools.so.  NOT part of the drop-in
 * boundary (include/baysar_b200.h): bench.py uses the peak-rate probes for its roofline denominators, tests/ and tools/
 * use the contraction hooks.  Errors: return code != 0, message via baysar_tools_last_error().
 */
#ifndef BAYSAR_B200_TOOLS_H
#define BAYSAR_B200_TOOLS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

const char* baysar_tools_last_error(void);

/* Peak-rate probe for roofline denominators: what = 0 FP32 FMA flop/s, 1 MUFU (rcp) op/s, 2 FP64 FMA flop/s. */
int baysar_microbench(int device, int what, double* result);

/* Unit-test hook for the tcgen05 instrument-function contraction (csrc/if_contract.cuh): for n_rows (multiple of 128)
 * rows of F floats, out[r][j] = max(sum_k taps[k] * spectra[r][j + (KI-1)/2 - k], row_min[r]) (zero padded "same"
 * convolution, scipy.signal.fftconvolve semantics of spectrometers.py:269-271).  out_dev must hold
 * n_rows * ceil(F/128)*128 floats; trapz_dev (optional) n_rows * ceil(F/128) doubles. */
int baysar_debug_if_contract(const float* spectra_dev, int64_t n_rows, int F, const double* taps_host, int KI,
                             const float* row_min_dev, float* out_dev, double* trapz_dev, int64_t* out_pitch,
                             int* n_tiles, float* elapsed_ms);

/* Unit-test hook for the folded pixel contraction (csrc/pix_contract.cuh): instrument convolution and the static
 * wavelength map (interp1d between fine-grid points pix_idx[p], pix_idx[p] + 1 with weight pix_frac[p]) as one
 * contraction.  out[r][c]: column pix_col[p] (>= 0) holds pixel p; explicit (unclipped) fine-grid columns
 * j = 0 .. x_l - 1 at col_left0 + j and j = F - x_r .. F - 1 at col_right0 + j - (F - x_r).
 * info[8] = out_pitch, col_left0, x_l, col_right0, x_r, column tiles, k-blocks per theta block, A ring stages.
 * tsum_dev (optional, n_rows * column tiles * 2 doubles): exact column sums per (row, folded tile, converter group).
 * out_dev == NULL: layout query only. */
int baysar_debug_pix_contract(const float* spectra_dev, int64_t n_rows, int F, const double* taps_host, int KI,
                              const int* pix_idx_host, const double* pix_frac_host, int P, int tile_n, float* out_dev,
                              int64_t out_capacity, double* tsum_dev, int64_t tsum_capacity, int64_t* info,
                              int* pix_col_host, float* elapsed_ms);

/* Host-only self check of the folded-contraction plan (runs without a GPU): one float64 spectrum through the plan's
 * tables with the kernel's index arithmetic.  result[9] = max |error| / max|conv| of the pixel columns, of the explicit
 * edge columns, relative error of the area identity, k-blocks per theta block, folded pixels, largest interior clip
 * violation / max|conv| (<= 0: the clip is inactive there), tap-table floats, A ring stages, relative error of the
 * column-sum identity (float32 bbar table against the float64 column sum). */
int baysar_debug_pix_plan_check(const double* spectrum_host, int F, const double* taps_host, int KI,
                                const int* pix_idx_host, const double* pix_frac_host, int P, int tile_n, double* result);

/* Issue-rate probe of kind::tf32 tcgen05.mma (csrc/mma_probe.cuh): cycles per 128 x tile_n x 8 MMA, slowest SM, for the
 * contraction kernels' issue pattern on constant operands.  form 0 = A from tensor memory, 1 = A from shared memory;
 * pattern 0 = kernels' pattern, 1 = one accumulator, 2 = hi*hi only. */
int baysar_debug_mma_probe(int device, int tile_n, int form, int n_acc_hi, int b_stages, int pattern, int iters,
                           double* cycles_per_mma);


#ifdef __cplusplus
}
#endif
#endif /* BAYSAR_B200_TOOLS_H */
