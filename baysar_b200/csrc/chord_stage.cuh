// K2 — spectral stage: one CTA per (theta, chord).  float32 element-wise work on the fine wavelength grid held in
// shared memory, float64 for tap tables, reductions and the pixel / likelihood epilogue.
//
// Replaces, per (theta, chord) (reference file:line):
//   BalmerHydrogenLine.__call__ (shape part)   baysar/linemodels.py:838-865
//   HydrogenLineShape.__call__, stehle_param   baysar/linemodels.py:688-723, :554-566
//   DopplerLine / Gaussian.make_peak           baysar/linemodels.py:283-300, baysar/lineshapes.py:126-131,206-217
//   ADAS406Lines.__call__ (shape part), XLine  baysar/linemodels.py:401-403,428 ; :213-233
//   SpectrometerChord.forward_model            baysar/spectrometers.py:197-335
//   SpectrometerChord.likelihood               baysar/spectrometers.py:154-195
//
// Algorithmic departures from the reference's arithmetic (validated on CPU in tests/test_kernel_model.py):
//   * direct "same" convolutions, out[j] = sum_k w[k] s[j + (K-1)/2 - k], instead of FFT convolutions;
//   * taps / Gaussian samples with exp(-z^2/2) < 1e-17 (|tap| < 1e-17 max) are dropped;
//   * the two trapz normalisations of the Stark (x) Doppler profile collapse into A = sum_i u[i] G[i], with G the
//     trapz weights pulled back through the convolution (constant inside, tap prefix sums at the two ends);
//   * uniform trapz weights on the arange grid; the four end outputs of the instrument convolution in float64.
//
// Shared-memory layout: two zero-haloed fine-grid buffers A (accumulated spectrum) and B (scratch / convolved
// spectrum), XOR-swizzled at float4 granularity: element i lives at i ^ ((i >> 3) & 4), i.e. odd 32-element blocks
// swap their float4 pairs.  Threads that own 8 consecutive convolution outputs read float4s at a 32-byte stride; the
// swizzle spreads a quarter-warp's eight reads over all eight 16-byte bank groups for every alignment (no padding).
#pragma once
#include "device_common.cuh"

// Doppler taps are kept while exp(-z^2/2) >= 4.6e-11 of the central tap: next to a Stark profile the neglected taps
// change a fine-grid sample by < 1.5e-8 relative (adjacent samples differ by at most ~10^2.5), below float32 resolution.
#define BAYSAR_Z_CUT_DOPPLER 6.9
#define BAYSAR_MAX_FUSED_LINES 8    // Balmer lines per pair of fused sweeps
#define BAYSAR_MAX_BALMER_LINES 16  // Balmer lines per chord (shape records in shared memory; checked by the model compiler)

struct ChordArgs {
  const double* theta;      // [N][n_params]
  int64_t n;
  const double* lines;      // [N][n_ulines][BAYSAR_LINE_STRIDE]  (K1 output)
  double* components;       // [N][n_chords + n_priors]; this kernel fills column `chord`
  double* spectra;          // optional [N][P]
  float* fine_pre;          // optional [N][F]
  float* fine_post;         // optional [N][F]
  uint32_t* status;         // [N] OR-ed
  // MODE 1 (spectral stage of the tensor-core path): zero-padded pre-convolution spectra and per-theta scalars
  float* spec;              // [rows][split_pitch], written at column split_pad + j
  int64_t split_pitch;
  int split_pad;
  double* aux;              // [N][BAYSAR_AUX_STRIDE]: pre-area, minimum, 4 float64 end outputs
  int skip_edge64;          // MODE 1: the consumer (folded pixel stage, no extrapolating pixel) does not read the float64 end outputs
  int keep_padding;         // MODE 1: the zero padding of the spectra rows is already in place (same layout as the last writer)
  // The chord's constant records, copied from the descriptor blob by the host (chord_args_set_chord): as kernel
  // parameters they are constant-bank operands, where every CTA used to start with a chain of dependent global loads
  // (model scalars -> chord record -> line records) executed by all of its threads.
  int chord;
  int n_params, n_ul, ncomp_cols;
  int kd_max;               // longest Doppler tap grid of the chord's Balmer lines
  int n_gauss_comps;        // Gaussian components of the chord's ADAS406 / X lines
  int ch_i[CH_I_COUNT];
  double ch_d[CH_D_COUNT];
};

// Host side: fill the constant part of ChordArgs for chord c from (host views of) the descriptor sections.
static inline void chord_args_set_chord(ChordArgs& a, const int64_t* I, const int* ch_i_all, const double* ch_d_all,
                                        const int* cl_i_all, int c) {
  a.chord = c;
  a.n_params = (int)I[I_N_PARAMS];
  a.n_ul = (int)I[I_N_ULINES];
  a.ncomp_cols = (int)I[I_N_CHORDS] + (int)I[I_N_PRIORS];
  for (int k = 0; k < CH_I_COUNT; ++k) a.ch_i[k] = ch_i_all[c * CH_I_COUNT + k];
  for (int k = 0; k < CH_D_COUNT; ++k) a.ch_d[k] = ch_d_all[c * CH_D_COUNT + k];
  a.kd_max = 0;
  a.n_gauss_comps = 0;
  for (int k = 0; k < a.ch_i[CH_N_LINES]; ++k) {
    const int* cl = cl_i_all + (a.ch_i[CH_LINE_OFF] + k) * CL_I_COUNT;
    if (cl[CL_KD] > a.kd_max) a.kd_max = cl[CL_KD];
    if (cl[CL_KIND] != 0) a.n_gauss_comps += cl[CL_NCOMP];
  }
}

#define BAYSAR_AUX_STRIDE 8
#define BAYSAR_BG_SLOT 60  // scratch slot of 10^theta_background (block_sum uses the first NV * NT / 32 <= 24 slots)

// Per-line constants of the fused narrow-Doppler Balmer passes (one 64-byte record per line, shared memory)
struct FusedLine {
  float jc_f, r, g_r, g_e;   // grid point nearest the centre, lambda[jc] - centre, gamma_rec^2.5, gamma_exc^2.5
  float s_r, s_e, c0, c1;    // amplitudes over the profile integrals; far-field series coefficients
  float c2, nfar;            // far field: groups whose points are all >= nfar grid points from the centre, where
                             // |d|^2.5 >= BAYSAR_FAR_RATIO * max(g_r, g_e) (and a three-tap Doppler kernel is a delta)
  int jL, jR;                // densely integrated window of sweep 1
  double sum_r, sum_e;       // line-of-sight integrals (recombination, excitation)
};
// Parameters of one Gaussian component, computed once per (theta, chord) (32 bytes)
struct alignas(16) GaussComp {
  double cw, inv_sigma;  // (Doppler-shifted) centre, 1 / sigma
  float amp;             // area-normalised amplitude over the intensity calibration
  int j_lo, j_hi, pad;   // fine-grid window |z| <= BAYSAR_Z_CUT
};
#define BAYSAR_WIN_GAMMAS 10.0   // sweep 1 sums +-max(BAYSAR_WIN_MIN, BAYSAR_WIN_GAMMAS gamma / delta) grid points around the
#define BAYSAR_WIN_MIN 48        // line centre point by point; the two tails beyond are integrated in closed form
#define BAYSAR_TAIL_TERMS 4      // terms of the tail series in g / |x|^2.5 (<= 1e-2.5 beyond 10 gamma: truncation < 1e-10)
#define BAYSAR_FAR_RATIO 500.f  // sweep 2: series in g / |d|^2.5 beyond this ratio (third-order term < 1e-8)
#define BAYSAR_FAR_SIGMAS 40.0  // multi-tap lines: the convolution is a moment series beyond this many Doppler widths
#ifndef BAYSAR_FUSED_GPT
#define BAYSAR_FUSED_GPT 4      // sweep 2: groups of four consecutive points a thread keeps in registers per chunk
#endif

__host__ __device__ static inline int pad_index(int i) { return i ^ ((i >> 3) & 4); }

struct ChordSmem {
  size_t buf_floats;   // physical floats per padded fine-grid buffer
  size_t taps_floats;  // floats in the q-indexed tap buffer
  size_t tapd_doubles; // doubles for Doppler taps + prefix sums
  size_t total;        // bytes
};

__host__ __device__ static inline ChordSmem chord_smem_layout(int F, int halo, int taps_max, int kd_max, int need_b = 1,
                                                              int n_lines = 0) {
  ChordSmem s;
  int f8 = (F + 7) / 8 * 8;
  int logical = f8 + 2 * halo + 16;
  s.buf_floats = (size_t)(logical + 8 + 3) / 4 * 4;
  s.taps_floats = (size_t)(taps_max + 8 + 3) / 4 * 4;
  s.tapd_doubles = 2 * (size_t)kd_max + 4;
  s.total = (need_b ? 2 : 1) * s.buf_floats * sizeof(float) + s.taps_floats * sizeof(float) + s.tapd_doubles * sizeof(double) +
            192 * sizeof(double) + 16 * sizeof(int) + 8 * BAYSAR_MAX_BALMER_LINES * sizeof(double) +
            4 * BAYSAR_MAX_BALMER_LINES * sizeof(float) + (size_t)(n_lines + 15) / 16 * 16;  // + one class byte per line
  s.total = (s.total + 15) / 16 * 16;
  return s;
}

// fine-grid coordinate, bit-identical to numpy.arange's fill: start + j * delta (two roundings, no FMA)
BAYSAR_DEV double grid_x(double x0, double delta, int j) { return __dadd_rn(x0, __dmul_rn((double)j, delta)); }

BAYSAR_DEV float4 ld4(const float* buf, int e) { return *reinterpret_cast<const float4*>(buf + pad_index(e)); }

// 8 consecutive outputs of a zero-padded "same" convolution.  `e0` = logical element index of (first output + q_start)
// in the padded buffer (a multiple of 4); tq[qi] = w[o - (q_start + qi)], nq a multiple of 4.
BAYSAR_DEV void conv8(const float* buf, int e0, const float* tq, int nq, float* acc) {
#pragma unroll
  for (int r = 0; r < 8; ++r) acc[r] = 0.f;
  float4 c0 = ld4(buf, e0), c1 = ld4(buf, e0 + 4);
  for (int qi = 0; qi < nq; qi += 4) {
    const float4 c2 = ld4(buf, e0 + qi + 8);
    const float4 w = *reinterpret_cast<const float4*>(tq + qi);
    const float s[12] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w, c2.x, c2.y, c2.z, c2.w};
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      acc[r] = fmaf(w.x, s[r], acc[r]);
      acc[r] = fmaf(w.y, s[r + 1], acc[r]);
      acc[r] = fmaf(w.z, s[r + 2], acc[r]);
      acc[r] = fmaf(w.w, s[r + 3], acc[r]);
    }
    c0 = c1;
    c1 = c2;
  }
}

// v ~ hi + lo with both terms representable in TF32 (10-bit mantissa): the operands of the 3xTF32 contraction
#ifndef BAYSAR_CPU_EMU
BAYSAR_DEV void split_tf32_pair(float v, float& hi, float& lo) {
  uint32_t h, l;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(v));
  hi = __uint_as_float(h);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(v - hi));
  lo = __uint_as_float(l);
}
#else
BAYSAR_DEV float emu_tf32(float v) {
  uint32_t u; std::memcpy(&u, &v, 4);
  u = (u + 0x1000u) & 0xFFFFE000u;
  float r; std::memcpy(&r, &u, 4);
  return r;
}
BAYSAR_DEV void split_tf32_pair(float v, float& hi, float& lo) { hi = emu_tf32(v); lo = emu_tf32(v - hi); }
#endif

// exp(arg) for arg <= 0 computed as 2^n * ex2(f): the argument reduction is done in float64 so that the result
// keeps float32 relative accuracy even for arg ~ -40 (a narrow Gaussian sampled far from its centre).
BAYSAR_DEV float exp_neg_f32(double arg) {
  double t = arg * 1.4426950408889634;
  double nn = rint(t);
  float f = (float)(t - nn);
  int e = (int)nn;
  if (e < -126) return 0.f;
  return exp2f(f) * __int_as_float((e + 127) << 23);
}

// Fetchers for the convolved fine-grid spectrum: shared memory (fused kernel) or global memory (tensor-core path)
struct FetchSmem {
  const float* buf; int halo;
  BAYSAR_DEV double operator()(int i) const { return (double)buf[pad_index(halo + i)]; }
};
struct FetchGlobal {
  const float* row;
  BAYSAR_DEV double operator()(int i) const { return (double)row[i]; }
};
// compact rows of the tensor-core path: only the columns a static wavelength map reads, packed in column order
struct FetchCompact {
  const float* row;
  const uint32_t* cmask;  // bit e of word w: column 32 w + e is stored
  const int* cbase;       // stored columns before word w
  BAYSAR_DEV double operator()(int i) const {
    const int w = i >> 5;
    return (double)row[cbase[w] + __popc(cmask[w] & ((1u << (i & 31)) - 1u))];
  }
};

// Folded rows of pix_contract.cuh: one column per pixel (instrument convolution and interp1d in one contraction) plus
// the explicit fine-grid columns of the two edge zones, where the clip at the pre-convolution minimum can bite.
struct FetchFolded {
  const float* row;        // [out_pitch] contraction output of this theta
  const int* pix_col;      // [P] column of pixel p, or -1: pixel reads the explicit edge columns
  int col_left0, x_l;      // explicit columns j = 0 .. x_l - 1 at row[col_left0 + j]
  int col_right0, x_r, F;  // explicit columns j = F - x_r .. F - 1 at row[col_right0 + j - (F - x_r)]
  float mn;                // clip level
  const float* kappa;      // [column tiles] bias correction of the folded columns (shared memory)
  int tile_shift;          // log2 of the column tile width
  int use_edge64;          // 0: the float64 end outputs were not formed (no pixel extrapolates): edge pixels read the columns
  BAYSAR_DEV double operator()(int i) const {
    const float v = i < x_l ? row[col_left0 + i] : row[col_right0 + (i - (F - x_r))];
    return (double)(v < mn ? mn : v);
  }
};

// Convolved spectrum at pixel p: interp1d between fine-grid points i and i + 1 (the four end outputs in float64)
template <class Fetch>
BAYSAR_DEV double pixel_value(const Fetch& fetch, int p, int i, double w, const double* edge64, int F) {
  const double v0 = (i < 2) ? edge64[i] : (i >= F - 2 ? edge64[i - (F - 4)] : fetch(i));
  const int i1 = i + 1;
  const double v1 = (i1 < 2) ? edge64[i1] : (i1 >= F - 2 ? edge64[i1 - (F - 4)] : fetch(i1));
  return v0 + w * (v1 - v0);
}
BAYSAR_DEV double pixel_value(const FetchFolded& fetch, int p, int i, double w, const double* edge64, int F) {
  const int col = fetch.pix_col[p];
  if (col >= 0) return (double)fetch.kappa[col >> fetch.tile_shift] * (double)fetch.row[col];
  if (!fetch.use_edge64) {
    const double v0 = fetch(i);
    return v0 + w * (fetch(i + 1) - v0);
  }
  return pixel_value<FetchFolded>(fetch, p, i, w, edge64, F);
}

// Area renormalisation, + background, wavelength map, clip, Cauchy likelihood (spectrometers.py:286-310, :154-195).
// `post` and `pre` are the trapz integrals after / before the instrument convolution; edge64 the four end outputs;
// bg = 10^theta_background (computed once per CTA by the caller: a float64 pow per thread was a fifth of the pixel
// kernels' instructions).
template <int NT, class Fetch>
BAYSAR_DEV void pixel_likelihood(const ModelDev& m, const int* ch, const double* chd, const double* th, int64_t n, int c,
                                 int ncomp_cols, double pre, double post, double bg, const double* edge64, Fetch fetch,
                                 double* scratch, double* components, double* spectra, float* fine_post, uint32_t& st) {
  const int tid = threadIdx.x;
  const int F = ch[CH_F], P = ch[CH_P];
  const double x0 = chd[CH_X0], delta = chd[CH_DELTA];
  const double ratio = post / pre;
  const double inv_ratio = 1.0 / ratio;
  const double ratio2 = (post / ratio) / pre;
  if (!(fabs(ratio2 - 1.0) <= 1e-8 + 1e-2)) st |= 1u << 9;  // np.isclose(ratio2, 1, rtol=1e-2)
  if (fine_post) {
    float* o_ = fine_post + n * (int64_t)F;
    for (int j = tid; j < F; j += NT) o_[j] = (float)(fetch(j) / ratio + bg);
  }
  const int poff = ch[CH_PIX_OFF];
  const int cw_idx = ch[CH_CALWAVE_IDX];
  const double cwl_map = cw_idx >= 0 ? th[cw_idx] : chd[CH_CALWAVE_CWL];
  const double disp_map = cw_idx >= 0 ? th[cw_idx + 1] : chd[CH_CALWAVE_DISP];
  double acc = 0.0;
  const bool gauss = m.I[I_GAUSS_LIKELIHOOD] != 0;
  for (int p = tid; p < P; p += NT) {
    int i; double w;
    if (cw_idx >= 0) {
      double px = ((double)p - chd[CH_PIX_MEAN]) * disp_map + cwl_map;
      double g = floor((px - x0) / delta);
      int jj = g < 0.0 ? 0 : (g > (double)(F - 1) ? F - 1 : (int)g);
      while (jj > 0 && grid_x(x0, delta, jj - 1) >= px) --jj;
      while (jj < F && grid_x(x0, delta, jj) < px) ++jj;
      i = jj - 1;
      i = i < 0 ? 0 : (i > F - 2 ? F - 2 : i);
      double xi = grid_x(x0, delta, i);
      w = (px - xi) / (grid_x(x0, delta, i + 1) - xi);
    } else {
      i = m.pix_idx[poff + p];
      w = m.pix_frac[poff + p];
    }
    double fm = pixel_value(fetch, p, i, w, edge64, F) * inv_ratio + bg;
    fm = fm < bg ? bg : fm;
    if (!(fm - fm == 0.0)) st |= 1u << 10;
    if (spectra) spectra[n * (int64_t)P + p] = fm;
    // Cauchy term log(1 + r^2), r = (y - fm) / err: the difference in float64 (it cancels), quotient and logarithm in
    // float32 (1 ulp log1pf: <= 1.2e-7 of each term, far inside the 1e-6 budget of the sum), the sum in float64.
    // Residuals beyond float32's comfortable range (|r| >= 1e15) take the float64 route.
    const double d = m.y[poff + p] - fm, er = m.err[poff + p];
    const float rf = (float)d / (float)er;
    const float ar = fabsf(rf);
    if (gauss) {  // likelihood(cauchy=False), spectrometers.py:187-188: -0.5 sum r^2 (float64: it is not the default path)
      const double r = d / er, r2 = r * r;
      if (r2 == 0.0) st |= 1u << 11;
      if (!(r2 - r2 == 0.0)) st |= 1u << 12;
      acc += 0.5 * r2;
    } else if (ar < 1e15f) {
      if (d == 0.0 || ar == 0.f) st |= 1u << 11;
      acc += (double)log1pf(ar * ar);
    } else {
      const double r = d / er, r2 = r * r;
      if (r2 == 0.0) st |= 1u << 11;
      if (!(r2 - r2 == 0.0)) st |= 1u << 12;
      acc += log(1.0 + r2);
    }
  }
  double red4[1] = {acc};
  block_sum<NT, 1>(red4, scratch);
  if (tid == 0) components[n * (int64_t)ncomp_cols + c] = -red4[0];
}

// trapz (unit spacing in the grid index, half weights at the two ends) of 1 / (|x_j|^2.5 + g), x_j = (j - jc) delta + r,
// over the grid indices [ja, jb] of a tail: both ends on one side of the line centre and beyond ~10 Stark half-widths, where
// g / |x|^2.5 <= 10^-2.5.  Integral of the four-term series in that ratio plus the first two Euler-Maclaurin end
// corrections (tests/kernel_model.py: < 2.1e-11 of the whole integral in float64; float32 here, the tails are a few per
// cent of it).  f_ends (may be null): the integrand at ja and jb.
BAYSAR_DEV float stark_tail_trapz(int ja, int jb, float jc_f, float r, float g, float d_hi, float d_lo, float df, float* f_ends) {
  float A[2], f1[2], f3[2];
  float sgn = 1.f;
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const float kf = (float)(e == 0 ? ja : jb) - jc_f;
    const float xs = fmaf(kf, d_hi, fmaf(kf, d_lo, r));
    if (e == 0) sgn = xs > 0.f ? 1.f : -1.f;
    const float x = fabsf(xs), sq = sqrt_approx(x), x25 = x * x * sq;
    const float ix15 = rcp_approx(x * sq), u = -g * rcp_approx(x25);  // x^-1.5, -g x^-2.5
    A[e] = -ix15 * (1.f / 1.5f + u * (1.f / 4.f + u * (1.f / 6.5f + u * (1.f / 9.f))));
    const float D1 = 2.5f * x * sq, D2 = 3.75f * sq, D3 = 1.875f * rcp_approx(sq);
    const float iD = rcp_approx(x25 + g), iD2 = iD * iD;
    f1[e] = -D1 * iD2;
    f3[e] = -D3 * iD2 + 6.f * D1 * D2 * iD2 * iD - 6.f * D1 * D1 * D1 * iD2 * iD2;
    if (f_ends) f_ends[e] = iD;
  }
  return sgn * ((A[1] - A[0]) * (1.f / df) + df * (f1[1] - f1[0]) * (1.f / 12.f) - df * df * df * (f3[1] - f3[0]) * (1.f / 720.f));
}

// Sweep 2 of the fused narrow-Doppler passes.  Far field, two fine-grid points (packed FP32 pipe, fma.rn.f32x2):
// v2 += s_r / (a + g_r) + s_e / (a + g_e) as the three-term series ia (c0 - ia (c1 - ia c2)), ia = 1 / a = rsqrt(|d|)^5,
// a = |d|^2.5 at the grid offsets kf2 from the line centre (one MUFU op per point instead of two).
BAYSAR_DEV void stark_far2(f32x2 kf2, f32x2& v2, f32x2 dhi2, f32x2 dlo2, f32x2 r2, f32x2 c0p, f32x2 mc1p, f32x2 c2p) {
  const f32x2 d2 = fma2(kf2, dhi2, fma2(kf2, dlo2, r2));
  float dx, dy;
  unpack2(d2, dx, dy);
  const f32x2 rs = pack2(rsqrt_approx(fabsf(dx)), rsqrt_approx(fabsf(dy)));
  const f32x2 rs2 = mul2(rs, rs);
  const f32x2 ia = mul2(mul2(rs2, rs2), rs);
  v2 = fma2(ia, fma2(ia, fma2(ia, c2p, mc1p), c0p), v2);
}
// The exact form next to the line centre, one point: kf grid points from the centre.
BAYSAR_DEV float stark_exact(float kf, float d_hi, float d_lo, float r, float g_r, float g_e, float s_r, float s_e) {
  const float d = fabsf(fmaf(kf, d_hi, fmaf(kf, d_lo, r)));
  const float a25 = d * d * sqrt_approx(d);
  const float dr = a25 + g_r, de = a25 + g_e;
  return fmaf(s_r, de, s_e * dr) * rcp_approx(dr * de);
}

// MINB = resident CTAs per SM the register allocation aims at: 3 (80 registers) when the shared-memory footprint allows
// no more (long fine grids), 4 (64 registers, no spills) or 5 (48 registers, 16 bytes of spills) for short grids.
// MODE 0: fused (spectrum -> instrument convolution on the FP32 pipe -> likelihood, everything in shared memory).
// MODE 1: spectral stage only: writes the pre-convolution spectrum and the per-theta scalars for the tensor-core
//         contraction (if_contract.cuh) and the pixel stage (pixel_stage_kernel below).
// TAPS: what the host guarantees about the Doppler kernels of the chord's Balmer lines for every theta inside the bounds
// (model.py): 0 = single taps only (CH_B_OPTIONAL) -- the three-tap variant of the fused passes and the multi-tap path
// are compiled out; 1 = at most three taps (CH_NARROW, cold neutrals on a fine grid) -- the multi-tap path is compiled
// out; 2 = anything.  The leaner builds keep the register budgets of the fused sweeps free of spills.
template <int NT, int MODE, int MINB = 3, int TAPS = 2>
__global__ void __launch_bounds__(NT, MINB) chord_stage_kernel(ModelDev m, ChordArgs a) {
  constexpr bool FUSED3 = TAPS >= 1;
#ifdef BAYSAR_CPU_EMU
  unsigned char* chord_smem_raw = emu_dynamic_smem();
#else
  extern __shared__ __align__(16) unsigned char chord_smem_raw[];
#endif
  pdl_trigger();
  pdl_wait();  // the line records come from the LOS stage
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t n = blockIdx.x;
  const int c = a.chord;
  const int n_params = a.n_params, n_ul = a.n_ul;
  const int ncomp_cols = a.ncomp_cols;
  const int* ch = a.ch_i;       // kernel parameters: constant-bank operands
  const double* chd = a.ch_d;
  const int F = ch[CH_F], HALO = ch[CH_HALO], TAPS_MAX = ch[CH_TAPS_MAX];
  const double x0 = chd[CH_X0], delta = chd[CH_DELTA], dl = chd[CH_DL];
  const double* th = a.theta + n * n_params;
  const double* lines_n = a.lines + n * (int64_t)n_ul * BAYSAR_LINE_STRIDE;

  // ---- shared memory carve-up (must match chord_smem_layout)
  const int kd_max = a.kd_max;
  // buffer B (scratch / convolved spectrum) is left out when the tensor-core path's spectral stage cannot need it
  const bool need_b = MODE == 0 || ch[CH_B_OPTIONAL] == 0;
  const ChordSmem lay = chord_smem_layout(F, HALO, TAPS_MAX, kd_max, need_b ? 1 : 0, ch[CH_N_LINES]);
  const int f8 = (F + 7) / 8 * 8;
  float* bufA = reinterpret_cast<float*>(chord_smem_raw);
  float* bufB = bufA + lay.buf_floats;
  float* tapq = bufA + (need_b ? 2 : 1) * lay.buf_floats;
  double* tapd = reinterpret_cast<double*>(tapq + lay.taps_floats);  // [kd_max] taps, then [kd_max + 1] prefix sums
  double* tappre = tapd + kd_max;
  double* scratch = tappre + kd_max + 4;  // 192 doubles: [0,128) block_sum (up to 16 values x 8 warps), [128,132) edge64
  int* iscr = reinterpret_cast<int*>(scratch + 192);  // 16 ints
  // constants of the Balmer lines handled by the fused narrow-Doppler passes (BAYSAR_MAX_BALMER_LINES x 64 bytes, slot =
  // rank of the line among the chord's Balmer lines); a pair of sweeps covers the slots [fb, fb + nf)
  FusedLine* fl_all = reinterpret_cast<FusedLine*>(iscr + 16);
  // Doppler taps of the fused lines, normalised to sum 1: weights of stark[j - 1], stark[j], stark[j + 1]; [3] != 0 marks
  // a three-tap line (the Doppler kernel of a cold neutral on a fine grid: central tap + two ~1e-7 neighbours)
  float* fw_all = reinterpret_cast<float*>(fl_all + BAYSAR_MAX_BALMER_LINES);
  // class of every line object, decided once by one thread per line (parallel across lines) before the serial loop over
  // the lines: 0 not a Balmer line, 1 single Doppler tap, 3 three taps, 2 multi-tap
  unsigned char* line_class = reinterpret_cast<unsigned char*>(fw_all + 4 * BAYSAR_MAX_BALMER_LINES);
  double* edge64 = scratch + 128;
#define A_AT(j) bufA[pad_index(HALO + (j))]
#define B_AT(j) bufB[pad_index(HALO + (j))]

  uint32_t st = 0;
  if (need_b) {
    for (size_t i = tid; i < lay.buf_floats; i += NT) bufA[i] = 0.f;
  } else {
    // only Balmer lines in the fused passes, whose last pass writes every group of four points: zero the haloes only
    const int interior_end = HALO + 4 * ((F + 3) >> 2);
    for (int i = tid; i < HALO; i += NT) bufA[pad_index(i)] = 0.f;
    for (int i = interior_end + tid; i < (int)lay.buf_floats; i += NT) bufA[pad_index(i)] = 0.f;
  }
  if (need_b) {
    for (int i = tid; i < HALO; i += NT) bufB[pad_index(i)] = 0.f;                       // left halo of B
    for (int i = HALO + F + tid; i < (int)lay.buf_floats; i += NT) bufB[pad_index(i)] = 0.f;  // tail + right halo of B
  }
  // ---- class of every Balmer line (block-uniform control flow below reads it back): the tap at the "same"-alignment
  // centre is inside the Doppler cut and either its neighbours are not (single tap), or they are and the next ones are
  // not (three taps); anything else takes the multi-tap path
  for (int k = tid; k < ch[CH_N_LINES]; k += NT) {
    const int* cl = m.cl_i + (ch[CH_LINE_OFF] + k) * CL_I_COUNT;
    unsigned char cls = 0;
    if (cl[CL_KIND] == 0) {
      const double* cld = m.cl_d + (ch[CH_LINE_OFF] + k) * CL_D_COUNT;
      const double* rec = lines_n + cl[CL_ULINE] * BAYSAR_LINE_STRIDE;
      const double cwl = rec[13], sigma = rec[14];
      const int KD = cl[CL_KD], o = (KD - 1) / 2;
      const double dshift = cwl - cld[CL_CWL0];
      double zc[5];
      const double zlim = BAYSAR_Z_CUT_DOPPLER * sigma;
#pragma unroll
      for (int t = 0; t < 5; ++t) {
        const double xk = __dadd_rn(cld[CL_DOP_START], __dmul_rn((double)(o - 2 + t), cld[CL_DOP_DELTA]));
        zc[t] = fabs((xk + dshift) - cwl);
      }
      const bool one_tap = KD >= 3 && zc[2] <= zlim && zc[1] > zlim && zc[3] > zlim;
      const bool three_taps = FUSED3 && KD >= 5 && zc[2] <= zlim && zc[1] <= zlim && zc[3] <= zlim && zc[0] > zlim && zc[4] > zlim;
      cls = one_tap ? 1 : (three_taps ? 3 : 2);
      if (!(sigma > 0.0)) st |= 1u << 5;
      if (cls != 2) {
        // shape record of a fused line: lambda_j - centre = (j - jc) delta + r with jc the grid point nearest the centre
        const double c0 = x0 - cwl;
        const int jc = (int)rint(-c0 / delta);
        FusedLine& L = fl_all[cl[CL_BRANK]];
        L.jc_f = (float)jc; L.r = (float)fma((double)jc, delta, c0); L.g_r = (float)rec[11]; L.g_e = (float)rec[12];
        L.sum_r = rec[0]; L.sum_e = rec[1];
        // densely integrated window of sweep 1: +-10 Stark half-widths (gamma = g^0.4), at least BAYSAR_WIN_MIN points
        // (float32 with fast powers: any width at or above ~10 gamma holds the tail series' accuracy)
        const float gmax = L.g_r > L.g_e ? L.g_r : L.g_e;
        float nw = ceilf((float)BAYSAR_WIN_GAMMAS * 1.001f * __powf(gmax, 0.4f) / (float)delta);
        nw = nw > (float)BAYSAR_WIN_MIN ? nw : (float)BAYSAR_WIN_MIN;  // NaN -> minimum
        const int n_w = nw < (float)F ? (int)nw : F;
        int jL = jc - n_w, jR = jc + n_w;
        jL = jL < 0 ? 0 : (jL > F - 1 ? F - 1 : jL);
        jR = jR < 0 ? 0 : (jR > F - 1 ? F - 1 : jR);
        L.jL = jL; L.jR = jR;
        float* w = fw_all + 4 * cl[CL_BRANK];
        if (three_taps) {
          // out[j] = g[o - 1] stark[j + 1] + g[o] stark[j] + g[o + 1] stark[j - 1], normalised by the tap sum (the
          // trapz of the convolved profile is the tap sum times the trapz of the Stark profile up to ~1e-12: the
          // end-point terms are g[o +- 1] / g[o] ~ 1e-7 times the far wing)
          const double gm = exp(-0.5 * (zc[1] / sigma) * (zc[1] / sigma)), g0 = exp(-0.5 * (zc[2] / sigma) * (zc[2] / sigma)),
                       gp = exp(-0.5 * (zc[3] / sigma) * (zc[3] / sigma));
          const double gs = gm + g0 + gp;
          w[0] = (float)(gp / gs); w[1] = (float)(g0 / gs); w[2] = (float)(gm / gs); w[3] = 1.f;
        } else {
          w[0] = 0.f; w[1] = 1.f; w[2] = 0.f; w[3] = 0.f;
        }
      }
    }
    line_class[k] = cls;
  }
  __syncthreads();

  const double cal = ch[CH_CAL_IDX] >= 0 ? pow(10.0, th[ch[CH_CAL_IDX]]) : 1.0;
  const double fwhm_to_sigma = 0.42466090014400953;  // 1 / sqrt(8 ln 2)
  const double c_light = 299792458.0;
  const float NTf = (float)NT;

  const float d_hi = (float)delta, d_lo = (float)(delta - (double)(float)delta);

  // ================= Balmer lines whose Doppler kernel is a single tap (sigma << grid step) ====================
  // Stark (x) delta = Stark: no convolution, no stash; ALL such lines share two sweeps over the grid and one block
  // reduction: sweep 1 integrates u_rec / u_exc of every line (trapz, half weights at the ends), sweep 2 adds
  // sum_b (rec_sum_b u_rec_b / I_rec_b + exc_sum_b u_exc_b / I_exc_b) to the spectrum.
  int nf = 0, fb = 0;  // lines of the current pair of sweeps: shape-record slots [fb, fb + nf)
  // accumulated by the LAST fused pass (which then also owns the pre-convolution area / minimum and, in MODE 1, the
  // hand-over of the spectrum) or by the separate pass further down
  float pre_f = 0.f;
  float mn = 3.402823466e38f;
  int bad = 0;
  bool pre_done = false;
  // buffer A holds something other than zeros (Gaussian components, an earlier group of lines): block-uniform
  bool a_dirty = false;
  auto flush_fused = [&](const bool final_pass) {
    FusedLine* fl = fl_all + fb;
    float* fw = fw_all + 4 * fb;
    // ---- sweep 1: I_rec, I_exc = trapz of the two Stark profiles of every line over the whole grid.
    // Only a window of +-10 Stark half-widths around the line centre (>= 48 points) is summed point by point.  On the two
    // tails g / |x|^2.5 <= 10^-2.5 and the integrand is smooth on the scale of the grid, so the trapezoid is the integral
    // plus the first two Euler-Maclaurin end corrections,
    //   T_1[a, b] = (1 / delta) (A(x_b) - A(x_a)) + (f1(b) - f1(a)) / 12 - (f3(b) - f3(a)) / 720,
    //   A(x) = - sum_k (-g)^k x^-(2.5 k + 1.5) / (2.5 k + 1.5)   (antiderivative of 1 / (x^2.5 + g), four terms),
    // f1, f3 the first and third derivative per grid index: a closed form formed by one thread per (line, profile).
    // Relative error of the integral < 1e-10 for every width and centre (tests/test_kernel_model.py).
    // The lines are dealt to the warps (a group of wpl warps per line): every warp pays the per-line set-up once, for
    // its own line, instead of all warps for all lines.
    {
      constexpr int NW = NT / 32;
      int wpl = 1;
      while (2 * wpl * nf <= NW) wpl *= 2;
      for (int b = warp / wpl; b < nf; b += NW / wpl) {
        const int sw = warp % wpl;
        const float4 cst = *reinterpret_cast<const float4*>(&fl[b].jc_f);  // jc, r, gamma_rec^2.5, gamma_exc^2.5
        const int jL = fl[b].jL, n_c = fl[b].jR - jL + 1;
        const float w_end = n_c == 1 ? 0.f : 0.5f;
        float sr = 0.f, se = 0.f;
        for (int t = sw * 32 + lane; t < n_c; t += 32 * wpl) {
          const float kf = (float)(jL + t) - cst.x;
          const float d = fabsf(fmaf(kf, d_hi, fmaf(kf, d_lo, cst.y)));
          const float a25 = d * d * sqrt_approx(d);
          const float dr = a25 + cst.z, de = a25 + cst.w;
          const float inv = ((t == 0 || t == n_c - 1) ? w_end : 1.f) * rcp_approx(dr * de);
          sr = fmaf(de, inv, sr);
          se = fmaf(dr, inv, se);
        }
        const double wr = warp_sum((double)sr), we = warp_sum((double)se);
        if (lane == 0) { scratch[(2 * b) * NW + sw] = wr; scratch[(2 * b + 1) * NW + sw] = we; }
      }
      __syncthreads();
      if (warp == 0) {
        // one thread per (line, recombination / excitation): window sum, closed-form tails, amplitude
        if (tid < 2 * nf) {
          FusedLine& L = fl[tid >> 1];
          const int q = tid & 1;
          double integ = 0.0;
          for (int w = 0; w < wpl; ++w) integ += scratch[tid * NW + w];
          // the two tails are at most a few per cent of the integral: float32 with approximate reciprocals is ample
          const float g = q == 0 ? L.g_r : L.g_e;
          float tails = 0.f;
#pragma unroll
          for (int side = 0; side < 2; ++side) {
            const int ja = side == 0 ? 0 : L.jR, jb = side == 0 ? L.jL : F - 1;
            if (jb > ja) tails += stark_tail_trapz(ja, jb, L.jc_f, L.r, g, d_hi, d_lo, (float)delta, nullptr);
          }
          integ += (double)tails;
          const double amp = (q == 0 ? L.sum_r : L.sum_e) / (dl * integ) / cal;
          if (!(amp - amp == 0.0)) atomicOr(&a.status[n], 1u << 6);
          (q == 0 ? L.s_r : L.s_e) = (float)amp;
        }
      __syncwarp();  // both amplitudes of a line come from neighbouring lanes of this warp
      if (tid < nf) {
        // sweep-2 constants: exact form next to the centre, far-field series
        //   s_r / (a + g_r) + s_e / (a + g_e) = (1/a) (c0 - (1/a) (c1 - (1/a) c2)) + O((g/a)^3),  a = |d|^2.5
        FusedLine& L = fl[tid];
        L.c0 = L.s_r + L.s_e;
        L.c1 = L.s_r * L.g_r + L.s_e * L.g_e;
        L.c2 = L.s_r * L.g_r * L.g_r + L.s_e * L.g_e * L.g_e;
        const float gmax = L.g_r > L.g_e ? L.g_r : L.g_e;
        // |d| >= d_far for every point of a group of four whose nearest point is nfar grid points from the centre
        // (|r| <= delta / 2)
        const float d_far = exp2f(0.4f * log2f(BAYSAR_FAR_RATIO * gmax)) * 1.0001f;
        float nfar = d_far / (float)delta + 0.5f;
        // Three Doppler taps (w-, w0, w+ ~ 1e-6, normalised): far from the centre the Stark profile S is smooth on the scale
        // of the grid and  w- S(j-1) + w0 S(j) + w+ S(j+1) = S(j) + (w+ - w-) S' + (w+ + w-) S''/2 + ...  with
        // |S'/S| <= 2.5 / k and |S''/S| <= 8.75 / k^2 at k grid points from the centre: beyond the k where both terms drop
        // below 1e-9 the kernel is a delta (one evaluation per point instead of three)
        const float* w = fw + 4 * tid;
        if (FUSED3 && w[3] != 0.f) {
          const float wsum = w[0] + w[2], wdif = fabsf(w[2] - w[0]);
          const float n2 = sqrt_approx(4.375f * wsum * 1e9f), n1 = 2.5f * wdif * 1e9f;
          const float n3 = (n2 > n1 ? n2 : n1) + 2.f;
          nfar = nfar > n3 ? nfar : n3;
        }
        L.nfar = nfar;
      }
      }
    }
    __syncthreads();
    // ---- sweep 2: spectrum += sum_b (s_rec_b / den_rec + s_exc_b / den_exc).  Every thread owns GPT groups of four
    // CONSECUTIVE fine-grid points per chunk (group g = chunk base + tid + i NT) and keeps their sums in registers while it
    // loops over the lines, so that the line constants are read once per line and thread, the result leaves as one
    // 128-bit shared-memory store (and one 128-bit global store in MODE 1) per group, and the far / near decision is one
    // test per group on the integer distance to the line centre.  Far-field points go two at a time through the packed
    // FP32 pipe (fma.rn.f32x2).
    constexpr int GPT = NT == 128 ? 2 * BAYSAR_FUSED_GPT : (MINB >= 6 ? BAYSAR_FUSED_GPT / 2 : BAYSAR_FUSED_GPT);
    float* sp = (MODE == 1) ? a.spec + n * a.split_pitch + a.split_pad : nullptr;
    const f32x2 dhi2 = pack2(d_hi, d_hi), dlo2 = pack2(d_lo, d_lo);
    const int n_groups = (F + 3) >> 2;
    for (int gbase = 0; gbase < n_groups; gbase += GPT * NT) {
      f32x2 v01[GPT], v23[GPT];
#pragma unroll
      for (int i = 0; i < GPT; ++i) { v01[i] = pack2(0.f, 0.f); v23[i] = pack2(0.f, 0.f); }
      // group i of this thread starts at grid point 4 (gbase + tid + i NT): an arithmetic progression, so the offsets from
      // a line centre advance by one packed add per group.  Groups past the end of the grid are evaluated where they
      // would lie (far field, finite) and never stored: no bounds test inside the line loop.
      const float j00 = (float)(4 * (gbase + tid));
      const f32x2 jb2 = pack2(j00, j00 + 1.f), two2 = pack2(2.f, 2.f), step2 = pack2((float)(4 * NT), (float)(4 * NT));
      for (int b = 0; b < nf; ++b) {
        const float4 cst = *reinterpret_cast<const float4*>(&fl[b].jc_f);  // jc, r, gamma_rec^2.5, gamma_exc^2.5
        const float4 ser = *reinterpret_cast<const float4*>(&fl[b].s_r);   // s_r, s_e, c0, c1
        const float2 cf = *reinterpret_cast<const float2*>(&fl[b].c2);     // c2, nfar
        const f32x2 r2 = pack2(cst.y, cst.y);
        const f32x2 c0p = pack2(ser.z, ser.z), mc1p = pack2(-ser.w, -ser.w), c2p = pack2(cf.x, cf.x);
        const float4 w4 = *reinterpret_cast<const float4*>(fw + 4 * b);
        const bool three = FUSED3 && w4.w != 0.f;
        const float far_lim = cf.y + 1.5f;  // |offset of the group's middle| from which all four points are >= nfar away
        f32x2 kfA = add2(jb2, pack2(-cst.x, -cst.x));
#pragma unroll
        for (int i = 0; i < GPT; ++i) {
          float kf0, kf1;
          unpack2(kfA, kf0, kf1);
          bool far = fabsf(kf0 + 1.5f) >= far_lim;
          if (FUSED3 && three) {
            // a three-tap line keeps the exact form in the first and the last group: the taps past the two grid ends are
            // zero padding ("same" convolution)
            const int g = gbase + tid + i * NT;
            far = far && g != 0 && g != n_groups - 1;
          }
          if (far) {
            stark_far2(kfA, v01[i], dhi2, dlo2, r2, c0p, mc1p, c2p);
            stark_far2(add2(kfA, two2), v23[i], dhi2, dlo2, r2, c0p, mc1p, c2p);
          } else {
            float o[4];
            if (!three) {
#pragma unroll
              for (int k = 0; k < 4; ++k) o[k] = stark_exact(kf0 + (float)k, d_hi, d_lo, cst.y, cst.z, cst.w, ser.x, ser.y);
            } else {
              // out[j] = w.x S(j - 1) + w.y S(j) + w.z S(j + 1); S vanishes outside the grid
              const int g = gbase + tid + i * NT;
              float sv[6];
#pragma unroll
              for (int k = 0; k < 6; ++k) {
                const int j = 4 * g - 1 + k;
                const float val = stark_exact(kf0 + (float)(k - 1), d_hi, d_lo, cst.y, cst.z, cst.w, ser.x, ser.y);
                sv[k] = (j >= 0 && j < F) ? val : 0.f;
              }
#pragma unroll
              for (int k = 0; k < 4; ++k) o[k] = fmaf(w4.x, sv[k], fmaf(w4.y, sv[k + 1], w4.z * sv[k + 2]));
            }
            v01[i] = add2(v01[i], pack2(o[0], o[1]));
            v23[i] = add2(v23[i], pack2(o[2], o[3]));
          }
          kfA = add2(kfA, step2);
        }
      }
#pragma unroll
      for (int i = 0; i < GPT; ++i) {
        const int g = gbase + tid + i * NT;
        if (g >= n_groups) continue;
        const int j = 4 * g;
        float v[4];
        unpack2(v01[i], v[0], v[1]);
        unpack2(v23[i], v[2], v[3]);
        float4* ap = reinterpret_cast<float4*>(bufA + pad_index(HALO + j));  // HALO and j are multiples of 4
        if (a_dirty) {
          const float4 prev = *ap;
          v[0] += prev.x; v[1] += prev.y; v[2] += prev.z; v[3] += prev.w;
        }
        if (j + 3 >= F) {  // last group of a grid that is not a multiple of four: the alignment tail stays zero
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (j + k >= F) v[k] = 0.f;
        }
        // the spectrum stays in shared memory only if something reads it there: an earlier or later pass, the in-kernel
        // convolution, the fine-grid diagnostic; the folded path's single fused pass hands it over from registers
        if (!(MODE == 1 && final_pass && !a_dirty && a.skip_edge64 && a.fine_pre == nullptr)) *ap = make_float4(v[0], v[1], v[2], v[3]);
        if (final_pass) {  // pre-convolution area / minimum (spectrometers.py:256, :271) and MODE 1 hand-over
          const float s4 = (v[0] + v[1]) + (v[2] + v[3]);
          pre_f += s4;
          if (!(s4 - s4 == 0.f)) bad = 1;
          if (j == 0) pre_f -= 0.5f * v[0];
          if (j + 3 >= F - 1) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (j + k == F - 1) pre_f -= 0.5f * v[k];
              if (j + k < F) mn = v[k] < mn ? v[k] : mn;
            }
          } else {
            const float m01 = v[0] < v[1] ? v[0] : v[1], m23 = v[2] < v[3] ? v[2] : v[3];
            const float m4 = m01 < m23 ? m01 : m23;
            mn = m4 < mn ? m4 : mn;
          }
          if (MODE == 1) *reinterpret_cast<float4*>(sp + j) = make_float4(v[0], v[1], v[2], v[3]);
        }
      }
    }
    a_dirty = true;
    if (final_pass) pre_done = true;
    __syncthreads();  // the constants may be overwritten by the next group
    nf = 0;
  };

  // ================= Gaussian multiplets (ADAS406Lines / XLine): linemodels.py:401-403,428 ; :213-233 ============
  // The float64 parameters of every component (Doppler-shifted centre, 1 / sigma, amplitude, sample window) are
  // computed ONCE by one thread per component into a scratch table (the interior of buffer B, dead until the first
  // Balmer line or the instrument convolution); the block then adds the components point by point with
  // thread-consistent ownership j = tid (mod NT), so there is no barrier between components.
  {
    GaussComp* gc = reinterpret_cast<GaussComp*>(bufB + ((HALO + 7) & ~7));
    const int gcap = (((HALO + F) & ~7) - ((HALO + 7) & ~7)) * (int)sizeof(float) / (int)sizeof(GaussComp);
    const int total = a.n_gauss_comps;
    for (int base = 0; base < total; base += gcap) {
      const int nb = total - base < gcap ? total - base : gcap;
      for (int g = tid; g < nb; g += NT) {
        // This chain runs on a few threads while the rest of the block waits at the barrier below, so it is kept
        // short: the line object of component base + g is found by a branch-free scan (the table reads of different
        // lines overlap), the reciprocals that do not depend on the line record are formed while its loads are in
        // flight, and the remaining divisions are multiplications by them (relative difference ~1e-16).
        const int idx = base + g, n_lines = ch[CH_N_LINES], line_off = ch[CH_LINE_OFF];
        int k = 0, first = 0, run = 0;
        for (int kk = 0; kk < n_lines; ++kk) {
          const int* clk = m.cl_i + (line_off + kk) * CL_I_COUNT;
          const int nc = clk[CL_KIND] != 0 ? clk[CL_NCOMP] : 0;
          if (idx >= run && idx < run + nc) { k = kk; first = run; }
          run += nc;
        }
        const int* cl = m.cl_i + (line_off + k) * CL_I_COUNT;
        const double* cld = m.cl_d + (line_off + k) * CL_D_COUNT;
        const double* rec = lines_n + cl[CL_ULINE] * BAYSAR_LINE_STRIDE;
        const double* cp = m.comp_d + (cl[CL_COMP_OFF] + (idx - first)) * 2;
        const double rec0 = rec[0], rec4 = rec[4];
        const double inv_cal = 1.0 / cal, inv_delta = 1.0 / delta;
        double cw = cp[0], fwhm;
        if (cl[CL_KIND] == 1) {
          const double inv_mass = 1.0 / cld[CL_MASS];
          if (cl[CL_SPECIES] >= 0) {
            const int vi = m.sp_i[cl[CL_SPECIES] * SP_I_COUNT + SP_VEL_IDX];
            if (vi >= 0) cw = cw * ((1e3 * th[vi]) * (1.0 / c_light) + 1.0);
          }
          fwhm = cw * 7.715e-5 * sqrt(rec4 * inv_mass);
        } else {
          fwhm = cld[CL_FWHM];
        }
        if (!(fwhm > 0.0) || !(cp[1] > 0.0)) st |= 1u << 5;
        const double sigma = fwhm * fwhm_to_sigma;
        const double inv_sigma = 1.0 / sigma;
        const double amp = (rec0 * (cp[1] * (inv_sigma * (1.0 / 2.5066282746310002)))) * inv_cal;
        if (!(amp - amp == 0.0)) st |= 1u << 6;
        const double half = BAYSAR_Z_CUT * sigma;
        const double jl = ceil((cw - half - x0) * inv_delta - 1e-9), jh = floor((cw + half - x0) * inv_delta + 1e-9);
        GaussComp c;
        c.cw = cw; c.inv_sigma = inv_sigma; c.amp = (float)amp;
        c.j_lo = jl < 0.0 ? 0 : (jl > (double)F ? F : (int)jl);
        c.j_hi = jh < -1.0 ? -1 : (jh > (double)(F - 1) ? F - 1 : (int)jh);
        if (!(jl == jl) || !(jh == jh)) { c.j_lo = 0; c.j_hi = -1; }
        c.pad = 0;
        gc[g] = c;
      }
      __syncthreads();
      for (int q = 0; q < nb; ++q) {
        const GaussComp c = gc[q];
        int j = c.j_lo + ((tid - (c.j_lo % NT)) + NT) % NT;  // this thread's first owned point in the window
        for (; j <= c.j_hi; j += NT) {
          const double z = (grid_x(x0, delta, j) - c.cw) * c.inv_sigma;
          if (fabs(z) <= BAYSAR_Z_CUT) A_AT(j) += c.amp * exp_neg_f32(-0.5 * z * z);
        }
      }
      __syncthreads();
    }
    if (total > 0) a_dirty = true;
  }

  // ================= line objects ==============================================================================
  for (int k = 0; k < ch[CH_N_LINES]; ++k) {
    const int cls = line_class[k];
    if (cls == 0) continue;  // Gaussian multiplets: added above
    const int* cl = m.cl_i + (ch[CH_LINE_OFF] + k) * CL_I_COUNT;
    {
      // ---------------- Balmer line: Stark (x) Doppler ----------------
      if (cls != 2) {  // single / three Doppler taps: its shape record is ready; consecutive ones share a pair of sweeps
        if (nf == 0) fb = cl[CL_BRANK];
        if (++nf == BAYSAR_MAX_FUSED_LINES) flush_fused(false);
        continue;
      }
      if constexpr (TAPS < 2) {  // this build serves chords whose Balmer lines are guaranteed narrow (see TAPS above):
        st |= 1u << 16;           // the multi-tap path is compiled out of it
        continue;
      } else {
      const double* cld = m.cl_d + (ch[CH_LINE_OFF] + k) * CL_D_COUNT;
      const double* rec = lines_n + cl[CL_ULINE] * BAYSAR_LINE_STRIDE;
      const double cwl0 = cld[CL_CWL0];
      const double cwl = rec[13], sigma = rec[14];  // shifted centre and Doppler sigma from the LOS stage
      const int KD = cl[CL_KD], o = (KD - 1) / 2;
      const double dshift = cwl - cwl0;
      if (nf > 0) flush_fused(false);  // keep the (float) summation order of the lines deterministic
      if (!need_b) { st |= 1u << 16; continue; }  // cannot happen: the host guarantees single-tap lines (CH_B_OPTIONAL)
      // tap window |z| <= Z_CUT
      if (tid == 0) { iscr[0] = KD; iscr[1] = -1; }
      __syncthreads();
      for (int q = tid; q < KD; q += NT) {
        double xk = __dadd_rn(cld[CL_DOP_START], __dmul_rn((double)q, cld[CL_DOP_DELTA]));
        double z = ((xk + dshift) - cwl) / sigma;
        if (fabs(z) <= BAYSAR_Z_CUT_DOPPLER) { atomicMin(&iscr[0], q); atomicMax(&iscr[1], q); }
      }
      __syncthreads();
      const int k_lo = iscr[0], k_hi = iscr[1] + 1;
      const int T = k_hi - k_lo;
      __syncthreads();
      if (T <= 0) { st |= 1u << 6; continue; }  // uniform across the block
      for (int q = tid; q < T; q += NT) {
        double xk = __dadd_rn(cld[CL_DOP_START], __dmul_rn((double)(k_lo + q), cld[CL_DOP_DELTA]));
        double z = ((xk + dshift) - cwl) / sigma;
        tapd[q] = exp(-0.5 * z * z);
      }
      // q-indexed float taps for conv8: q = o - k in [o - k_hi + 1, o - k_lo]
      const int q_min = o - k_hi + 1, q_max = o - k_lo;
      const int q_start = q_min & ~3;
      const int nq = (q_max - q_start + 1 + 3) & ~3;
      __syncthreads();
      for (int qi = tid; qi < nq; qi += NT) {
        int kk = o - (q_start + qi);
        tapq[qi] = (kk >= k_lo && kk < k_hi) ? (float)tapd[kk - k_lo] : 0.f;
      }
      if (warp == 0) {  // exclusive prefix sums of the taps (float64), one contiguous chunk per lane
        int chunk = (T + 31) / 32, b = lane * chunk, e = b + chunk;
        e = e > T ? T : e;
        double s = 0.0;
        for (int q = b; q < e; ++q) s += tapd[q];
        double incl = s;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
          double v = __shfl_up_sync(0xffffffffu, incl, off);
          if (lane >= off) incl += v;
        }
        double run = incl - s;
        for (int q = b; q < e; ++q) { tappre[q] = run; run += tapd[q]; }
        if (lane == 31) tappre[T] = incl;
        // moments of the taps about the output point, mu_n = sum_u g_u (u delta)^n / sum_u g_u (u = kk - o), and from
        // them the coefficients of the far-field polynomials P_p(y) = sum_n mu_n (p)_n / n! y^n (see the far pass below)
        double mo[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        for (int q = lane; q < T; q += 32) {
          const double xu = (double)(k_lo + q - o) * delta;
          double w = tapd[q];
#pragma unroll
          for (int n_ = 0; n_ < 6; ++n_) { w *= xu; mo[n_] += w; }
        }
#pragma unroll
        for (int n_ = 0; n_ < 6; ++n_) mo[n_] = warp_sum(mo[n_]);
        const double ig = 1.0 / __shfl_sync(0xffffffffu, incl, 31);
        if (lane == 0) {
          float* fco = reinterpret_cast<float*>(scratch + 136);
          const double r25[6] = {2.5, 4.375, 6.5625, 9.0234375, 11.73046875, 14.6630859375};  // (2.5)_n / n!
          const double r50[4] = {5.0, 15.0, 35.0, 70.0};                                       // (5)_n / n!
          const double r75[2] = {7.5, 31.875};                                                 // (7.5)_n / n!
#pragma unroll
          for (int n_ = 0; n_ < 6; ++n_) fco[n_] = (float)(mo[n_] * ig * r25[n_]);
#pragma unroll
          for (int n_ = 0; n_ < 4; ++n_) fco[6 + n_] = (float)(mo[n_] * ig * r50[n_]);
#pragma unroll
          for (int n_ = 0; n_ < 2; ++n_) fco[10 + n_] = (float)(mo[n_] * ig * r75[n_]);
        }
      }
      __syncthreads();
      const double g_full = tappre[T];
      // interior points: every tap of the window lands on a valid output with full trapz weight
      const int i_lo = o - k_lo + 1;          // first interior i
      const int i_hi = F - 1 + o - k_hi;      // last interior i
      const float g25_rec = (float)rec[11], g25_exc = (float)rec[12];
      // lambda_j - cwl = (j - jc) * delta + r, with jc the grid point nearest the line centre and |r| <= delta / 2:
      // the large, cancelling parts are removed in float64 once, the per-point work is two float32 FMAs with a
      // float-float split of delta (error ~1 ulp of the difference itself, even next to the line centre).
      const double c0 = x0 - cwl;
      const int jc = (int)rint(-c0 / delta);
      const float r_f = (float)fma((double)jc, delta, c0);
      // Where the convolution is taken explicitly.  Far from the line centre the Stark pair is smooth on the scale of
      // the Doppler taps, and sum_u g_u m(x - u delta) = sum_n (-1)^n G_n m^(n)(x) / n! with the raw tap moments G_n;
      // for the far-field series m = c0 |d|^-2.5 - c1 |d|^-5 + c2 |d|^-7.5 every derivative is again a power of |d|, so
      //   (m (*) g)(x) = G_0 ia (c0 P_2.5(y) - ia (c1 P_5(y) - ia c2 P_7.5(y))),  ia = |d|^-2.5,  y = 1 / d,
      //   P_p(y) = sum_n mu_n (p)_n / n! y^n  (n <= 6, 4, 2),
      // one rsqrt and ~20 FMAs per point instead of two passes and T taps (relative error < 1e-8 beyond
      // BAYSAR_FAR_SIGMAS Doppler widths and the far ratio of the series; tests/test_kernel_model.py).  Explicit stay a
      // window around the centre and the two end zones where the taps leave the grid (the reference's "same"
      // convolution pads with zeros); when they would cover most of the grid anyway the whole grid is explicit.
      const int u_min = k_lo - o, u_max = k_hi - 1 - o;  // tap offsets: out[j] = sum_u g_u m[j - u]
      const int ext = (u_max > -u_min ? u_max : -u_min) + 8;
      int sa[3], sb[3];  // explicit segments [sa, sb] (empty: sb < sa), ascending, block-uniform
      bool far_on = false;
      {
        const float gmx = g25_rec > g25_exc ? g25_rec : g25_exc;
        const float wsig = (float)BAYSAR_FAR_SIGMAS * (float)sigma, wrat = __powf(BAYSAR_FAR_RATIO * gmx, 0.4f) * 1.001f;
        float wf = ceilf((wsig > wrat ? wsig : wrat) / (float)delta) + 2.f;
        wf = wf < (float)F ? wf : (float)F;  // NaN -> F: everything explicit
        const int Wp = wf == wf ? (int)wf : F;
        // end zones: outputs whose taps reach beyond the grid, and inputs whose taps do (the weighted zone of pass 1)
        const int ez = (u_max > -u_min ? u_max : -u_min) + 1;
        int a0 = 0, b0 = ez;                                                 // left end zone
        int a1 = jc - Wp < 0 ? 0 : jc - Wp, b1 = jc + Wp > F - 1 ? F - 1 : jc + Wp;  // window (empty if off the grid)
        int a2 = F - 1 - ez, b2 = F - 1;                                     // right end zone
        if (b0 > F - 1) b0 = F - 1;
        if (a2 < 0) a2 = 0;
        // whole blocks of eight outputs: the convolution's blocks tile a segment exactly and no 128-bit group of the far
        // pass straddles a boundary, so the two passes touch disjoint elements of the spectrum and need no barrier
        b0 |= 7; a1 &= ~7; b1 |= 7; a2 &= ~7; b2 |= 7;
        // merge neighbours whose pass-2 ranges (segment +- ext) would overlap; an empty segment has sb < sa
        // (scalars and fully unrolled loops only: a dynamically indexed array would live in local memory)
        if (b1 >= a1 && b0 >= a0 && a1 <= b0 + 2 * ext + 1) { b0 = b1 > b0 ? b1 : b0; b1 = a1 - 1; }
        if (b2 >= a2) {
          if (b1 >= a1) { if (a2 <= b1 + 2 * ext + 1) { b1 = b2 > b1 ? b2 : b1; b2 = a2 - 1; } }
          else if (b0 >= a0 && a2 <= b0 + 2 * ext + 1) { b0 = b2 > b0 ? b2 : b0; b2 = a2 - 1; }
        }
        sa[0] = a0; sb[0] = b0; sa[1] = a1; sb[1] = b1; sa[2] = a2; sb[2] = b2;
        int covered = 0;
#pragma unroll
        for (int q = 0; q < 3; ++q) covered += sb[q] >= sa[q] ? sb[q] - sa[q] + 1 + 2 * ext : 0;
        far_on = T >= 12 && covered < (F * 3) / 5;
        if (!far_on) { sa[0] = 0; sb[0] = F - 1; sb[1] = sa[1] - 1; sb[2] = sa[2] - 1; }
      }
      // pass 1: a25 = |d|^2.5 stashed in B;  A_e = sum_i u_e[i] G[i] with u_rec = den_exc * inv, u_exc = den_rec * inv,
      // inv = 1 / (den_rec * den_exc): one reciprocal serves both profiles.  Point by point on the explicit segments (the
      // stash also on the tap reach around them); the far parts in between are interior (weight G_0) and their sums
      // are closed forms: sum_{j=A}^{B} f = trapz + (f_A + f_B) / 2 with the tail integral of sweep 1.
      float sr = 0.f, se = 0.f;
      double er = 0.0, ee = 0.0;
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        if (sb[q] < sa[q]) continue;
        const int lo_ = sa[q] - ext < 0 ? 0 : sa[q] - ext, hi_ = sb[q] + ext > F - 1 ? F - 1 : sb[q] + ext;
        for (int j = lo_ + tid; j <= hi_; j += NT) {
          const float kf = (float)(j - jc);
          float d = fabsf(fmaf(kf, d_hi, fmaf(kf, d_lo, r_f)));
          float a25 = d * d * sqrt_approx(d);
          B_AT(j) = a25;
          if (j < sa[q] || j > sb[q]) continue;  // tap reach only: its sum belongs to a far part
          float dr = a25 + g25_rec, de = a25 + g25_exc;
          float inv = rcp_approx(dr * de);
          float ur = de * inv, ue = dr * inv;
          if (j >= i_lo && j <= i_hi) { sr += ur; se += ue; }
          else {
            int kmin = o - j, kmax = F - 1 + o - j;  // taps that reach outputs 0 and F-1
            int lo = kmin > k_lo ? kmin : k_lo, hi = kmax + 1 < k_hi ? kmax + 1 : k_hi;
            double G = 0.0;
            if (hi > lo) {
              G = tappre[hi - k_lo] - tappre[lo - k_lo];
              if (kmin >= lo && kmin < hi) G -= 0.5 * tapd[kmin - k_lo];
              if (kmax >= lo && kmax < hi) G -= 0.5 * tapd[kmax - k_lo];
            }
            er += (double)ur * G; ee += (double)ue * G;
          }
        }
      }
      if (far_on && tid < 8) {  // far part tid >> 1 (the gaps before, between and after the segments), profile tid & 1
        int pa = 0, pb = -1, prev = -1, np_ = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const bool seg = q < 3 && sb[q < 3 ? q : 0] >= sa[q < 3 ? q : 0];
          if (q < 3 && !seg) continue;
          const int ga = prev + 1, gb = q < 3 ? sa[q] - 1 : F - 1;
          if (gb >= ga) { if (np_ == (tid >> 1)) { pa = ga; pb = gb; } ++np_; }
          if (q < 3) prev = sb[q];
        }
        double far_sum = 0.0;
        if (pb >= pa) {
          const float g = (tid & 1) ? g25_exc : g25_rec;
          float fe2[2];
          const float tz = pb > pa ? stark_tail_trapz(pa, pb, (float)jc, r_f, g, d_hi, d_lo, (float)delta, fe2) : 0.f;
          if (pb == pa) {
            const float kf = (float)(pa - jc), d = fabsf(fmaf(kf, d_hi, fmaf(kf, d_lo, r_f)));
            fe2[0] = fe2[1] = rcp_approx(d * d * sqrt_approx(d) + g);  // a single point: the sum is its value
          }
          far_sum = (double)tz + 0.5 * ((double)fe2[0] + (double)fe2[1]);
        }
        scratch[150 + tid] = far_sum;
      }
      double red[2] = {(double)sr * g_full + er, (double)se * g_full + ee};
      block_sum<NT, 2>(red, scratch);
      if (far_on) {
        red[0] += g_full * (scratch[150] + scratch[152] + scratch[154] + scratch[156]);
        red[1] += g_full * (scratch[151] + scratch[153] + scratch[155] + scratch[157]);
      }
      const double s_rec = rec[0] / (dl * red[0]) / cal, s_exc = rec[1] / (dl * red[1]) / cal;
      if (!(s_rec - s_rec == 0.0) || !(s_exc - s_exc == 0.0)) st |= 1u << 6;
      const float fr = (float)s_rec, fe = (float)s_exc;
      // pass 2: m = s_rec u_rec + s_exc u_exc = (s_rec den_exc + s_exc den_rec) * inv  (in place over the stash), on the
      // explicit segments and the tap reach around them
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        if (sb[q] < sa[q]) continue;
        const int lo = sa[q] - ext < 0 ? 0 : sa[q] - ext, hi = sb[q] + ext > F - 1 ? F - 1 : sb[q] + ext;
        for (int j = lo + tid; j <= hi; j += NT) {
          float a25 = B_AT(j);
          float dr = a25 + g25_rec, de = a25 + g25_exc;
          B_AT(j) = fmaf(fr, de, fe * dr) * rcp_approx(dr * de);
        }
      }
      __syncthreads();
      // A += m (*) g on the explicit segments
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        if (sb[q] < sa[q]) continue;
        const int ja = sa[q], jb = sb[q] < F - 1 ? sb[q] : F - 1;
        for (int j0 = ja + 8 * tid; j0 <= jb; j0 += 8 * NT) {
          float acc[8];
          conv8(bufB, HALO + j0 + q_start, tapq, nq, acc);
          float4* dst = reinterpret_cast<float4*>(bufA + pad_index(HALO + j0));
          float4 v = dst[0];
          v.x += acc[0]; v.y += acc[1]; v.z += acc[2]; v.w += acc[3];
          dst[0] = v;
          dst = reinterpret_cast<float4*>(bufA + pad_index(HALO + j0 + 4));
          v = dst[0];
          v.x += acc[4]; v.y += acc[5]; v.z += acc[6]; v.w += acc[7];
          dst[0] = v;
        }
      }
      if (far_on) {  // (no barrier: disjoint elements, see above)
        const float* fco = reinterpret_cast<const float*>(scratch + 136);
        // coefficients in registers (the loop writes shared memory: the compiler would re-read them every point); the
        // amplitudes and G_0 folded in: value = ia (q1(y) - ia (q2(y) - ia q3(y)))
        const float gf = (float)g_full;
        const float c0f = gf * (fr + fe), c1f = gf * fmaf(fr, g25_rec, fe * g25_exc),
                    c2f = gf * fmaf(fr * g25_rec, g25_rec, fe * g25_exc * g25_exc);
        const float a0 = c0f, a1 = c0f * fco[0], a2 = c0f * fco[1], a3 = c0f * fco[2], a4 = c0f * fco[3], a5 = c0f * fco[4],
                    a6 = c0f * fco[5];
        const float e0 = c1f, e1 = c1f * fco[6], e2 = c1f * fco[7], e3 = c1f * fco[8], e4 = c1f * fco[9];
        const float h0 = c2f, h1 = c2f * fco[10], h2 = c2f * fco[11];
        // four consecutive points per thread and trip (128-bit read-modify-write of the spectrum); a group that touches an
        // explicit segment is handled point by point
        for (int j0 = 4 * tid; j0 < F; j0 += 4 * NT) {
          if ((j0 >= sa[0] && j0 <= sb[0]) || (j0 >= sa[1] && j0 <= sb[1]) || (j0 >= sa[2] && j0 <= sb[2])) continue;
          const bool clear = j0 + 3 < F;
          float4* ap = reinterpret_cast<float4*>(bufA + pad_index(HALO + j0));
          float4 v4 = *ap;
          float v[4] = {v4.x, v4.y, v4.z, v4.w};
          const float kf0 = (float)(j0 - jc);
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const int j = j0 + r;
            if (!clear && j >= F) continue;
            const float kf2 = kf0 + (float)r;
            const float dd = fmaf(kf2, d_hi, fmaf(kf2, d_lo, r_f));
            const float rs = rsqrt_approx(fabsf(dd)), iq = rs * rs;
            const float y = dd < 0.f ? -iq : iq, ia = iq * iq * rs;
            const float q1 = fmaf(y, fmaf(y, fmaf(y, fmaf(y, fmaf(y, fmaf(y, a6, a5), a4), a3), a2), a1), a0);
            const float q2 = fmaf(y, fmaf(y, fmaf(y, fmaf(y, e4, e3), e2), e1), e0);
            const float q3 = fmaf(y, fmaf(y, h2, h1), h0);
            v[r] = fmaf(ia, fmaf(-ia, fmaf(-ia, q3, q2), q1), v[r]);
          }
          *ap = make_float4(v[0], v[1], v[2], v[3]);
        }
      }
      __syncthreads();  // the convolution's last block may have added to the alignment tail
      for (int j = F + tid; j < f8; j += NT) A_AT(j) = 0.f;  // keep the alignment tail clean
      a_dirty = true;
      __syncthreads();
      }  // TAPS == 2 build
    }
  }
  if (nf > 0) flush_fused(true);
  if (!need_b && !a_dirty) {  // no line wrote the interior (a chord without lines, or every line flagged): it is zero
    for (int j = tid; j < 4 * ((F + 3) >> 2); j += NT) A_AT(j) = 0.f;
  }
  __syncthreads();

  // ================= pre-convolution area and minimum (spectrometers.py:256, :271) ===============================
  if (!pre_done) {
    for (int j = tid; j < F; j += NT) {
      float v = A_AT(j);
      pre_f += (j == 0 || j == F - 1) ? 0.5f * v : v;   // <= 40 terms per thread in float32, float64 across threads
      mn = v < mn ? v : mn;
      if (!(v - v == 0.f)) bad = 1;
    }
  }
  double pre_s = (double)pre_f;
  if (a.fine_pre) {
    float* o_ = a.fine_pre + n * (int64_t)F;
    for (int j = tid; j < F; j += NT) o_[j] = A_AT(j);
  }
  mn = warp_min_f(mn);
  if (lane == 0) reinterpret_cast<float*>(iscr)[4 + warp] = mn;  // iscr[4 .. 4 + NT/32)
  double red1[1] = {pre_s};
  block_sum<NT, 1>(red1, scratch);
  for (int w = 0; w < NT / 32; ++w) { float o_ = reinterpret_cast<float*>(iscr)[4 + w]; mn = o_ < mn ? o_ : mn; }
  const double pre = dl * red1[0];
  if (__syncthreads_or(bad)) st |= 1u << 6;

  if (MODE == 1 && a.skip_edge64) {
    // ================= hand the spectrum to the folded tensor-core contraction (no float64 end outputs needed) =====
    float* sp = a.spec + n * a.split_pitch + a.split_pad;
    if (!a.keep_padding) {
      for (int j = tid - a.split_pad; j < 0; j += NT) sp[j] = 0.f;
      for (int64_t j = F + tid; j < a.split_pitch - a.split_pad; j += NT) sp[j] = 0.f;
    }
    if (!pre_done) for (int j = tid; j < F; j += NT) sp[j] = A_AT(j);
    if (tid == 0) {
      double* ax = a.aux + n * BAYSAR_AUX_STRIDE;
      ax[0] = pre; ax[1] = (double)mn; ax[2] = 0.0; ax[3] = 0.0; ax[4] = 0.0; ax[5] = 0.0; ax[6] = 0.0; ax[7] = 0.0;
    }
    st = __syncthreads_or((int)st);
    if (tid == 0 && st) atomicOr(&a.status[n], st);
    return;
  }

  // ================= instrument function taps (spectrometers.py:229-253) ========================================
  int KI, k_lo, k_hi;
  if (ch[CH_CALINT_IDX] >= 0) {
    // sampled Gaussian instrument function over the whole fine grid (plasmas.py:96-98), pixel normalised
    KI = F;
    const double fw = pow(10.0, th[ch[CH_CALINT_IDX]]);
    const double sigma = fw * fwhm_to_sigma, xm = chd[CH_X_MEAN];
    if (!(fw > 0.0)) st |= 1u << 5;
    if (tid == 0) { iscr[0] = F; iscr[1] = -1; }
    __syncthreads();
    double gs = 0.0;
    for (int q = tid; q < F; q += NT) {
      double z = (grid_x(x0, delta, q) - xm) / sigma;
      if (fabs(z) <= BAYSAR_Z_CUT) { atomicMin(&iscr[0], q); atomicMax(&iscr[1], q); gs += exp(-0.5 * z * z); }
    }
    __syncthreads();
    k_lo = iscr[0]; k_hi = iscr[1] + 1;
    double red2[1] = {gs};
    block_sum<NT, 1>(red2, scratch);
    if (!(red2[0] > 0.0)) st |= 1u << 7;
    int o = (KI - 1) / 2;
    // clamp the window to the shared-memory capacity sized from the parameter bounds
    int lim = HALO - 8 < TAPS_MAX / 2 - 8 ? HALO - 8 : TAPS_MAX / 2 - 8;
    if (k_hi - 1 - o > lim) { k_hi = o + lim + 1; st |= 1u << 16; }
    if (o - k_lo > lim) { k_lo = o - lim; st |= 1u << 16; }
    if (k_hi <= k_lo) { k_lo = o; k_hi = o + 1; }
    const int q_min = o - k_hi + 1, q_max = o - k_lo;
    const int q_start = q_min & ~3;
    const int nq = (q_max - q_start + 1 + 3) & ~3;
    for (int qi = tid; qi < nq; qi += NT) {
      int kk = o - (q_start + qi);
      float w = 0.f;
      if (kk >= k_lo && kk < k_hi) {
        double z = (grid_x(x0, delta, kk) - xm) / sigma;
        w = (float)(exp(-0.5 * z * z) / red2[0]);
      }
      tapq[qi] = w;
    }
  } else {
    KI = ch[CH_KI]; k_lo = ch[CH_IF_KLO]; k_hi = ch[CH_IF_KHI];
    const int o = (KI - 1) / 2;
    const int q_min = o - k_hi + 1, q_max = o - k_lo;
    const int q_start = q_min & ~3;
    const int nq = (q_max - q_start + 1 + 3) & ~3;
    const double* w = m.ifn + ch[CH_IF_OFF];
    for (int qi = tid; qi < nq; qi += NT) {
      int kk = o - (q_start + qi);
      tapq[qi] = (kk >= k_lo && kk < k_hi) ? (float)w[kk] : 0.f;
    }
  }
  const int o_if = (KI - 1) / 2;
  const int q_min_if = o_if - k_hi + 1, q_max_if = o_if - k_lo;
  const int q_start_if = q_min_if & ~3;
  const int nq_if = (q_max_if - q_start_if + 1 + 3) & ~3;
  __syncthreads();

  // the four end outputs in float64: they set the slope interp1d extrapolates with (sampled wavelength maps)
  if (warp < 4) {
    int j = warp < 2 ? warp : F - 4 + warp;
    double s = 0.0;
    for (int qi = lane; qi < nq_if; qi += 32) s += (double)tapq[qi] * (double)A_AT(j + q_start_if + qi);
    s = warp_sum(s);
    if (lane == 0) edge64[warp] = s < (double)mn ? (double)mn : s;
  }

  if (MODE == 1) {
    // ================= hand the spectrum to the tensor-core contraction =========================================
    __syncthreads();
    float* sp = a.spec + n * a.split_pitch + a.split_pad;
    // the workspace is shared by chords with different paddings: (re)write this row's zero padding as well
    if (!a.keep_padding) {
      for (int j = tid - a.split_pad; j < 0; j += NT) sp[j] = 0.f;
      for (int64_t j = F + tid; j < a.split_pitch - a.split_pad; j += NT) sp[j] = 0.f;
    }
    if (!pre_done) for (int j = tid; j < F; j += NT) sp[j] = A_AT(j);
    if (tid == 0) {
      double* ax = a.aux + n * BAYSAR_AUX_STRIDE;
      ax[0] = pre; ax[1] = (double)mn; ax[2] = edge64[0]; ax[3] = edge64[1]; ax[4] = edge64[2]; ax[5] = edge64[3];
      ax[6] = 0.0; ax[7] = 0.0;
    }
    st = __syncthreads_or((int)st);
    if (tid == 0 && st) atomicOr(&a.status[n], st);
    return;
  }

  // ================= instrument convolution, clip at the pre-convolution minimum (spectrometers.py:269-271) =====
  double post_s = 0.0;
  bad = 0;
  for (int j0 = 8 * tid; j0 < F; j0 += 8 * NT) {
    float acc[8];
    conv8(bufA, HALO + j0 + q_start_if, tapq, nq_if, acc);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      acc[q] = acc[q] < mn ? mn : acc[q];
      int j = j0 + q;
      if (j < F) {
        double w = (j == 0 || j == F - 1) ? 0.5 : 1.0;
        post_s += w * (double)acc[q];
        if (!(acc[q] - acc[q] == 0.f)) bad = 1;
      }
    }
    *reinterpret_cast<float4*>(bufB + pad_index(HALO + j0)) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    *reinterpret_cast<float4*>(bufB + pad_index(HALO + j0 + 4)) = make_float4(acc[4], acc[5], acc[6], acc[7]);
  }
  if (tid == 0) scratch[BAYSAR_BG_SLOT] = pow(10.0, th[ch[CH_BG_IDX]]);
  double red3[1] = {post_s};
  block_sum<NT, 1>(red3, scratch);
  if (__syncthreads_or(bad)) st |= 1u << 8;
  const double post = dl * red3[0];
  FetchSmem fetch{bufB, HALO};
  pixel_likelihood<NT>(m, ch, chd, th, n, c, ncomp_cols, pre, post, scratch[BAYSAR_BG_SLOT], edge64, fetch, scratch,
                       a.components, a.spectra, a.fine_post, st);
  st = __syncthreads_or((int)st);
  if (tid == 0 && st) atomicOr(&a.status[n], st);
#undef A_AT
#undef B_AT
}

// K2c — pixel / likelihood stage of the tensor-core path: one CTA per (theta, chord).  Reads the convolved spectrum
// the contraction kernel left in global memory (L2-resident), its per-tile trapz partials (summed in tile order:
// deterministic) and the scalars of the spectral stage.
struct PixelArgs {
  const double* theta;
  int64_t n;
  int chord;
  const float* conv;        // [rows][conv_pitch]: every column, or (cmask != nullptr) the compact columns
  int64_t conv_pitch;
  const uint32_t* cmask;
  const int* cbase;
  const double* trapz_part; // [rows][n_tiles]
  int n_tiles;
  const double* aux;        // [N][BAYSAR_AUX_STRIDE]
  double* components;
  double* spectra;          // optional [N][P]
  float* fine_post;         // optional [N][F]
  uint32_t* status;
};

template <int NT>
__global__ void __launch_bounds__(NT) pixel_stage_kernel(ModelDev m, PixelArgs a) {
#ifdef BAYSAR_CPU_EMU
  double* scratch = reinterpret_cast<double*>(emu_dynamic_smem());
#else
  __shared__ double scratch[64];
#endif
  const int tid = threadIdx.x;
  const int64_t n = blockIdx.x;
  const int c = a.chord;
  const int64_t* I = m.I;
  const int n_params = (int)I[I_N_PARAMS];
  const int ncomp_cols = (int)I[I_N_CHORDS] + (int)I[I_N_PRIORS];
  const int* ch = m.ch_i + c * CH_I_COUNT;
  const double* chd = m.ch_d + c * CH_D_COUNT;
  const double* th = a.theta + n * n_params;
  const double* ax = a.aux + n * BAYSAR_AUX_STRIDE;
  uint32_t st = 0;
  // trapz of the convolved spectrum: the contraction's per-tile partial sums, one per thread, reduced in a fixed
  // order (deterministic)
  const double* part = a.trapz_part + n * (int64_t)a.n_tiles;
  if (tid == 0) scratch[BAYSAR_BG_SLOT] = pow(10.0, th[ch[CH_BG_IDX]]);
  double red_p[1] = {0.0};
  for (int t = tid; t < a.n_tiles; t += NT) red_p[0] += part[t];
  block_sum<NT, 1>(red_p, scratch);
  const double bg = scratch[BAYSAR_BG_SLOT];
  const double post_s = red_p[0];
  if (!(post_s - post_s == 0.0)) st |= 1u << 8;
  const double post = chd[CH_DL] * post_s;
  if (a.cmask) {
    FetchCompact fetch{a.conv + n * a.conv_pitch, a.cmask, a.cbase};
    pixel_likelihood<NT>(m, ch, chd, th, n, c, ncomp_cols, ax[0], post, bg, ax + 2, fetch, scratch, a.components,
                         a.spectra, nullptr, st);
  } else {
    FetchGlobal fetch{a.conv + n * a.conv_pitch};
    pixel_likelihood<NT>(m, ch, chd, th, n, c, ncomp_cols, ax[0], post, bg, ax + 2, fetch, scratch, a.components,
                         a.spectra, a.fine_post, st);
  }
  st = __syncthreads_or((int)st);
  if (tid == 0 && st) atomicOr(&a.status[n], st);
}

// K2c' — pixel / likelihood stage behind pix_contract.cuh: one CTA per (theta, chord).
// The trapz of the clipped, convolved spectrum is not summed from fine-grid columns (they are never formed) but from
// the identity  sum_j c_j conv_j = sum_s S_s G_s,  G_s = sum_j c_j w[j + o - s]  (c = trapz weights):  G equals the
// tap sum W times c_s except within a tap half-width of the two ends, so
//     post / dl = W pre / dl + sum_{s in end zones} S_s (G_s - W c_s) + sum_{j in edge zones} c_j (clip(conv_j) - conv_j)
// with the explicit edge columns of the contraction for the last sum (the clip is inactive elsewhere: w >= 0).
// Bias correction: tensor-memory accumulation truncates, so every folded column comes out low by a few 1e-7 .. 1e-6
// (relative).  The reference's renormalisation by the convolved area cancelled the common part of that in the
// every-column path; here the same is achieved per column tile with the column-sum identity
//     sum_{c in tile} C_c = sum_s S_s bbar_s,   bbar_s = sum_{c in tile} B[s, c]
// whose right-hand side the contraction's converter threads accumulate from the A operand they hold anyway (float32
// products, float64 sums): kappa_tile = (exact sum) / (sum of the contraction's columns) multiplies the tile's columns.
struct PixelFoldArgs {
  const double* theta;
  int64_t n;
  int chord;
  const float* cp;          // [rows][cp_pitch] contraction output
  int64_t cp_pitch;
  const int* pix_col;       // [P]
  int col_left0, x_l, col_right0, x_r;
  const float* spec;        // [rows][spec_pitch] pre-convolution spectra, element s at spec_pad + s
  int64_t spec_pitch;
  int spec_pad;
  const double* dtab;       // [n_dl + n_dr]: G_s - W c_s for s = 0 .. n_dl - 1, then s = F - n_dr .. F - 1
  int n_dl, n_dr;
  double wsum;              // W
  const float* inv_err;     // [P] float32 1 / err of the chord's pixels (log-posterior-only fast loop), may be nullptr
  const double* tsum;       // [rows][n_col_tiles][pxc::TSUM_SLOTS] exact column sums per tile (converter groups), may be nullptr
  int n_col_tiles, n_fold_tiles, tile_n;
  double* logp;             // log-posterior path of single-chord models: also do K3's job (sum of components, flags)
  const double* aux;        // [N][BAYSAR_AUX_STRIDE]
  int use_edge64;           // 1: aux carries the four float64 end outputs of the convolution (some pixel extrapolates)
  double* components;
  double* spectra;          // optional [N][P]
  uint32_t* status;
};

// Register budget: the kernel waits on global-load latency (row staging, per-pixel tables), so resident CTAs matter more
// than registers: 14 CTAs of 128 threads per SM (32 registers, ~100 bytes of spills) run the 4096-theta demo step in 2 waves
// instead of 4 at the compiler's own 64 registers (measured 49.7 -> 45.1 us at 10 CTAs -> 40.6 us at 14).
#ifndef BAYSAR_PIXFOLD_MINB
#define BAYSAR_PIXFOLD_MINB 14
#endif
template <int NT>
__global__ void __launch_bounds__(NT, BAYSAR_PIXFOLD_MINB) pixel_fold_kernel(ModelDev m, PixelFoldArgs a) {
  pdl_trigger();
  pdl_wait();  // the contraction's columns and tile sums
#ifdef BAYSAR_CPU_EMU
  double* scratch = reinterpret_cast<double*>(emu_dynamic_smem());
  float* row = reinterpret_cast<float*>(scratch + 160);
#else
  __shared__ double scratch[64 + 32 + 64];  // reductions, 64 float tile corrections, 64 float64 tile scales
  extern __shared__ __align__(16) float row[];  // this theta's contraction output (cp_pitch floats)
#endif
  const int tid = threadIdx.x;
  const int64_t n = blockIdx.x;
  const int c = a.chord;
  const int64_t* I = m.I;
  const int n_params = (int)I[I_N_PARAMS];
  const int ncomp_cols = (int)I[I_N_CHORDS] + (int)I[I_N_PRIORS];
  const int* ch = m.ch_i + c * CH_I_COUNT;
  const double* chd = m.ch_d + c * CH_D_COUNT;
  const double* th = a.theta + n * n_params;
  const double* ax = a.aux + n * BAYSAR_AUX_STRIDE;
  const int F = ch[CH_F];
  const float mn = (float)ax[1];
  const float* sp = a.spec + n * a.spec_pitch + a.spec_pad;
  uint32_t st = 0;
  // Stage the row in shared memory with one round of coalesced 16-byte loads: everything below reads it several times
  // (edge columns, tile sums, pixels through the pixel -> column map), and as dependent global loads those reads
  // made the kernel latency-bound.
  {
    const float4* src = reinterpret_cast<const float4*>(a.cp + n * a.cp_pitch);
    float4* dst = reinterpret_cast<float4*>(row);
    for (int i = tid; i < (int)(a.cp_pitch / 4); i += NT) dst[i] = src[i];
  }
  double part = 0.0;
  for (int t = tid; t < a.n_dl + a.n_dr; t += NT) {  // end zones of the pulled-back trapz weights
    const int s = t < a.n_dl ? t : F - a.n_dr + (t - a.n_dl);
    part += (double)sp[s] * a.dtab[t];
  }
  if (tid == 0) scratch[BAYSAR_BG_SLOT] = pow(10.0, th[ch[CH_BG_IDX]]);
  __syncthreads();
  for (int t = tid; t < a.x_l + a.x_r; t += NT) {  // clip correction of the explicit edge columns
    const int j = t < a.x_l ? t : F - a.x_r + (t - a.x_l);
    const float v = t < a.x_l ? row[a.col_left0 + t] : row[a.col_right0 + (t - a.x_l)];
    if (!(v - v == 0.f)) st |= 1u << 8;
    if (v < mn) part += ((j == 0 || j == F - 1) ? 0.5 : 1.0) * ((double)mn - (double)v);
  }
  double red_p[1] = {part};
  block_sum<NT, 1>(red_p, scratch);
  const double bg = scratch[BAYSAR_BG_SLOT];
  const double pre = ax[0];
  const double post = a.wsum * pre + chd[CH_DL] * red_p[0];
  if (!(post - post == 0.0)) st |= 1u << 8;
  // per-tile bias correction: one warp per folded column tile
  float* kappa_s = reinterpret_cast<float*>(scratch + 64);
  const int tile_shift = a.tile_n == 64 ? 6 : 7;
  for (int t = tid >> 5; t < a.n_fold_tiles; t += NT / 32) {
    float kap = 1.f;
    if (a.tsum) {
      double cs = 0.0;
      for (int e = tid & 31; e < a.tile_n; e += 32) cs += (double)row[t * a.tile_n + e];
      cs = warp_sum(cs);
      const double* ts = a.tsum + (n * a.n_col_tiles + t) * 3;  // pxc::TSUM_SLOTS
      const double k = ((ts[0] + ts[1]) + ts[2]) / cs;
      if (fabs(k - 1.0) < 1e-4) kap = (float)k;  // a correction, not a repair: anything larger is not truncation bias
    }
    if ((tid & 31) == 0) kappa_s[t] = kap;
  }
  __syncthreads();
  FetchFolded fetch{row, a.pix_col, a.col_left0, a.x_l, a.col_right0, a.x_r, F, mn, kappa_s, tile_shift, a.use_edge64};
  if (a.spectra || !a.inv_err) {
    pixel_likelihood<NT>(m, ch, chd, th, n, c, ncomp_cols, pre, post, bg, ax + 2, fetch, scratch, a.components, a.spectra,
                         nullptr, st);
  } else {
    // log-posterior only: the same arithmetic as pixel_likelihood with the per-pixel constants folded (tile scale
    // kappa / ratio in shared memory, float32 1 / err table) — 4 M pixels per step make this loop a kernel of its own
    const double ratio = post / pre, inv_ratio = 1.0 / ratio;
    const double ratio2 = (post / ratio) / pre;
    if (!(fabs(ratio2 - 1.0) <= 1e-8 + 1e-2)) st |= 1u << 9;
    double* scale_s = scratch + 96;
    if (tid < a.n_fold_tiles) scale_s[tid] = (double)kappa_s[tid] * inv_ratio;
    __syncthreads();
    const int P = ch[CH_P], poff = ch[CH_PIX_OFF];
    double acc = 0.0;
    const bool gauss = m.I[I_GAUSS_LIKELIHOOD] != 0;
#pragma unroll 4
    for (int p = tid; p < P; p += NT) {
      const int col = a.pix_col[p];
      double fm;
      if (col >= 0) {
        fm = (double)row[col] * scale_s[col >> tile_shift] + bg;
      } else {
        fm = pixel_value(fetch, p, m.pix_idx[poff + p], m.pix_frac[poff + p], ax + 2, F) * inv_ratio + bg;
      }
      fm = fm < bg ? bg : fm;
      const double d = m.y[poff + p] - fm;
      if (!(d - d == 0.0)) st |= 1u << 10;
      const float rf = (float)d * a.inv_err[p];
      const float ar = fabsf(rf);
      if (gauss) {  // likelihood(cauchy=False), spectrometers.py:187-188
        const double r = d / m.err[poff + p], r2 = r * r;
        if (r2 == 0.0) st |= 1u << 11;
        if (!(r2 - r2 == 0.0)) st |= 1u << 12;
        acc += 0.5 * r2;
      } else if (ar < 1e15f) {
        if (d == 0.0 || ar == 0.f) st |= 1u << 11;
        acc += (double)log1pf(ar * ar);
      } else {
        const double r = d / m.err[poff + p], r2 = r * r;
        if (r2 == 0.0) st |= 1u << 11;
        if (!(r2 - r2 == 0.0)) st |= 1u << 12;
        acc += log(1.0 + r2);
      }
    }
    double red4[1] = {acc};
    block_sum<NT, 1>(red4, scratch);
    if (tid == 0) a.components[n * (int64_t)ncomp_cols + c] = -red4[0];
    st = __syncthreads_or((int)st);
    if (tid == 0) {
      if (a.logp) {  // single-chord model: logP = sum of the components in list order, output flags (finalize.cuh)
        const double* comp = a.components + n * (int64_t)ncomp_cols;
        double s = 0.0;
        for (int k = 0; k < ncomp_cols; ++k) s = s + (k == c ? -red4[0] : comp[k]);
        if (s != s) st |= 1u << 13;
        else if (s - s != 0.0) st |= 1u << 14;
        else if (s >= 0.0) st |= 1u << 15;
        a.logp[n] = s;
      }
      if (st) atomicOr(&a.status[n], st);
    }
    return;
  }
  st = __syncthreads_or((int)st);
  if (tid == 0 && st) atomicOr(&a.status[n], st);
}
