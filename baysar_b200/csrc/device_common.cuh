// Shared device helpers: model view, searches, table interpolation, reductions.
#pragma once
#include <stdint.h>

#include "layout.h"

#ifndef BAYSAR_CPU_EMU
#include <cuda_runtime.h>
#define BAYSAR_DEV __device__ __forceinline__
#else
#include "cpu_emu.h"  // tests/emu: std::thread-based execution of the same kernel source (test-only)
#endif

#define BAYSAR_LINE_STRIDE 16
#define BAYSAR_HDR_STRIDE 4  // per-theta header written by the LOS stage: ne_max, reserved...
#define BAYSAR_Z_CUT 8.85    // exp(-z*z/2) < 1e-17 beyond this

// Device-side view of one compiled model: pointers into the single HBM copy of the descriptor blob.
struct ModelDev {
  const int64_t* I;
  const double* D;
  const double* los_x;
  const double* los_w;
  const double* prof_d;
  const int* sp_i;
  const double* sp_d;
  const int* elem_dens_idx;
  const double* a11_ne;
  const double* a11_te;
  const double* a11_scd;
  const double* h_ne;
  const double* h_te;
  const double* h_tab;
  const double* tau_grid;
  const double* spl_tx;
  const double* spl_ty;
  const double* spl_c;
  const int* ul_i;
  const double* ul_d;
  const int* pr_i;
  const double* pr_d;
  const int* cso_off;
  const int* cso_ul;
  const int* ch_i;
  const double* ch_d;
  const int* cl_i;
  const double* cl_d;
  const double* comp_d;
  const double* y;
  const double* err;
  const double* ifn;
  const int* pix_idx;
  const double* pix_frac;
  const double* mesh_x;
  const int* mesh_lo;
  const int* pr_aux;
};

// approximate reciprocal / square root (one MUFU op each, ~1 ulp): the float32 spectral sweeps do not need IEEE rounding
#ifndef BAYSAR_CPU_EMU
BAYSAR_DEV float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
BAYSAR_DEV float sqrt_approx(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
BAYSAR_DEV float rsqrt_approx(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
#else
BAYSAR_DEV float rcp_approx(float x) { return 1.0f / x; }
BAYSAR_DEV float sqrt_approx(float x) { return sqrtf(x); }
BAYSAR_DEV float rsqrt_approx(float x) { return 1.0f / sqrtf(x); }
#endif

// Packed FP32 pairs (fma.rn.f32x2 and friends, sm_100+): two FMAs per lane and issue slot
#ifndef BAYSAR_CPU_EMU
struct f32x2 { unsigned long long v; };
BAYSAR_DEV f32x2 pack2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi)); return r; }
BAYSAR_DEV void unpack2(f32x2 a, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v)); }
BAYSAR_DEV f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return r; }
BAYSAR_DEV f32x2 mul2(f32x2 a, f32x2 b) { f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
BAYSAR_DEV f32x2 add2(f32x2 a, f32x2 b) { f32x2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
#else
struct f32x2 { float lo, hi; };
BAYSAR_DEV f32x2 pack2(float lo, float hi) { return f32x2{lo, hi}; }
BAYSAR_DEV void unpack2(f32x2 a, float& lo, float& hi) { lo = a.lo; hi = a.hi; }
BAYSAR_DEV f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { return f32x2{fmaf(a.lo, b.lo, c.lo), fmaf(a.hi, b.hi, c.hi)}; }
BAYSAR_DEV f32x2 mul2(f32x2 a, f32x2 b) { return f32x2{a.lo * b.lo, a.hi * b.hi}; }
BAYSAR_DEV f32x2 add2(f32x2 a, f32x2 b) { return f32x2{a.lo + b.lo, a.hi + b.hi}; }
#endif

// np.clip(v, lo, None): NaN propagates (fmax would swallow it)
BAYSAR_DEV double clip_lo(double v, double lo) { return v < lo ? lo : v; }

// np.nan_to_num: NaN -> 0, +-inf -> +-DBL_MAX
BAYSAR_DEV double nan_to_num(double v) {
  if (v != v) return 0.0;
  if (v > 1.7976931348623157e308) return 1.7976931348623157e308;
  if (v < -1.7976931348623157e308) return -1.7976931348623157e308;
  return v;
}

// max([0, x]) of the reference's gaussian_{high,low}_pass_cost (priors.py:4-11): NaN -> 0
BAYSAR_DEV double pos_part(double x) { return x > 0.0 ? x : 0.0; }
// (the division by the width is a multiplication by its reciprocal: inside a loop over line-of-sight points the
// reciprocal is loop-invariant and leaves the loop; relative difference ~1e-16)
BAYSAR_DEV double high_pass(double v, double thr, double err) {
  double t = pos_part(thr - v) * (1.0 / err);
  return -0.5 * (t * t);
}
BAYSAR_DEV double low_pass(double v, double thr, double err) {
  double t = pos_part(v - thr) * (1.0 / err);
  return -0.5 * (t * t);
}

// numpy.searchsorted(g, x, side="left")
BAYSAR_DEV int lower_bound(const double* g, int n, double x) {
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (g[mid] < x) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// cell + unclamped weight of scipy's linear RegularGridInterpolator (bounds_error=False, fill_value=None)
BAYSAR_DEV void linear_cell(const double* g, int n, double x, int& i, double& a) {
  int k = lower_bound(g, n, x) - 1;
  k = k < 0 ? 0 : (k > n - 2 ? n - 2 : k);
  i = k;
  a = (x - g[k]) / (g[k + 1] - g[k]);
}

BAYSAR_DEV double bilinear(const double* t, int n1, int i, int j, double a, double b) {
  const double* r0 = t + (int64_t)i * n1 + j;
  const double* r1 = r0 + n1;
  return r0[0] * (1 - a) * (1 - b) + r0[1] * (1 - a) * b + r1[0] * a * (1 - b) + r1[1] * a * b;
}

// FITPACK fpbisp/fpbspl for one coordinate, degree 3: clamp to the knot span, locate the interval
// t[l] <= x < t[l+1] and return the 4 non-zero B-spline basis values (scipy RectBivariateSpline.ev -> bispeu).
BAYSAR_DEV void bspline_basis3(const double* t, int n, double x, int& l, double* h) {
  const int k = 3;
  double tb = t[k], te = t[n - k - 1];
  if (x < tb) x = tb;
  if (x > te) x = te;
  // largest l in [k, n - k - 2] with t[l] <= x (FITPACK's forward scan, as a bisection)
  int lo = k, hi = n - k - 2;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (x >= t[mid]) lo = mid; else hi = mid - 1;
  }
  l = lo;
  double hh[3];
  h[0] = 1.0;
  for (int j = 1; j <= k; ++j) {
    for (int i = 0; i < j; ++i) hh[i] = h[i];
    h[0] = 0.0;
    for (int i = 0; i < j; ++i) {
      int li = l + i + 1;
      int lj = li - j;
      double f = hh[i] / (t[li] - t[lj]);
      h[i] = h[i] + f * (t[li] - x);
      h[i + 1] = f * (x - t[lj]);
    }
  }
}

// Two tau slices at once: c holds (slice i, slice i + 1) coefficient pairs.  sum_i hx[i] (sum_j c[i][j] hy[j]) per slice:
// 2 x 20 fused multiply-adds from 16 sixteen-byte loads (FITPACK's fpbisp forms the same sums term by term).
// hx / hy are strided (structure-of-arrays basis cache: element q of LOS point l at q * stride + l).
struct alignas(16) Double2 { double x, y; };  // 16-byte aligned: one LDG.128 per pair (unaligned it was two LDG.64)
BAYSAR_DEV void bspline_eval_pair(const double* c, int ncy, int lx, int ly, const double* hx, const double* hy, int stride,
                                  double& vi, double& vj) {
  const Double2* base = reinterpret_cast<const Double2*>(c) + (int64_t)(lx - 3) * ncy + (ly - 3);
  double si = 0.0, sj = 0.0;
  const double y0 = hy[0], y1 = hy[stride], y2 = hy[2 * stride], y3 = hy[3 * stride];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const Double2 c0 = base[i * ncy], c1 = base[i * ncy + 1], c2 = base[i * ncy + 2], c3 = base[i * ncy + 3];
    const double ri = fma(c3.x, y3, fma(c2.x, y2, fma(c1.x, y1, c0.x * y0)));
    const double rj = fma(c3.y, y3, fma(c2.y, y2, fma(c1.y, y1, c0.y * y0)));
    const double xw = hx[i * stride];
    si = fma(ri, xw, si);
    sj = fma(rj, xw, sj);
  }
  vi = si;
  vj = sj;
}

// ---- warp / block reductions ------------------------------------------------------------------------------
BAYSAR_DEV double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
BAYSAR_DEV double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    double w = __shfl_xor_sync(0xffffffffu, v, o);
    v = w > v ? w : v;
  }
  return v;
}
BAYSAR_DEV float warp_min_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float w = __shfl_xor_sync(0xffffffffu, v, o);
    v = w < v ? w : v;
  }
  return v;
}

// Sum NV doubles over the block.  scratch: NV * (NT/32) doubles.  Result valid in every thread.
template <int NT, int NV>
BAYSAR_DEV void block_sum(double* v, double* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double s = warp_sum(v[k]);
    if (lane == 0) scratch[k * (NT / 32) + warp] = s;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < NT / 32; ++w) s += scratch[k * (NT / 32) + w];
    v[k] = s;
  }
  __syncthreads();
}
