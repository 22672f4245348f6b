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

// Tables derived from the interpolation axes at model creation (los_tables.h): reciprocal spacings, search buckets over
// the leading bits of a double (hi32 >> BAYSAR_KEY_SHIFT: sign, exponent, three mantissa bits).
#define BAYSAR_KEY_SHIFT 17
enum LosAxis { AX_A11_NE, AX_A11_TE, AX_H_NE, AX_H_TE, AX_SPL_X, AX_SPL_Y, AX_COUNT };
struct GridKey {  // bucket table of one sorted axis
  int off;   // offset of the table in LosTablesDev::key_tbl
  int key0;  // key of bucket 0
  int nb;    // number of buckets
  int pad;
};
struct LosTablesDev {
  const double* recip[AX_COUNT];  // linear axes: [n - 1] 1 / (g[k + 1] - g[k]); spline axes: [3][n] 1 / (t[k + j] - t[k]), j = 1..3
  const int* key_tbl;
  GridKey key[AX_COUNT];
};

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
  const double* a11_acd;
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
  const int* gr_i;
  const int* gr_ul;
  LosTablesDev lt;
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

// Programmatic dependent launch (host_plans.h launch_k): pdl_trigger lets the next kernel of the stream be scheduled once
// every CTA of this grid has started; pdl_wait blocks until the previous grid has completed and flushed its writes.
// Every kernel of an evaluation calls both before its first global-memory access, so stream order is preserved.
#ifndef BAYSAR_CPU_EMU
BAYSAR_DEV void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
BAYSAR_DEV void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#else
BAYSAR_DEV void pdl_trigger() {}
BAYSAR_DEV void pdl_wait() {}
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

// ---- searches that start from a bucket table (LosTablesDev) instead of bisecting: one table read from the leading bits
// of x, then a forward scan that almost always ends after zero or one step
#ifndef BAYSAR_CPU_EMU
BAYSAR_DEV int hi_word(double x) { return __double2hiint(x); }
BAYSAR_DEV int lo_word(double x) { return __double2loint(x); }
BAYSAR_DEV double from_words(int hi, int lo) { return __hiloint2double(hi, lo); }
#else
BAYSAR_DEV int hi_word(double x) { int64_t b; std::memcpy(&b, &x, 8); return (int)(b >> 32); }
BAYSAR_DEV int lo_word(double x) { int64_t b; std::memcpy(&b, &x, 8); return (int)(b & 0xffffffffLL); }
BAYSAR_DEV double from_words(int hi, int lo) {
  const int64_t b = (int64_t)(((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo);
  double x; std::memcpy(&x, &b, 8); return x;
}
#endif
BAYSAR_DEV int bucket_start(const LosTablesDev& lt, int ax, double x) {
  const GridKey k = lt.key[ax];
  int b = (hi_word(x) >> BAYSAR_KEY_SHIFT) - k.key0;
  b = b < 0 ? 0 : (b >= k.nb ? k.nb - 1 : b);
  return lt.key_tbl[k.off + b];
}
// numpy.searchsorted(g, x, side="left"); `i` from bucket_start (a lower-bound table): never above the answer
BAYSAR_DEV int lower_bound_from(const double* g, int n, double x, int i) {
  if (x != x) return 0;  // what the bisection returns for NaN
  while (i < n && g[i] < x) ++i;
  return i;
}
// cell + unclamped weight as linear_cell below, the division replaced by the tabulated reciprocal spacing
BAYSAR_DEV void linear_cell_fast(const double* g, int n, double x, const LosTablesDev& lt, int ax, int& i, double& a) {
  int k = lower_bound_from(g, n, x, bucket_start(lt, ax, x)) - 1;
  k = k < 0 ? 0 : (k > n - 2 ? n - 2 : k);
  i = k;
  a = (x - g[k]) * lt.recip[ax][k];
}
// bspline_basis3 below with the interval located from the bucket table (an upper-bound table) and the six divisions of
// the Cox-de Boor recurrence replaced by tabulated reciprocals R_j[k] = 1 / (t[k + j] - t[k])
BAYSAR_DEV void bspline_basis3_fast(const double* t, int n, double x, const LosTablesDev& lt, int ax, int& l, double* h) {
  const double tb = t[3], te = t[n - 4];
  if (x < tb) x = tb;
  if (x > te) x = te;
  int i = x != x ? 0 : bucket_start(lt, ax, x);
  while (i < n && t[i] <= x) ++i;
  l = i - 1 < 3 ? 3 : (i - 1 > n - 5 ? n - 5 : i - 1);
  const double* R = lt.recip[ax];
  double hh[3];
  h[0] = 1.0;
#pragma unroll
  for (int j = 1; j <= 3; ++j) {
#pragma unroll
    for (int q = 0; q < j; ++q) hh[q] = h[q];
    h[0] = 0.0;
#pragma unroll
    for (int q = 0; q < j; ++q) {
      const int li = l + q + 1, lj = li - j;
      const double f = hh[q] * R[(j - 1) * n + lj];
      h[q] = h[q] + f * (t[li] - x);
      h[q + 1] = f * (x - t[lj]);
    }
  }
}

// exp(x) = 2^n e^r, n = rint(x / ln 2), |r| <= ln 2 / 2, Taylor polynomial of degree 12 (truncation 2e-16; ~1 ulp overall;
// the library routine outside |x| < 690 and for NaN).  No table: the LOS stage is bound by the L1 data pipe, and a
// 2^(k/64) table costs a gather of ~4 wavefronts per call.
BAYSAR_DEV double exp_fast(double x) {
  if (!(fabs(x) < 690.0)) return exp(x);
  const double magic = 6755399441055744.0;  // 1.5 * 2^52: the integer lands in the low word
  const double tt = fma(x, 1.4426950408889634, magic);
  const int n = lo_word(tt);
  const double tn = tt - magic;
  double r = fma(tn, -0.6931471805599453, x);
  r = fma(tn, -2.3190468138462996e-17, r);
  double p = 1.0 / 479001600.0;
  p = fma(r, p, 1.0 / 39916800.0);
  p = fma(r, p, 1.0 / 3628800.0);
  p = fma(r, p, 1.0 / 362880.0);
  p = fma(r, p, 1.0 / 40320.0);
  p = fma(r, p, 1.0 / 5040.0);
  p = fma(r, p, 1.0 / 720.0);
  p = fma(r, p, 1.0 / 120.0);
  p = fma(r, p, 1.0 / 24.0);
  p = fma(r, p, 1.0 / 6.0);
  p = fma(r, p, 0.5);
  p = fma(r, p, 1.0);
  p = fma(r, p, 1.0);
  return from_words(hi_word(p) + (n << 20), lo_word(p));
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

// two tables at the same cell with the four corner weights formed once
BAYSAR_DEV void bilinear2(const double* t0, const double* t1, int n1, int i, int j, double a, double b, double& v0, double& v1) {
  const int64_t o = (int64_t)i * n1 + j;
  const double w00 = (1 - a) * (1 - b), w01 = (1 - a) * b, w10 = a * (1 - b), w11 = a * b;
  const double* p = t0 + o;
  v0 = p[0] * w00 + p[1] * w01 + p[n1] * w10 + p[n1 + 1] * w11;
  p = t1 + o;
  v1 = p[0] * w00 + p[1] * w01 + p[n1] * w10 + p[n1 + 1] * w11;
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
BAYSAR_DEV void bspline_eval_pair(const char* cell, size_t row_bytes, const double* hx, const double* hy, int stride,
                                  double& vi, double& vj) {
  // `cell`: address of coefficient pair (lx - 3, ly - 3); the four rows are row_bytes apart (explicit pointers: with
  // an int row index the compiler re-derived a 64-bit address for every one of the 16 loads)
  const Double2* r0 = reinterpret_cast<const Double2*>(cell);
  const Double2* r1 = reinterpret_cast<const Double2*>(cell + row_bytes);
  const Double2* r2 = reinterpret_cast<const Double2*>(cell + 2 * row_bytes);
  const Double2* r3 = reinterpret_cast<const Double2*>(cell + 3 * row_bytes);
  const double y0 = hy[0], y1 = hy[stride], y2 = hy[2 * stride], y3 = hy[3 * stride];
  double si, sj;
  {
    const Double2 c0 = r0[0], c1 = r0[1], c2 = r0[2], c3 = r0[3];
    const double xw = hx[0];
    si = fma(c3.x, y3, fma(c2.x, y2, fma(c1.x, y1, c0.x * y0))) * xw;
    sj = fma(c3.y, y3, fma(c2.y, y2, fma(c1.y, y1, c0.y * y0))) * xw;
  }
  {
    const Double2 c0 = r1[0], c1 = r1[1], c2 = r1[2], c3 = r1[3];
    const double xw = hx[stride];
    si = fma(fma(c3.x, y3, fma(c2.x, y2, fma(c1.x, y1, c0.x * y0))), xw, si);
    sj = fma(fma(c3.y, y3, fma(c2.y, y2, fma(c1.y, y1, c0.y * y0))), xw, sj);
  }
  {
    const Double2 c0 = r2[0], c1 = r2[1], c2 = r2[2], c3 = r2[3];
    const double xw = hx[2 * stride];
    si = fma(fma(c3.x, y3, fma(c2.x, y2, fma(c1.x, y1, c0.x * y0))), xw, si);
    sj = fma(fma(c3.y, y3, fma(c2.y, y2, fma(c1.y, y1, c0.y * y0))), xw, sj);
  }
  {
    const Double2 c0 = r3[0], c1 = r3[1], c2 = r3[2], c3 = r3[3];
    const double xw = hx[3 * stride];
    si = fma(fma(c3.x, y3, fma(c2.x, y2, fma(c1.x, y1, c0.x * y0))), xw, si);
    sj = fma(fma(c3.y, y3, fma(c2.y, y2, fma(c1.y, y1, c0.y * y0))), xw, sj);
  }
  vi = si;
  vj = sj;
}

// One line of an ADAS line group (layout.h SEC_SPL_C): every (ix, iy) entry holds, for each of the group's lines, 32 bytes
// = (tau slice i, slice i + 1) of column iy and of column iy + 1, so a row of four columns is two 256-bit loads `2 entry`
// bytes apart, and lanes that evaluate different lines at the same spline cell read neighbouring bytes of one cache line.
struct alignas(32) Double4 { double a, b, c, d; };
#ifndef BAYSAR_CPU_EMU
BAYSAR_DEV Double4 ld256(const char* p) {
  Double4 r;
  asm volatile("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(r.a), "=d"(r.b), "=d"(r.c), "=d"(r.d) : "l"(p));
  return r;
}
#else
BAYSAR_DEV Double4 ld256(const char* p) { Double4 r; std::memcpy(&r, p, 32); return r; }
#endif
BAYSAR_DEV void bspline_eval_group(const char* cell, size_t entry, size_t row_bytes, const double* hx, const double* hy,
                                   int stride, double& vi, double& vj) {
  const double y0 = hy[0], y1 = hy[stride], y2 = hy[2 * stride], y3 = hy[3 * stride];
  double si = 0.0, sj = 0.0;
#ifndef BAYSAR_LOS_GATHER_AHEAD
#define BAYSAR_LOS_GATHER_AHEAD 4  // rows whose gathers are issued together: 1.771 -> 1.743 ms on the multi-species batch (0 = one row at a time)
#endif
#if BAYSAR_LOS_GATHER_AHEAD == 0
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const char* rp = cell + r * row_bytes;
    const Double4 A = ld256(rp), B = ld256(rp + 2 * entry);
    const double ri = fma(B.c, y3, fma(B.a, y2, fma(A.c, y1, A.a * y0)));
    const double rj = fma(B.d, y3, fma(B.b, y2, fma(A.d, y1, A.b * y0)));
    const double xw = hx[r * stride];
    si = fma(ri, xw, si);
    sj = fma(rj, xw, sj);
  }
#else
  // the gathers of BAYSAR_LOS_GATHER_AHEAD rows are issued before the first of them is used (the loads are volatile asm:
  // their order is the program's), so a trip waits for that many rows' round trips at once instead of one after another
  constexpr int G = BAYSAR_LOS_GATHER_AHEAD;
#pragma unroll
  for (int r0 = 0; r0 < 4; r0 += G) {
    Double4 A[G], B[G];
#pragma unroll
    for (int r = 0; r < G; ++r) {
      const char* rp = cell + (r0 + r) * row_bytes;
      A[r] = ld256(rp);
      B[r] = ld256(rp + 2 * entry);
    }
#pragma unroll
    for (int r = 0; r < G; ++r) {
      const double ri = fma(B[r].c, y3, fma(B[r].a, y2, fma(A[r].c, y1, A[r].a * y0)));
      const double rj = fma(B[r].d, y3, fma(B[r].b, y2, fma(A[r].d, y1, A[r].b * y0)));
      const double xw = hx[(r0 + r) * stride];
      si = fma(ri, xw, si);
      sj = fma(rj, xw, sj);
    }
  }
#endif
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
