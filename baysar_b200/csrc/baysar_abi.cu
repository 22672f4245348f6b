// C ABI of the B200-native BaySAR posterior path (include/baysar_b200.h): model upload, workspace, launches.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <new>
#include <string>
#include <vector>

#include "../../include/baysar_b200.h"
#include "chord_stage.cuh"
#include "ensemble.cuh"
#include "finalize.cuh"
#include "if_contract.cuh"
#include "los_stage.cuh"
#include "mma_probe.cuh"
#include "pix_contract.cuh"

static thread_local std::string g_last_error;

static int fail(const std::string& msg) {
  g_last_error = msg;
  return 1;
}
#define CUDA_OK(expr)                                                                          \
  do {                                                                                         \
    cudaError_t e_ = (expr);                                                                   \
    if (e_ != cudaSuccess) return fail(std::string(#expr) + ": " + cudaGetErrorString(e_));    \
  } while (0)

namespace {
// Host-side packing of one chord's instrument function for the tensor-core contraction (if_contract.cuh).
struct IfcPlan {
  int m_lo = 0, m_hi = 0, pad_left = 0, n_tiles = 0, taps_len = 0;
  int64_t pitch = 0, out_pitch = 0;
  std::vector<float> taps_rev;  // [2][4][taps_len]
  // compact output for a static wavelength map: needed-column bit mask / running count per 32 columns
  std::vector<uint32_t> cmask;
  std::vector<int> cbase;
  int64_t c_pitch = 0;
};

// Host-side plan of the folded pixel contraction (pix_contract.cuh) of one chord: column tiles with their bands, the
// zero-padded tap table, the end-zone weights of the area identity, and per-batch-size schedules.
struct PxcSchedule {
  int ctas = 0, slots = 0;
  int *d_sched = nullptr, *d_cnt = nullptr;
};
struct PxcPlan {
  bool on = false;
  int tn = 128, pad_left = 0, n_col_tiles = 0, wlen = 0, ring = 0;
  int e_l = 0, e_r = 0, x_l = 0, x_r = 0, col_left0 = 0, col_right0 = 0, n_dl = 0, n_dr = 0;
  int64_t pitch = 0, out_pitch = 0;
  double wsum = 0.0;
  std::vector<int> tile_kcol0, tile_nkb, tile_boff, col_idx, pix_col;
  int n_fold_tiles = 0;
  std::vector<float> inv_err;  // 1 / err of the chord's pixels (pixel_fold_kernel's log-posterior loop)
  float* d_inv_err = nullptr;
  std::vector<float> col_frac, wtab, bbar;  // bbar: per folded tile, sum over its columns of B[s, c] (bias correction)
  std::vector<double> dtab;
  int *d_tile_kcol0 = nullptr, *d_tile_nkb = nullptr, *d_tile_boff = nullptr, *d_col_idx = nullptr, *d_pix_col = nullptr;
  float *d_col_frac = nullptr, *d_wtab = nullptr, *d_bbar = nullptr;
  double* d_dtab = nullptr;
  std::map<int, PxcSchedule> schedules;  // by number of theta blocks
};

static float tf32_round(float v) {  // round to nearest, ties away from zero, 10-bit mantissa (cvt.rna.tf32.f32)
  uint32_t u;
  memcpy(&u, &v, 4);
  u = (u + 0x1000u) & 0xFFFFE000u;
  float r;
  memcpy(&r, &u, 4);
  return r;
}

static inline int floor_div(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

}  // namespace

struct baysar_model {
  int device = 0;
  unsigned char* blob_dev = nullptr;
  size_t blob_bytes = 0;
  std::vector<unsigned char> blob_host;  // header + small tables are read on the host for launch configuration
  ModelDev dev;
  const int64_t* I = nullptr;   // host view
  const int* ch_i = nullptr;    // host view
  const int* cl_i = nullptr;
  // workspace (grows on demand)
  int64_t cap = 0;
  double* lines = nullptr;
  double* comps = nullptr;
  double* hdr = nullptr;
  uint32_t* status = nullptr;
  double* theta_stage = nullptr;  // for the *_host entry point
  double* logp_stage = nullptr;     // [n] float64 logP, then [n] uint32 status
  void* out_pinned = nullptr;       // pinned host mirror of logp_stage
  int64_t stage_cap = 0;
  size_t chord_smem_max = 0;            // fused launches covering several chords
  std::vector<size_t> chord_smem;       // per chord (the tensor path's spectral stage may drop buffer B)
  // tensor-core path (if_contract.cuh): per chord plan (empty taps_rev = chord stays on the fused kernel)
  std::vector<IfcPlan> ifc;
  std::vector<float*> ifc_taps_dev;
  std::vector<uint32_t*> ifc_cmask_dev;
  std::vector<int*> ifc_cbase_dev;
  float* conv_c = nullptr;
  int64_t conv_c_pitch_max = 0;
  std::vector<PxcPlan> pxc;  // folded pixel contraction (static wavelength map, non-negative instrument function)
  bool any_ifc = false;
  int64_t ifc_rows = 0;              // rows of the split / conv workspaces (multiple of 128)
  int64_t ifc_budget_bytes = (int64_t)24 << 30;  // HBM budget of those workspaces (BAYSAR_IFC_BUDGET_MB at create)
  float *spec = nullptr, *conv = nullptr;  // zero-padded FP32 spectra (A operand), convolved spectra
  int sm_count = 0;
  double *trapz_part = nullptr, *aux = nullptr;
  int64_t split_pitch_max = 0, conv_pitch_max = 0;
  int tiles_max = 0;
  int last_launches = 3;
  // optional per-stage timing (bench.py roofline): events tagged with the stage that ends at them
  bool profiling = false;
  std::vector<cudaEvent_t> prof_events;
  std::vector<int> prof_slots;
};

namespace {

struct SectionTable {
  const int64_t* hdr;
  const unsigned char* base;
  const void* ptr(int s) const { return base + hdr[3 + 2 * s]; }
  int64_t count(int s) const { return hdr[3 + 2 * s + 1]; }
};

// Peak-rate probes for the roofline denominators MEASURED_PEAKS.json does not carry (FP32 FMA, MUFU, FP64 FMA):
// 16 independent dependency chains per thread, enough resident warps to saturate the pipe.
__global__ void microbench_ffma_kernel(float* out, int iters, float a) {
  float v[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) v[k] = threadIdx.x * 1e-3f + k;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 16; ++k) v[k] = fmaf(v[k], a, 0.5f);
  }
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 16; ++k) s += v[k];
  if (s == 12345.678f) out[threadIdx.x] = s;
}
__global__ void microbench_mufu_kernel(float* out, int iters, float a) {
  float v[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) v[k] = threadIdx.x * 1e-3f + k + 1.f;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 16; ++k) asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(v[k]) : "f"(v[k]));
  }
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 16; ++k) s += v[k];
  if (s == 12345.678f) out[threadIdx.x] = s * a;
}
// FFMA with all three operands in registers (the form real kernels issue; the immediate form above is the pipe's
// best case) and the packed FFMA2 (fma.rn.f32x2, two FMAs per lane and instruction, sm_100+)
__global__ void microbench_ffma_reg_kernel(float* out, int iters, float a, float b) {
  float v[16];
  const float ar = a + (float)(threadIdx.x & 1) * 1e-9f, br = b + (float)(threadIdx.x & 2) * 1e-9f;
#pragma unroll
  for (int k = 0; k < 16; ++k) v[k] = threadIdx.x * 1e-3f + k;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 16; ++k) v[k] = fmaf(v[k], ar, br);
  }
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 16; ++k) s += v[k];
  if (s == 12345.678f) out[threadIdx.x] = s;
}
__global__ void microbench_ffma2_kernel(float* out, int iters, float a, float b) {
  unsigned long long v[16], ar, br;
  const float a0 = a + (float)(threadIdx.x & 1) * 1e-9f, b0 = b + (float)(threadIdx.x & 2) * 1e-9f;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ar) : "f"(a0), "f"(a0));
  asm("mov.b64 %0, {%1, %2};" : "=l"(br) : "f"(b0), "f"(b0));
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const float x = threadIdx.x * 1e-3f + k;
    asm("mov.b64 %0, {%1, %2};" : "=l"(v[k]) : "f"(x), "f"(x + 0.5f));
  }
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 16; ++k) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v[k]) : "l"(ar), "l"(br));
  }
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v[k]));
    s += lo + hi;
  }
  if (s == 12345.678f) out[threadIdx.x] = s;
}
__global__ void microbench_dfma_kernel(double* out, int iters, double a) {
  double v[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) v[k] = threadIdx.x * 1e-3 + k;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 16; ++k) v[k] = fma(v[k], a, 0.5);
  }
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < 16; ++k) s += v[k];
  if (s == 12345.678) out[threadIdx.x] = s;
}

IfcPlan make_ifc_plan(const double* w, int KI, int k_lo, int k_hi, int F) {
  IfcPlan pl;
  const int o = (KI - 1) / 2;
  pl.m_lo = floor_div(o - k_hi + 1, ifc::TILE_K);
  pl.m_hi = floor_div(ifc::TILE_N - 1 + o - k_lo, ifc::TILE_K);
  pl.pad_left = -ifc::TILE_K * pl.m_lo;
  if (pl.pad_left < 0) pl.pad_left = 0;
  pl.n_tiles = (F + ifc::TILE_N - 1) / ifc::TILE_N;
  int64_t last_col = (int64_t)pl.pad_left + (int64_t)(pl.n_tiles - 1) * ifc::TILE_N + ifc::TILE_K * pl.m_hi + ifc::TILE_K;
  if (last_col < pl.pad_left + F) last_col = pl.pad_left + F;
  pl.pitch = (last_col + 31) / 32 * 32;
  pl.out_pitch = (int64_t)pl.n_tiles * ifc::TILE_N;
  pl.taps_len = (164 + ifc::TILE_K * (pl.m_hi - pl.m_lo) + 3) / 4 * 4;
  const int x_max = ifc::TILE_N - 1 + o - ifc::TILE_K * pl.m_lo;
  pl.taps_rev.assign((size_t)8 * pl.taps_len, 0.f);
  for (int s = 0; s < 4; ++s) {
    for (int i = 0; i < pl.taps_len; ++i) {
      const int x = x_max - (i + s);  // copy_s[i] = rev[i + s] = w[x_max - i - s]
      float v = (x >= k_lo && x < k_hi) ? (float)w[x] : 0.f;
      float hi = tf32_round(v), lo = tf32_round(v - hi);
      pl.taps_rev[(size_t)s * pl.taps_len + i] = hi;
      pl.taps_rev[(size_t)(4 + s) * pl.taps_len + i] = lo;
    }
  }
  return pl;
}

// Columns of the folded contraction: one per pixel (i_p, a_p), then the explicit fine-grid columns of the two edge
// zones; every group padded to whole tiles by repeating its last column.
PxcPlan make_pxc_plan(const double* w, int KI, int k_lo, int k_hi, int F, int P, const int* pix_idx, const double* pix_frac,
                      int tn) {
  PxcPlan pl;
  pl.tn = tn;
  const int o = (KI - 1) / 2;
  for (int k = k_lo; k < k_hi; ++k)
    if (!(w[k] >= 0.0)) return pl;  // the clip argument needs non-negative taps
  pl.e_l = std::max(0, k_hi - 1 - o);
  pl.e_r = std::max(0, o - k_lo);
  pl.x_l = pl.e_l + 1;
  pl.x_r = pl.e_r + 1;
  if (F < 4 * (pl.e_l + pl.e_r + 2) + 2 * tn || P < 1) return pl;
  std::vector<int> ci;
  std::vector<double> ca;
  auto pad_group = [&]() { while (ci.size() % tn) { ci.push_back(ci.back()); ca.push_back(ca.back()); } };
  pl.pix_col.assign(P, -1);
  for (int p = 0; p < P; ++p) {
    const int i = std::min(std::max(pix_idx[p], 0), F - 2);
    const bool edge = pix_idx[p] != i || i < pl.e_l || i + 1 >= F - pl.e_r;
    if (!edge) pl.pix_col[p] = (int)ci.size();
    ci.push_back(i);
    ca.push_back(edge ? 0.0 : pix_frac[p]);
  }
  pad_group();
  pl.col_left0 = (int)ci.size();
  for (int j = 0; j < pl.x_l; ++j) { ci.push_back(j); ca.push_back(0.0); }
  pad_group();
  pl.col_right0 = (int)ci.size();
  for (int j = F - pl.x_r; j < F; ++j) { ci.push_back(j); ca.push_back(0.0); }
  pad_group();
  pl.n_col_tiles = (int)ci.size() / tn;
  pl.out_pitch = (int64_t)ci.size();
  std::vector<int> k0(pl.n_col_tiles);
  pl.tile_nkb.assign(pl.n_col_tiles, 0);
  int min_k0 = 0, max_end = F;
  for (int t = 0; t < pl.n_col_tiles; ++t) {
    int i_min = F, i_max = 0;
    for (int c = t * tn; c < (t + 1) * tn; ++c) {
      i_min = std::min(i_min, ci[c]);
      i_max = std::max(i_max, ci[c] + (ca[c] != 0.0 ? 1 : 0));
    }
    const int s_min = i_min + o - (k_hi - 1), s_max = i_max + o - k_lo;
    k0[t] = floor_div(s_min, ifc::TILE_K) * ifc::TILE_K;
    pl.tile_nkb[t] = std::max((s_max - k0[t]) / ifc::TILE_K + 1, pxc::MIN_KB);
    min_k0 = std::min(min_k0, k0[t]);
    max_end = std::max(max_end, k0[t] + ifc::TILE_K * pl.tile_nkb[t]);
  }
  pl.pad_left = -min_k0;  // multiple of 16
  pl.pitch = ((int64_t)pl.pad_left + max_end + 31) / 32 * 32;
  pl.tile_kcol0.resize(pl.n_col_tiles);
  // tap index of (column c, source s): u = i_c + o - s; the table holds u in [u_base, u_base + wlen)
  int u_base = 1 << 30, u_top = -(1 << 30);
  for (int t = 0; t < pl.n_col_tiles; ++t) {
    pl.tile_kcol0[t] = pl.pad_left + k0[t];
    for (int c = t * tn; c < (t + 1) * tn; ++c) {
      u_base = std::min(u_base, ci[c] + o - (k0[t] + ifc::TILE_K * pl.tile_nkb[t] - 1));
      u_top = std::max(u_top, ci[c] + 1 + o - k0[t]);
    }
  }
  u_base -= 4;  // the builders read five consecutive taps from x - 3
  pl.wlen = (u_top - u_base + 1 + 4 + 3) / 4 * 4;
  pl.wtab.assign(pl.wlen, 0.f);
  for (int x = 0; x < pl.wlen; ++x) {
    const int u = x + u_base;
    if (u >= k_lo && u < k_hi) pl.wtab[x] = (float)w[u];
  }
  pl.col_idx.resize(ci.size());
  pl.col_frac.resize(ci.size());
  for (int t = 0; t < pl.n_col_tiles; ++t)
    for (int c = t * tn; c < (t + 1) * tn; ++c) {
      pl.col_idx[c] = ci[c] + o - k0[t] - u_base;
      pl.col_frac[c] = (float)ca[c];
    }
  // area identity: G_s - W c_s at the two ends (exactly zero in between: same summation order as W)
  double wsum = 0.0;
  for (int k = k_lo; k < k_hi; ++k) wsum += w[k];
  pl.wsum = wsum;
  const int zone = std::min(F / 2, pl.e_l + pl.e_r + 4);
  auto d_of = [&](int s) {
    double g = 0.0;
    for (int k = k_lo; k < k_hi; ++k) {
      const int j = s - o + k;
      if (j >= 0 && j < F) g += ((j == 0 || j == F - 1) ? 0.5 : 1.0) * w[k];
    }
    return g - wsum * ((s == 0 || s == F - 1) ? 0.5 : 1.0);
  };
  std::vector<double> dl(zone), dr(zone);
  for (int s = 0; s < zone; ++s) { dl[s] = d_of(s); dr[s] = d_of(F - zone + s); }
  pl.n_dl = zone;
  while (pl.n_dl > 0 && dl[pl.n_dl - 1] == 0.0) --pl.n_dl;
  int first = 0;
  while (first < zone && dr[first] == 0.0) ++first;
  pl.n_dr = zone - first;
  pl.dtab.assign(dl.begin(), dl.begin() + pl.n_dl);
  pl.dtab.insert(pl.dtab.end(), dr.begin() + first, dr.end());
  // column-sum identity per folded tile (bias correction of the truncating TMEM accumulation, pixel_fold_kernel):
  //   sum_{c in tile} C[c] = sum_s S_s bbar_s,  bbar_s = sum_{c in tile} B[s, c] over the tile's band (padding
  //   columns repeat the last pixel and are part of both sides)
  pl.n_fold_tiles = pl.col_left0 / tn;
  pl.tile_boff.assign(pl.n_col_tiles, -1);
  if (pl.n_fold_tiles <= 64) {
    for (int t = 0; t < pl.n_fold_tiles; ++t) {
      pl.tile_boff[t] = (int)pl.bbar.size();
      const int len = ifc::TILE_K * pl.tile_nkb[t];
      std::vector<double> bb(len, 0.0);
      for (int c = t * tn; c < (t + 1) * tn; ++c)
        for (int x = 0; x < len; ++x) {
          const int idx = pl.col_idx[c] - x;  // the builders' index arithmetic
          bb[x] += (double)pl.wtab[idx] + (double)pl.col_frac[c] * ((double)pl.wtab[idx + 1] - (double)pl.wtab[idx]);
        }
      for (int x = 0; x < len; ++x) pl.bbar.push_back((float)bb[x]);
    }
  }
  if (pl.bbar.empty()) pl.bbar.push_back(0.f);
  if (pl.dtab.empty()) pl.dtab.push_back(0.0);
  pl.ring = tn == 64 ? pxc::ring_stages<64>(pl.wlen) : pxc::ring_stages<128>(pl.wlen);
  pl.on = pl.ring >= 2 && pl.n_col_tiles < 32768;
  return pl;
}

template <class T>
int upload(T** dst, const std::vector<T>& v) {
  CUDA_OK(cudaMalloc(dst, sizeof(T) * v.size()));
  CUDA_OK(cudaMemcpy(*dst, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice));
  return 0;
}
int upload_pxc_plan(PxcPlan& pl) {
  if (upload(&pl.d_tile_kcol0, pl.tile_kcol0) || upload(&pl.d_tile_nkb, pl.tile_nkb) || upload(&pl.d_col_idx, pl.col_idx) ||
      upload(&pl.d_tile_boff, pl.tile_boff) ||
      upload(&pl.d_pix_col, pl.pix_col) || upload(&pl.d_col_frac, pl.col_frac) || upload(&pl.d_wtab, pl.wtab) ||
      upload(&pl.d_dtab, pl.dtab) || upload(&pl.d_bbar, pl.bbar))
    return 1;
  if (!pl.inv_err.empty() && upload(&pl.d_inv_err, pl.inv_err)) return 1;
  return 0;
}
void free_pxc_plan(PxcPlan& pl) {
  cudaFree(pl.d_tile_kcol0); cudaFree(pl.d_tile_nkb); cudaFree(pl.d_tile_boff); cudaFree(pl.d_col_idx); cudaFree(pl.d_pix_col);
  cudaFree(pl.d_col_frac); cudaFree(pl.d_wtab); cudaFree(pl.d_dtab); cudaFree(pl.d_bbar); cudaFree(pl.d_inv_err);
  for (auto& kv : pl.schedules) { cudaFree(kv.second.d_sched); cudaFree(kv.second.d_cnt); }
  pl.schedules.clear();
}

// Longest-first deal of the (column tile, theta block) pairs to `sms` persistent CTAs (cost = k-blocks + a constant
// for the epilogue); a CTA runs its tiles in the order dealt.  Cached per number of theta blocks.
int pxc_schedule(PxcPlan& pl, int n_rb, int sms, const PxcSchedule** out) {
  auto it = pl.schedules.find(n_rb);
  if (it == pl.schedules.end()) {
    std::vector<int> order(pl.n_col_tiles);
    for (int t = 0; t < pl.n_col_tiles; ++t) order[t] = t;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return pl.tile_nkb[a] > pl.tile_nkb[b]; });
    const int64_t tiles = (int64_t)n_rb * pl.n_col_tiles;
    PxcSchedule sc;
    sc.ctas = (int)std::min<int64_t>(tiles, sms);
    std::vector<std::vector<int>> lists(sc.ctas);
    std::vector<int64_t> load(sc.ctas, 0);
    for (int t : order)
      for (int rb = 0; rb < n_rb; ++rb) {
        int best = 0;
        for (int c = 1; c < sc.ctas; ++c)
          if (load[c] < load[best]) best = c;
        lists[best].push_back((t << 16) | rb);
        load[best] += pl.tile_nkb[t] + 10;
      }
    for (auto& l : lists) sc.slots = std::max(sc.slots, (int)l.size());
    std::vector<int> flat((size_t)sc.ctas * sc.slots, 0), cnt(sc.ctas);
    for (int c = 0; c < sc.ctas; ++c) {
      cnt[c] = (int)lists[c].size();
      std::copy(lists[c].begin(), lists[c].end(), flat.begin() + (size_t)c * sc.slots);
    }
    if (upload(&sc.d_sched, flat) || upload(&sc.d_cnt, cnt)) return 1;
    it = pl.schedules.emplace(n_rb, sc).first;
  }
  *out = &it->second;
  return 0;
}

template <int TN, bool TRACE>
int launch_pxc_t(const pxc::Params& p, const CUtensorMap& tmap, int ctas, size_t smem, cudaStream_t stream) {
  CUDA_OK(cudaFuncSetAttribute(pxc::pix_contract_kernel<TN, TRACE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  pxc::pix_contract_kernel<TN, TRACE><<<ctas, pxc::THREADS, smem, stream>>>(p, tmap);
  CUDA_OK(cudaGetLastError());
  return 0;
}

// Tensor map of the zero-padded spectra for the contraction's A loads: float32 {pitch, rows}, box {16, 128}, SWIZZLE_64B
int make_spec_tmap(const float* spec, int64_t pitch, int64_t rows, CUtensorMap* out) {
  typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static encode_fn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) return fail("cuTensorMapEncodeTiled not available in this driver");
    encode = reinterpret_cast<encode_fn>(fn);
  }
  const cuuint64_t dims[2] = {(cuuint64_t)pitch, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)pitch * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)ifc::TILE_K, (cuuint32_t)ifc::TILE_M};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(spec), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
  return 0;
}

int launch_pxc(PxcPlan& pl, const float* spec, int n_rb, float* out, double* tsum, int sms, cudaStream_t stream,
               long long* dbg_cycles = nullptr) {
  if (n_rb > 65535) return fail("too many theta blocks for one folded contraction launch");
  const PxcSchedule* sc = nullptr;
  if (pxc_schedule(pl, n_rb, sms, &sc)) return 1;
  pxc::Params p{};
  p.s = spec; p.pitch = pl.pitch; p.tile_kcol0 = pl.d_tile_kcol0; p.tile_nkb = pl.d_tile_nkb; p.col_idx = pl.d_col_idx;
  p.col_frac = pl.d_col_frac; p.wtab = pl.d_wtab; p.wlen = pl.wlen; p.sched = sc->d_sched; p.sched_cnt = sc->d_cnt;
  p.sched_slots = sc->slots; p.out = out; p.out_pitch = pl.out_pitch; p.ring = pl.ring; p.dbg_cycles = dbg_cycles;
  p.bbar = (tsum && pl.n_fold_tiles <= 64) ? pl.d_bbar : nullptr; p.tile_boff = pl.d_tile_boff; p.tsum = tsum;
  p.n_col_tiles = pl.n_col_tiles;
  // A loader: 0 = cp.async by two loader warps (default), 1 = one TMA box per k-block (opt-in: 2 % faster on the Balmer
  // demo, but one build of it was seen to deliver stale A rows in cold first launches; not root-caused)
  static const int loader = getenv("BAYSAR_PXC_LOADER") ? atoi(getenv("BAYSAR_PXC_LOADER")) : 0;
  p.use_tma = loader;
  static const int sleep_epi = getenv("BAYSAR_PXC_SLEEP_EPI") ? atoi(getenv("BAYSAR_PXC_SLEEP_EPI")) : 0;
  static const int sleep_build = getenv("BAYSAR_PXC_SLEEP_BUILD") ? atoi(getenv("BAYSAR_PXC_SLEEP_BUILD")) : 0;
  p.sleep_epi = (unsigned)sleep_epi; p.sleep_build = (unsigned)sleep_build;
  CUtensorMap tmap;
  if (make_spec_tmap(spec, pl.pitch, (int64_t)n_rb * ifc::TILE_M, &tmap)) return 1;
  if (pl.tn == 64) {
    const size_t smem = pxc::smem_bytes<64>(pl.ring, pl.wlen);
    return dbg_cycles ? launch_pxc_t<64, true>(p, tmap, sc->ctas, smem, stream)
                      : launch_pxc_t<64, false>(p, tmap, sc->ctas, smem, stream);
  }
  const size_t smem = pxc::smem_bytes<128>(pl.ring, pl.wlen);
  return dbg_cycles ? launch_pxc_t<128, true>(p, tmap, sc->ctas, smem, stream)
                    : launch_pxc_t<128, false>(p, tmap, sc->ctas, smem, stream);
}

int ensure_workspace(baysar_model* m, int64_t n) {
  if (n <= m->cap) return 0;
  int64_t cap = n < 1024 ? 1024 : n;
  const int64_t n_ul = m->I[I_N_ULINES] > 0 ? m->I[I_N_ULINES] : 1;
  const int64_t ncols = m->I[I_N_CHORDS] + m->I[I_N_PRIORS];
  cudaFree(m->lines); cudaFree(m->comps); cudaFree(m->hdr); cudaFree(m->status);
  m->lines = nullptr; m->comps = nullptr; m->hdr = nullptr; m->status = nullptr; m->cap = 0;
  CUDA_OK(cudaMalloc(&m->lines, sizeof(double) * cap * n_ul * BAYSAR_LINE_STRIDE));
  CUDA_OK(cudaMalloc(&m->comps, sizeof(double) * cap * ncols));
  CUDA_OK(cudaMalloc(&m->hdr, sizeof(double) * cap * BAYSAR_HDR_STRIDE));
  CUDA_OK(cudaMalloc(&m->status, sizeof(uint32_t) * cap));
  m->cap = cap;
  return 0;
}

int launch_los(baysar_model* m, const double* theta, int64_t n, double* lines, double* comps, double* profiles,
               uint32_t* status, cudaStream_t stream) {
  LosArgs a;
  a.theta = theta; a.n = n; a.lines = lines; a.components = comps; a.hdr = m->hdr; a.profiles = profiles;
  a.status = status;
  const int L = (int)m->I[I_L];
  const size_t per_warp = los_smem_per_warp(L, (int)m->I[I_N_PARAMS], (int)m->I[I_N_SPECIES]);
  if (per_warp > 200 * 1024) return fail("line-of-sight grid too long for the shared-memory LOS stage");
  // Register budget by batch size.  The kernel is latency-bound at 16 warps per SM (4 CTAs, 128 registers, no spills):
  // the best choice for many waves.  A batch of a few thousand thetas, though, is one or two waves, and then fitting
  // every CTA into ONE wave beats the spills of a tighter budget (4096 thetas: 1024 CTAs = 1.7 waves of 592 at 0.080 ms,
  // one wave of 7 CTAs per SM at 72 registers at 0.068 ms).  BAYSAR_LOS_MINB overrides.
  static const int minb_env = getenv("BAYSAR_LOS_MINB") ? atoi(getenv("BAYSAR_LOS_MINB")) : 0;
  int minb = 4;
  if (minb_env > 0) {
    minb = minb_env;
  } else {
    const int64_t ctas = (n + 3) / 4;
    if (ctas > (int64_t)4 * m->sm_count)
      for (int b = 5; b <= 7; ++b)
        if (ctas <= (int64_t)b * m->sm_count) { minb = b; break; }
  }
  // Long lines of sight: the per-point cache (128 B per point and theta) starves the warp-per-theta form of occupancy
  // (L = 1000: one warp per SM).  From BAYSAR_LOS_TEAM_L points on (default: as soon as four thetas no longer fit one
  // CTA's 200 KB) one CTA of four warps works on ONE theta, the points spread over its 128 threads.
  static const int team_l = getenv("BAYSAR_LOS_TEAM_L") ? atoi(getenv("BAYSAR_LOS_TEAM_L")) : 0;
  const bool team = (team_l > 0 && L >= team_l) || 4 * per_warp > 200 * 1024;
  if (team && per_warp + 128 <= 200 * 1024) {
    const size_t smem = per_warp + 128;
    if (2 * smem > 200 * 1024) {  // one CTA per SM anyway: eight warps per theta
      auto kern = los_stage_kernel<8, 1, 1, 8>;
      CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      kern<<<(unsigned)n, 256, smem, stream>>>(m->dev, a);
    } else {
      auto kern = los_stage_kernel<4, 1, 1, 4>;
      CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      kern<<<(unsigned)n, 128, smem, stream>>>(m->dev, a);
    }
  } else if (4 * per_warp <= 200 * 1024) {
    size_t smem = 4 * per_warp;
    static const int unr = getenv("BAYSAR_LOS_UNROLL") ? atoi(getenv("BAYSAR_LOS_UNROLL")) : 1;
    auto kern = minb >= 6 ? los_stage_kernel<4, 6> : (minb == 5 ? los_stage_kernel<4, 5> : los_stage_kernel<4, 4>);
    if (minb == 7) kern = los_stage_kernel<4, 7>;
    if (minb >= 8) kern = los_stage_kernel<4, 8>;
    if (unr == 2) kern = minb >= 5 ? los_stage_kernel<4, 5, 2> : (minb == 3 ? los_stage_kernel<4, 3, 2> : los_stage_kernel<4, 4, 2>);
    CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)((n + 3) / 4), 128, smem, stream>>>(m->dev, a);
  } else {
    size_t smem = per_warp;
    CUDA_OK(cudaFuncSetAttribute(los_stage_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    los_stage_kernel<1><<<(unsigned)n, 32, smem, stream>>>(m->dev, a);
  }
  CUDA_OK(cudaGetLastError());
  return 0;
}

int ensure_ifc_workspace(baysar_model* m, int64_t n) {
  // rows handled per pass of the tensor-core path: bounded so that the split / convolved spectra stay within a
  // fixed HBM budget (8 bytes of operands + 4 bytes of output per fine-grid point and row)
  const int64_t budget = m->ifc_budget_bytes;
  const int64_t per_row = 4 * (m->split_pitch_max + m->conv_pitch_max + m->conv_c_pitch_max) + 8 * m->tiles_max;
  int64_t rows = (n + ifc::TILE_M - 1) / ifc::TILE_M * ifc::TILE_M;
  int64_t max_rows = budget / per_row / ifc::TILE_M * ifc::TILE_M;
  if (max_rows < ifc::TILE_M) max_rows = ifc::TILE_M;
  if (rows > max_rows) rows = max_rows;
  if (rows <= m->ifc_rows) return 0;
  cudaFree(m->spec); cudaFree(m->conv); cudaFree(m->conv_c); cudaFree(m->trapz_part); cudaFree(m->aux);
  m->spec = m->conv = m->conv_c = nullptr; m->trapz_part = m->aux = nullptr; m->ifc_rows = 0;
  CUDA_OK(cudaMalloc(&m->spec, sizeof(float) * rows * m->split_pitch_max));
  CUDA_OK(cudaMalloc(&m->conv, sizeof(float) * rows * m->conv_pitch_max));
  if (m->conv_c_pitch_max > 0) CUDA_OK(cudaMalloc(&m->conv_c, sizeof(float) * rows * m->conv_c_pitch_max));
  CUDA_OK(cudaMalloc(&m->trapz_part, sizeof(double) * rows * m->tiles_max));
  CUDA_OK(cudaMalloc(&m->aux, sizeof(double) * rows * BAYSAR_AUX_STRIDE));
  // rows past the end of a batch are multiplied too (M = 128 granularity): keep them finite
  CUDA_OK(cudaMemset(m->spec, 0, sizeof(float) * rows * m->split_pitch_max));
  m->ifc_rows = rows;
  return 0;
}

// Launch the chord kernel with the register budget that matches how many CTAs the shared-memory footprint lets an SM hold
template <int MODE>
int launch_chord_kernel(baysar_model* m, const ChordArgs& a, dim3 grid, size_t smem, cudaStream_t stream, bool single_tap_only = false) {
  constexpr int NT = 256;
  static const int cap = getenv("BAYSAR_CHORD_MINB") ? atoi(getenv("BAYSAR_CHORD_MINB")) : 5;
  const int by_smem = (int)((size_t)232448 / (smem + 1024));
  const int minb = by_smem < cap ? by_smem : cap;
  auto kern = minb >= 5 ? chord_stage_kernel<NT, MODE, 5> : (minb == 4 ? chord_stage_kernel<NT, MODE, 4> : chord_stage_kernel<NT, MODE, 3>);
  if (single_tap_only)  // every Balmer line of the chord is a guaranteed single Doppler tap: leaner build
    kern = minb >= 5 ? chord_stage_kernel<NT, MODE, 5, false>
                     : (minb == 4 ? chord_stage_kernel<NT, MODE, 4, false> : chord_stage_kernel<NT, MODE, 3, false>);
  CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, NT, smem, stream>>>(m->dev, a);
  CUDA_OK(cudaGetLastError());
  return 0;
}

// stage tags for the optional event log
enum { ST_BEGIN = -1, ST_LOS = 0, ST_CHORD = 1, ST_CONTRACT = 2, ST_PIXEL = 3, ST_FINAL = 4, ST_COUNT = 5 };

int mark(baysar_model* m, int slot, cudaStream_t stream) {
  if (!m->profiling) return 0;
  cudaEvent_t ev;
  CUDA_OK(cudaEventCreate(&ev));
  CUDA_OK(cudaEventRecord(ev, stream));
  m->prof_events.push_back(ev);
  m->prof_slots.push_back(slot);
  return 0;
}

int launch_chords(baysar_model* m, const double* theta, int64_t n, int chord_begin, int n_chords,
                  const double* lines, double* comps, double* spectra, float* fine_pre, float* fine_post,
                  uint32_t* status, cudaStream_t stream, double* logp = nullptr, bool* finalized = nullptr) {
  constexpr int NT = 256;
  const int n_params = (int)m->I[I_N_PARAMS];
  const int ncols = (int)(m->I[I_N_CHORDS] + m->I[I_N_PRIORS]);
  bool any = false;
  for (int c = chord_begin; c < chord_begin + n_chords; ++c) any = any || !m->ifc[c].taps_rev.empty();
  if (!any) {
    ChordArgs a{};
    a.theta = theta; a.n = n; a.chord_begin = chord_begin; a.lines = lines; a.components = comps;
    a.spectra = spectra; a.fine_pre = fine_pre; a.fine_post = fine_post; a.status = status;
    dim3 grid((unsigned)n, (unsigned)n_chords);
    if (launch_chord_kernel<0>(m, a, grid, m->chord_smem_max, stream)) return 1;
    m->last_launches += 1;
    if (mark(m, ST_CHORD, stream)) return 1;
    return 0;
  }
  if (ensure_ifc_workspace(m, n)) return 1;
  for (int c = chord_begin; c < chord_begin + n_chords; ++c) {
    const IfcPlan& pl = m->ifc[c];
    const int P = m->ch_i[c * CH_I_COUNT + CH_P], F = m->ch_i[c * CH_I_COUNT + CH_F];
    if (pl.taps_rev.empty()) {
      ChordArgs a{};
      a.theta = theta; a.n = n; a.chord_begin = c; a.lines = lines; a.components = comps;
      a.spectra = spectra; a.fine_pre = fine_pre; a.fine_post = fine_post; a.status = status;
      if (launch_chord_kernel<0>(m, a, dim3((unsigned)n, 1), m->chord_smem[c], stream)) return 1;
      m->last_launches += 1;
      if (mark(m, ST_CHORD, stream)) return 1;
      continue;
    }
    for (int64_t r0 = 0; r0 < n; r0 += m->ifc_rows) {
      const int64_t rows = (n - r0) < m->ifc_rows ? (n - r0) : m->ifc_rows;
      const int64_t rows128 = (rows + ifc::TILE_M - 1) / ifc::TILE_M * ifc::TILE_M;
      // K2a: spectrum -> TF32-split rows + per-theta scalars
      ChordArgs a{};
      a.theta = theta + r0 * n_params; a.n = rows; a.chord_begin = c;
      a.lines = lines + r0 * m->I[I_N_ULINES] * BAYSAR_LINE_STRIDE;
      a.components = comps + r0 * ncols; a.spectra = nullptr;
      a.fine_pre = fine_pre ? fine_pre + r0 * F : nullptr; a.fine_post = nullptr; a.status = status + r0;
      PxcPlan& fx = m->pxc[c];
      const bool folded = fx.on && !fine_post;  // the fine-grid diagnostic needs every convolved column
      a.spec = m->spec; a.split_pitch = folded ? fx.pitch : pl.pitch; a.split_pad = folded ? fx.pad_left : pl.pad_left;
      a.aux = m->aux;
      if (launch_chord_kernel<1>(m, a, dim3((unsigned)rows, 1), m->chord_smem[c], stream,
                                 m->ch_i[c * CH_I_COUNT + CH_B_OPTIONAL] != 0))
        return 1;
      if (mark(m, ST_CHORD, stream)) return 1;
      if (folded) {
        // K2b': instrument function and wavelength map as one tcgen05 contraction; K2c': area identity + likelihood
        if (launch_pxc(fx, m->spec, (int)(rows128 / ifc::TILE_M), m->conv, m->trapz_part, m->sm_count, stream)) return 1;
        if (mark(m, ST_CONTRACT, stream)) return 1;
        PixelFoldArgs pf{};
        pf.theta = theta + r0 * n_params; pf.n = rows; pf.chord = c; pf.cp = m->conv; pf.cp_pitch = fx.out_pitch;
        pf.pix_col = fx.d_pix_col; pf.col_left0 = fx.col_left0; pf.x_l = fx.x_l; pf.col_right0 = fx.col_right0; pf.x_r = fx.x_r;
        pf.spec = m->spec; pf.spec_pitch = fx.pitch; pf.spec_pad = fx.pad_left; pf.dtab = fx.d_dtab; pf.n_dl = fx.n_dl;
        pf.n_dr = fx.n_dr; pf.wsum = fx.wsum; pf.inv_err = fx.d_inv_err; pf.tsum = fx.n_fold_tiles <= 64 ? m->trapz_part : nullptr;
        pf.n_col_tiles = fx.n_col_tiles; pf.n_fold_tiles = fx.n_fold_tiles <= 64 ? fx.n_fold_tiles : 0; pf.tile_n = fx.tn; pf.aux = m->aux; pf.components = comps + r0 * ncols;
        pf.spectra = spectra ? spectra + r0 * P : nullptr; pf.status = status + r0;
        {
          // a single-chord model gets its logP from the same kernel (no finalize launch)
          const bool fuse = logp && !pf.spectra && pf.inv_err && m->I[I_N_CHORDS] == 1;
          pf.logp = fuse ? logp + r0 : nullptr;
          if (fuse && finalized) *finalized = true;
          const size_t row_bytes = sizeof(float) * (size_t)fx.out_pitch;
          if (row_bytes > 160 * 1024) return fail("contraction row too long for the pixel stage's shared-memory copy");
          CUDA_OK(cudaFuncSetAttribute(pixel_fold_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)row_bytes));
          pixel_fold_kernel<128><<<(unsigned)rows, 128, row_bytes, stream>>>(m->dev, pf);
        }
        CUDA_OK(cudaGetLastError());
        if (mark(m, ST_PIXEL, stream)) return 1;
        m->last_launches += 3;
        continue;
      }
      // K2b: tcgen05 contraction with the banded-Toeplitz instrument matrix
      ifc::Params p{};
      p.s = m->spec; p.pitch = pl.pitch; p.pad_left = pl.pad_left; p.F = F;
      p.n_row_blocks = (int)(rows128 / ifc::TILE_M); p.m_lo = pl.m_lo; p.m_hi = pl.m_hi; p.taps_rev = m->ifc_taps_dev[c];
      p.taps_len = pl.taps_len; p.row_min = nullptr; p.aux = m->aux; p.aux_rows = (int)rows;
      p.out = m->conv; p.out_pitch = pl.out_pitch; p.trapz_part = m->trapz_part; p.n_tiles = pl.n_tiles;
      p.ring = ifc::ring_stages(pl.m_hi - pl.m_lo + 1);
      // static wavelength map and no fine-grid diagnostics requested: store only the columns the pixel stage reads
      const bool compact = m->ifc_cmask_dev[c] != nullptr && !fine_post;
      if (compact) {
        p.cmask = m->ifc_cmask_dev[c]; p.cbase = m->ifc_cbase_dev[c]; p.out_c = m->conv_c; p.out_c_pitch = pl.c_pitch;
      }
      const size_t smem = ifc::smem_bytes(pl.m_hi - pl.m_lo + 1);
      CUDA_OK(cudaFuncSetAttribute(ifc::if_contract_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      const int64_t tiles = (int64_t)p.n_row_blocks * pl.n_tiles;
      const unsigned ctas = (unsigned)(tiles < m->sm_count ? tiles : m->sm_count);  // persistent: one CTA per SM
      ifc::if_contract_kernel<false><<<ctas, ifc::THREADS, smem, stream>>>(p);
      CUDA_OK(cudaGetLastError());
      if (mark(m, ST_CONTRACT, stream)) return 1;
      // K2c: area renormalisation, wavelength map, likelihood
      PixelArgs pa{};
      pa.theta = theta + r0 * n_params; pa.n = rows; pa.chord = c; pa.conv = m->conv; pa.conv_pitch = pl.out_pitch;
      if (compact) { pa.conv = m->conv_c; pa.conv_pitch = pl.c_pitch; pa.cmask = p.cmask; pa.cbase = p.cbase; }
      pa.trapz_part = m->trapz_part; pa.n_tiles = pl.n_tiles; pa.aux = m->aux; pa.components = comps + r0 * ncols;
      pa.spectra = spectra ? spectra + r0 * P : nullptr; pa.fine_post = fine_post ? fine_post + r0 * F : nullptr;
      pa.status = status + r0;
      pixel_stage_kernel<128><<<(unsigned)rows, 128, 0, stream>>>(m->dev, pa);
      CUDA_OK(cudaGetLastError());
      if (mark(m, ST_PIXEL, stream)) return 1;
      m->last_launches += 3;
    }
  }
  return 0;
}

}  // namespace

extern "C" {

const char* baysar_last_error(void) { return g_last_error.c_str(); }
const char* baysar_version(void) { return "baysar_b200 0.1 (sm_100a)"; }

int baysar_model_create(const void* desc_blob, size_t nbytes, int device, baysar_model_t** out) {
  if (!desc_blob || !out || nbytes < 3 * sizeof(int64_t)) return fail("bad descriptor blob");
  const int64_t* hdr = static_cast<const int64_t*>(desc_blob);
  if (hdr[0] != BAYSAR_BLOB_MAGIC) return fail("descriptor magic mismatch");
  if (hdr[1] != BAYSAR_BLOB_VERSION) return fail("descriptor version mismatch (rebuild the model with this library)");
  if (hdr[2] != SEC_COUNT) return fail("descriptor section count mismatch");
  for (int s = 0; s < SEC_COUNT; ++s) {
    if (hdr[3 + 2 * s] < 0 || (size_t)hdr[3 + 2 * s] > nbytes) return fail("descriptor section offset out of range");
  }
  CUDA_OK(cudaSetDevice(device));
  baysar_model* m = new (std::nothrow) baysar_model();
  if (!m) return fail("out of host memory");
  m->device = device;
  m->blob_bytes = nbytes;
  m->blob_host.assign(static_cast<const unsigned char*>(desc_blob), static_cast<const unsigned char*>(desc_blob) + nbytes);
  cudaError_t e = cudaMalloc(&m->blob_dev, nbytes);
  if (e != cudaSuccess) { delete m; return fail(std::string("cudaMalloc(blob): ") + cudaGetErrorString(e)); }
  e = cudaMemcpy(m->blob_dev, desc_blob, nbytes, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { cudaFree(m->blob_dev); delete m; return fail(std::string("cudaMemcpy(blob): ") + cudaGetErrorString(e)); }

  SectionTable host{reinterpret_cast<const int64_t*>(m->blob_host.data()), m->blob_host.data()};
  SectionTable dev{reinterpret_cast<const int64_t*>(m->blob_host.data()), m->blob_dev};
  if (host.count(SEC_I) != I_COUNT) { baysar_model_destroy(m); return fail("descriptor I_COUNT mismatch"); }
  m->I = static_cast<const int64_t*>(host.ptr(SEC_I));
  m->ch_i = static_cast<const int*>(host.ptr(SEC_CH_I));
  m->cl_i = static_cast<const int*>(host.ptr(SEC_CL_I));
  ModelDev& d = m->dev;
  d.I = static_cast<const int64_t*>(dev.ptr(SEC_I));
  d.D = static_cast<const double*>(dev.ptr(SEC_D));
  d.los_x = static_cast<const double*>(dev.ptr(SEC_LOS_X));
  d.los_w = static_cast<const double*>(dev.ptr(SEC_LOS_W));
  d.prof_d = static_cast<const double*>(dev.ptr(SEC_PROF_D));
  d.sp_i = static_cast<const int*>(dev.ptr(SEC_SP_I));
  d.sp_d = static_cast<const double*>(dev.ptr(SEC_SP_D));
  d.elem_dens_idx = static_cast<const int*>(dev.ptr(SEC_ELEM_DENS_IDX));
  d.a11_ne = static_cast<const double*>(dev.ptr(SEC_A11_NE));
  d.a11_te = static_cast<const double*>(dev.ptr(SEC_A11_TE));
  d.a11_scd = static_cast<const double*>(dev.ptr(SEC_A11_SCD));
  d.h_ne = static_cast<const double*>(dev.ptr(SEC_H_NE));
  d.h_te = static_cast<const double*>(dev.ptr(SEC_H_TE));
  d.h_tab = static_cast<const double*>(dev.ptr(SEC_H_TAB));
  d.tau_grid = static_cast<const double*>(dev.ptr(SEC_TAU_GRID));
  d.spl_tx = static_cast<const double*>(dev.ptr(SEC_SPL_TX));
  d.spl_ty = static_cast<const double*>(dev.ptr(SEC_SPL_TY));
  d.spl_c = static_cast<const double*>(dev.ptr(SEC_SPL_C));
  d.ul_i = static_cast<const int*>(dev.ptr(SEC_UL_I));
  d.ul_d = static_cast<const double*>(dev.ptr(SEC_UL_D));
  d.pr_i = static_cast<const int*>(dev.ptr(SEC_PR_I));
  d.pr_d = static_cast<const double*>(dev.ptr(SEC_PR_D));
  d.cso_off = static_cast<const int*>(dev.ptr(SEC_CSO_OFF));
  d.cso_ul = static_cast<const int*>(dev.ptr(SEC_CSO_UL));
  d.ch_i = static_cast<const int*>(dev.ptr(SEC_CH_I));
  d.ch_d = static_cast<const double*>(dev.ptr(SEC_CH_D));
  d.cl_i = static_cast<const int*>(dev.ptr(SEC_CL_I));
  d.cl_d = static_cast<const double*>(dev.ptr(SEC_CL_D));
  d.comp_d = static_cast<const double*>(dev.ptr(SEC_COMP_D));
  d.y = static_cast<const double*>(dev.ptr(SEC_Y));
  d.err = static_cast<const double*>(dev.ptr(SEC_ERR));
  d.ifn = static_cast<const double*>(dev.ptr(SEC_IF));
  d.pix_idx = static_cast<const int*>(dev.ptr(SEC_PIX_IDX));
  d.pix_frac = static_cast<const double*>(dev.ptr(SEC_PIX_FRAC));
  d.mesh_x = static_cast<const double*>(dev.ptr(SEC_MESH_X));
  d.mesh_lo = static_cast<const int*>(dev.ptr(SEC_MESH_LO));
  d.pr_aux = static_cast<const int*>(dev.ptr(SEC_PR_AUX));

  // shared memory the chord kernel needs: the largest chord decides (one launch covers all chords)
  const int n_chords = (int)m->I[I_N_CHORDS];
  size_t smem_max = 0;
  for (int c = 0; c < n_chords; ++c) {
    const int* ch = m->ch_i + c * CH_I_COUNT;
    int kd_max = 0;
    for (int k = 0; k < ch[CH_N_LINES]; ++k) {
      int kd = m->cl_i[(ch[CH_LINE_OFF] + k) * CL_I_COUNT + CL_KD];
      kd_max = kd > kd_max ? kd : kd_max;
    }
    if (ch[CH_F] < 8) { baysar_model_destroy(m); return fail("fine grid shorter than 8 points"); }
    if (ch[CH_HALO] % 4 != 0) { baysar_model_destroy(m); return fail("CH_HALO must be a multiple of 4"); }
    ChordSmem s = chord_smem_layout(ch[CH_F], ch[CH_HALO], ch[CH_TAPS_MAX], kd_max);
    smem_max = s.total > smem_max ? s.total : smem_max;
    m->chord_smem.push_back(s.total);
  }
  if (smem_max > 227 * 1024) {
    baysar_model_destroy(m);
    return fail("fine grid too long for the shared-memory chord stage (needs " + std::to_string(smem_max) + " B)");
  }
  m->chord_smem_max = smem_max;
  cudaDeviceGetAttribute(&m->sm_count, cudaDevAttrMultiProcessorCount, device);
  if (m->sm_count <= 0) m->sm_count = 148;
  // tensor-core contraction for chords with a long, theta-independent instrument function
  const char* env = getenv("BAYSAR_IFC");
  const bool ifc_enabled = !(env && env[0] == '0');
  const char* env_taps = getenv("BAYSAR_IFC_MIN_TAPS");
  const int min_taps = env_taps ? atoi(env_taps) : 160;
  const char* env_budget = getenv("BAYSAR_IFC_BUDGET_MB");
  if (env_budget) m->ifc_budget_bytes = (int64_t)atoll(env_budget) << 20;
  const char* env_pxc = getenv("BAYSAR_PXC");  // 0: keep the every-column contraction; 64 / 128: column tile width
  const bool pxc_enabled = !(env_pxc && env_pxc[0] == '0');
  const int pxc_tn = (env_pxc && atoi(env_pxc) == 64) ? 64 : 128;
  m->pxc.resize(n_chords);
  m->ifc.resize(n_chords);
  m->ifc_taps_dev.assign(n_chords, nullptr);
  m->ifc_cmask_dev.assign(n_chords, nullptr);
  m->ifc_cbase_dev.assign(n_chords, nullptr);
  const int* pix_idx_host = static_cast<const int*>(host.ptr(SEC_PIX_IDX));
  const double* ifn_host = static_cast<const double*>(host.ptr(SEC_IF));
  for (int c = 0; c < n_chords; ++c) {
    const int* ch = m->ch_i + c * CH_I_COUNT;
    const int window = ch[CH_IF_KHI] - ch[CH_IF_KLO];
    if (!ifc_enabled || ch[CH_CALINT_IDX] >= 0 || window < min_taps || ch[CH_F] < 256) continue;
    m->ifc[c] = make_ifc_plan(ifn_host + ch[CH_IF_OFF], ch[CH_KI], ch[CH_IF_KLO], ch[CH_IF_KHI], ch[CH_F]);
    const IfcPlan& pl = m->ifc[c];
    if (pl.m_hi - pl.m_lo + 1 > ifc::MAX_KB) { m->ifc[c] = IfcPlan(); continue; }  // band too wide for the master tile
    cudaError_t e2 = cudaMalloc(&m->ifc_taps_dev[c], sizeof(float) * pl.taps_rev.size());
    if (e2 == cudaSuccess)
      e2 = cudaMemcpy(m->ifc_taps_dev[c], pl.taps_rev.data(), sizeof(float) * pl.taps_rev.size(), cudaMemcpyHostToDevice);
    if (e2 != cudaSuccess) { baysar_model_destroy(m); return fail(std::string("ifc taps upload: ") + cudaGetErrorString(e2)); }
    m->any_ifc = true;
    if (ch[CH_B_OPTIONAL]) {  // MODE 1 never touches buffer B on this chord
      int kd_max = 0;
      for (int k = 0; k < ch[CH_N_LINES]; ++k) {
        int kd = m->cl_i[(ch[CH_LINE_OFF] + k) * CL_I_COUNT + CL_KD];
        kd_max = kd > kd_max ? kd : kd_max;
      }
      m->chord_smem[c] = chord_smem_layout(ch[CH_F], ch[CH_HALO], ch[CH_TAPS_MAX], kd_max, 0).total;
    }
    if (ch[CH_CALWAVE_IDX] < 0) {  // static wavelength map: the pixel stage reads columns i_p and i_p + 1 only
      IfcPlan& plw = m->ifc[c];
      const int words = plw.n_tiles * (ifc::TILE_N / 32);
      plw.cmask.assign(words, 0u);
      plw.cbase.assign(words, 0);
      for (int px = 0; px < ch[CH_P]; ++px) {
        const int i = pix_idx_host[ch[CH_PIX_OFF] + px];
        for (int j = i; j <= i + 1; ++j)
          if (j >= 0 && j < ch[CH_F]) plw.cmask[j >> 5] |= 1u << (j & 31);
      }
      int run = 0;
      for (int w = 0; w < words; ++w) { plw.cbase[w] = run; run += __builtin_popcount(plw.cmask[w]); }
      plw.c_pitch = (run + 31) / 32 * 32;
      cudaError_t e3 = cudaMalloc(&m->ifc_cmask_dev[c], sizeof(uint32_t) * words);
      if (e3 == cudaSuccess) e3 = cudaMalloc(&m->ifc_cbase_dev[c], sizeof(int) * words);
      if (e3 == cudaSuccess) e3 = cudaMemcpy(m->ifc_cmask_dev[c], plw.cmask.data(), sizeof(uint32_t) * words, cudaMemcpyHostToDevice);
      if (e3 == cudaSuccess) e3 = cudaMemcpy(m->ifc_cbase_dev[c], plw.cbase.data(), sizeof(int) * words, cudaMemcpyHostToDevice);
      if (e3 != cudaSuccess) { baysar_model_destroy(m); return fail(std::string("compact map upload: ") + cudaGetErrorString(e3)); }
      m->conv_c_pitch_max = plw.c_pitch > m->conv_c_pitch_max ? plw.c_pitch : m->conv_c_pitch_max;
    }
    if (ch[CH_CALWAVE_IDX] < 0 && pxc_enabled) {  // static wavelength map: fold interp1d into the contraction
      const double* pix_frac_host = static_cast<const double*>(host.ptr(SEC_PIX_FRAC));
      PxcPlan fx = make_pxc_plan(ifn_host + ch[CH_IF_OFF], ch[CH_KI], ch[CH_IF_KLO], ch[CH_IF_KHI], ch[CH_F], ch[CH_P],
                                 pix_idx_host + ch[CH_PIX_OFF], pix_frac_host + ch[CH_PIX_OFF], pxc_tn);
      if (fx.on) {
        const double* err_host = static_cast<const double*>(host.ptr(SEC_ERR)) + ch[CH_PIX_OFF];
        fx.inv_err.resize(ch[CH_P]);
        for (int px = 0; px < ch[CH_P]; ++px) fx.inv_err[px] = (float)(1.0 / err_host[px]);
        if (upload_pxc_plan(fx)) { free_pxc_plan(fx); baysar_model_destroy(m); return 1; }
        m->pxc[c] = fx;
        m->split_pitch_max = fx.pitch > m->split_pitch_max ? fx.pitch : m->split_pitch_max;
        m->conv_pitch_max = fx.out_pitch > m->conv_pitch_max ? fx.out_pitch : m->conv_pitch_max;
        m->tiles_max = 2 * fx.n_col_tiles > m->tiles_max ? 2 * fx.n_col_tiles : m->tiles_max;  // tsum lives in trapz_part
      }
    }
    m->split_pitch_max = pl.pitch > m->split_pitch_max ? pl.pitch : m->split_pitch_max;
    m->conv_pitch_max = pl.out_pitch > m->conv_pitch_max ? pl.out_pitch : m->conv_pitch_max;
    m->tiles_max = pl.n_tiles > m->tiles_max ? pl.n_tiles : m->tiles_max;
  }
  *out = m;
  return 0;
}

void baysar_model_destroy(baysar_model_t* m) {
  if (!m) return;
  cudaSetDevice(m->device);
  cudaFree(m->blob_dev); cudaFree(m->lines); cudaFree(m->comps); cudaFree(m->hdr); cudaFree(m->status);
  cudaFree(m->theta_stage); cudaFree(m->logp_stage); cudaFreeHost(m->out_pinned);
  cudaFree(m->spec); cudaFree(m->conv); cudaFree(m->conv_c); cudaFree(m->trapz_part); cudaFree(m->aux);
  for (float* t : m->ifc_taps_dev) cudaFree(t);
  for (uint32_t* t : m->ifc_cmask_dev) cudaFree(t);
  for (int* t : m->ifc_cbase_dev) cudaFree(t);
  for (PxcPlan& fx : m->pxc) free_pxc_plan(fx);
  for (cudaEvent_t ev : m->prof_events) cudaEventDestroy(ev);
  delete m;
}

int64_t baysar_model_query(const baysar_model_t* m, int what, int arg) {
  if (!m) return -1;
  switch (what) {
    case 0: return m->I[I_N_PARAMS];
    case 1: return m->I[I_N_CHORDS];
    case 2: return m->I[I_N_ULINES];
    case 3: return m->I[I_N_PRIORS];
    case 4: return m->I[I_P_MAX];
    case 5: return m->I[I_F_MAX];
    case 6: return (arg >= 0 && arg < m->I[I_N_CHORDS]) ? m->ch_i[arg * CH_I_COUNT + CH_P] : -1;
    case 7: return (arg >= 0 && arg < m->I[I_N_CHORDS]) ? m->ch_i[arg * CH_I_COUNT + CH_F] : -1;
    case 8: return m->I[I_L];
    case 9: return m->last_launches;
    case 11: return m->any_ifc ? 1 : 0;
    case 12: {  // folded pixel contraction of chord `arg`: k-blocks per theta block (0: not folded)
      if (arg < 0 || arg >= m->I[I_N_CHORDS] || !m->pxc[arg].on) return 0;
      int64_t kb = 0;
      for (int t : m->pxc[arg].tile_nkb) kb += t;
      return kb;
    }
    case 13: return (arg >= 0 && arg < m->I[I_N_CHORDS] && m->pxc[arg].on) ? m->pxc[arg].tn : 0;
    case 10: return (int64_t)m->chord_smem_max;
    default: return -1;
  }
}

int baysar_eval_logp(baysar_model_t* m, const double* theta, int64_t n, double* logp, uint32_t* status,
                     double* components, void* stream_) {
  if (!m || !theta || !logp) return fail("null argument");
  if (n <= 0) return 0;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CUDA_OK(cudaSetDevice(m->device));
  if (ensure_workspace(m, n)) return 1;
  uint32_t* st = status ? status : m->status;
  double* comps = components ? components : m->comps;
  m->last_launches = 0;
  if (mark(m, ST_BEGIN, stream)) return 1;
  if (launch_los(m, theta, n, m->lines, comps, nullptr, st, stream)) return 1;
  m->last_launches += 1;
  if (mark(m, ST_LOS, stream)) return 1;
  bool finalized = false;
  if (launch_chords(m, theta, n, 0, (int)m->I[I_N_CHORDS], m->lines, comps, nullptr, nullptr, nullptr, st, stream, logp,
                    &finalized))
    return 1;
  if (!finalized) {
    const int ncols = (int)(m->I[I_N_CHORDS] + m->I[I_N_PRIORS]);
    finalize_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(comps, ncols, n, logp, st);
    CUDA_OK(cudaGetLastError());
    m->last_launches += 1;
  }
  if (mark(m, ST_FINAL, stream)) return 1;
  return 0;
}

int baysar_eval_spectra(baysar_model_t* m, const double* theta, int64_t n, int chord, double* spectra,
                        float* fine_pre, float* fine_post, uint32_t* status, void* stream_) {
  if (!m || !theta || !spectra) return fail("null argument");
  if (chord < 0 || chord >= m->I[I_N_CHORDS]) return fail("chord index out of range");
  if (n <= 0) return 0;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CUDA_OK(cudaSetDevice(m->device));
  if (ensure_workspace(m, n)) return 1;
  uint32_t* st = status ? status : m->status;
  if (launch_los(m, theta, n, m->lines, m->comps, nullptr, st, stream)) return 1;
  return launch_chords(m, theta, n, chord, 1, m->lines, m->comps, spectra, fine_pre, fine_post, st, stream);
}

int baysar_eval_lines(baysar_model_t* m, const double* theta, int64_t n, double* lines, double* profiles,
                      uint32_t* status, void* stream_) {
  if (!m || !theta || !lines) return fail("null argument");
  if (n <= 0) return 0;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CUDA_OK(cudaSetDevice(m->device));
  if (ensure_workspace(m, n)) return 1;
  uint32_t* st = status ? status : m->status;
  return launch_los(m, theta, n, lines, m->comps, profiles, st, stream);
}

int baysar_eval_logp_host(baysar_model_t* m, const double* theta_host, int64_t n, double* logp_host,
                          uint32_t* status_host) {
  if (!m || !theta_host || !logp_host) return fail("null argument");
  if (n <= 0) return 0;
  CUDA_OK(cudaSetDevice(m->device));
  const int64_t np = m->I[I_N_PARAMS];
  if (n > m->stage_cap) {
    cudaFree(m->theta_stage); cudaFree(m->logp_stage); cudaFreeHost(m->out_pinned);
    m->theta_stage = nullptr; m->logp_stage = nullptr; m->out_pinned = nullptr; m->stage_cap = 0;
    int64_t cap = n < 1024 ? 1024 : n;
    CUDA_OK(cudaMalloc(&m->theta_stage, sizeof(double) * cap * np));
    // logP and the status words of one call sit back to back ([n] float64, [n] uint32): one device-to-host copy into a
    // pinned staging buffer (a copy into the caller's pageable arrays is staged by the driver, twice)
    CUDA_OK(cudaMalloc(&m->logp_stage, (sizeof(double) + sizeof(uint32_t)) * cap));
    CUDA_OK(cudaHostAlloc(&m->out_pinned, (sizeof(double) + sizeof(uint32_t)) * cap, cudaHostAllocDefault));
    m->stage_cap = cap;
  }
  uint32_t* status_dev = reinterpret_cast<uint32_t*>(m->logp_stage + n);
  CUDA_OK(cudaMemcpyAsync(m->theta_stage, theta_host, sizeof(double) * n * np, cudaMemcpyHostToDevice, 0));
  if (baysar_eval_logp(m, m->theta_stage, n, m->logp_stage, status_dev, nullptr, nullptr)) return 1;
  const size_t out_bytes = (sizeof(double) + (status_host ? sizeof(uint32_t) : 0)) * (size_t)n;
  CUDA_OK(cudaMemcpyAsync(m->out_pinned, m->logp_stage, out_bytes, cudaMemcpyDeviceToHost, 0));
  CUDA_OK(cudaStreamSynchronize(0));
  memcpy(logp_host, m->out_pinned, sizeof(double) * n);
  if (status_host) memcpy(status_host, static_cast<unsigned char*>(m->out_pinned) + sizeof(double) * n, sizeof(uint32_t) * n);
  return 0;
}

int baysar_model_set_profiling(baysar_model_t* m, int on) {
  if (!m) return fail("null model");
  for (cudaEvent_t ev : m->prof_events) cudaEventDestroy(ev);
  m->prof_events.clear();
  m->prof_slots.clear();
  m->profiling = on != 0;
  return 0;
}

int baysar_model_get_profile(baysar_model_t* m, double* stage_ms, int64_t* calls) {
  if (!m || !stage_ms || !calls) return fail("null argument");
  CUDA_OK(cudaSetDevice(m->device));
  for (int k = 0; k < ST_COUNT; ++k) stage_ms[k] = 0.0;
  *calls = 0;
  for (size_t e = 0; e < m->prof_events.size(); ++e) {
    if (m->prof_slots[e] == ST_BEGIN) { *calls += 1; continue; }
    CUDA_OK(cudaEventSynchronize(m->prof_events[e]));
    float ms = 0.f;
    CUDA_OK(cudaEventElapsedTime(&ms, m->prof_events[e - 1], m->prof_events[e]));
    stage_ms[m->prof_slots[e]] += ms;
  }
  for (cudaEvent_t ev : m->prof_events) cudaEventDestroy(ev);
  m->prof_events.clear();
  m->prof_slots.clear();
  return 0;
}

int baysar_debug_if_contract(const float* spectra_dev, int64_t n_rows, int F, const double* taps_host, int KI,
                             const float* row_min_dev, float* out_dev, double* trapz_dev, int64_t* out_pitch,
                             int* n_tiles, float* elapsed_ms) {
  // Unit-test hook for the tcgen05 contraction: out[r][j] = max(sum_k w[k] S[r][j + (KI-1)/2 - k], row_min[r]).
  if (!spectra_dev || !taps_host || !out_dev) return fail("null argument");
  if (n_rows % ifc::TILE_M != 0) return fail("n_rows must be a multiple of 128");
  double wmax = 0.0;
  for (int k = 0; k < KI; ++k) wmax = fabs(taps_host[k]) > wmax ? fabs(taps_host[k]) : wmax;
  int k_lo = KI, k_hi = 0;
  for (int k = 0; k < KI; ++k)
    if (fabs(taps_host[k]) > 1e-17 * wmax) { k_lo = k < k_lo ? k : k_lo; k_hi = k + 1; }
  IfcPlan pl = make_ifc_plan(taps_host, KI, k_lo, k_hi, F);
  if (out_pitch) *out_pitch = pl.out_pitch;
  if (n_tiles) *n_tiles = pl.n_tiles;
  if (pl.m_hi - pl.m_lo + 1 > ifc::MAX_KB) return fail("instrument function too wide for the tensor-core contraction");
  float *spec = nullptr, *taps = nullptr;
  CUDA_OK(cudaMalloc(&spec, sizeof(float) * n_rows * pl.pitch));
  CUDA_OK(cudaMalloc(&taps, sizeof(float) * pl.taps_rev.size()));
  CUDA_OK(cudaMemset(spec, 0, sizeof(float) * n_rows * pl.pitch));
  CUDA_OK(cudaMemcpy(taps, pl.taps_rev.data(), sizeof(float) * pl.taps_rev.size(), cudaMemcpyHostToDevice));
  ifc::pad_rows_kernel<<<dim3(8, (unsigned)n_rows), 256>>>(spectra_dev, n_rows, F, spec, pl.pitch, pl.pad_left);
  CUDA_OK(cudaGetLastError());
  ifc::Params p{};
  p.s = spec; p.pitch = pl.pitch; p.pad_left = pl.pad_left; p.F = F; p.n_row_blocks = (int)(n_rows / ifc::TILE_M);
  p.m_lo = pl.m_lo; p.m_hi = pl.m_hi; p.taps_rev = taps; p.taps_len = pl.taps_len; p.row_min = row_min_dev; p.aux = nullptr; p.aux_rows = 0;
  p.out = out_dev; p.out_pitch = pl.out_pitch; p.trapz_part = trapz_dev; p.n_tiles = pl.n_tiles;
  p.ring = ifc::ring_stages(pl.m_hi - pl.m_lo + 1);
  p.debug = getenv("BAYSAR_IFC_DEBUG") ? atoi(getenv("BAYSAR_IFC_DEBUG")) : 0;
  long long* dbg = nullptr;
  if (getenv("BAYSAR_IFC_TRACE")) {
    CUDA_OK(cudaMalloc(&dbg, sizeof(long long) * 8 * 256));
    CUDA_OK(cudaMemset(dbg, 0, sizeof(long long) * 8 * 256));
    p.dbg_cycles = dbg;
  }
  const size_t smem = ifc::smem_bytes(pl.m_hi - pl.m_lo + 1);
  const bool dbg_kernel = p.debug != 0 || dbg != nullptr;
  CUDA_OK(cudaFuncSetAttribute(ifc::if_contract_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CUDA_OK(cudaFuncSetAttribute(ifc::if_contract_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1;
  CUDA_OK(cudaEventCreate(&e0));
  CUDA_OK(cudaEventCreate(&e1));
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int64_t tiles = (int64_t)p.n_row_blocks * pl.n_tiles;
  const unsigned ctas = (unsigned)(tiles < sms ? tiles : sms);
  for (int rep = 0; rep < 2; ++rep) {
    CUDA_OK(cudaEventRecord(e0, 0));
    if (dbg_kernel) ifc::if_contract_kernel<true><<<ctas, ifc::THREADS, smem>>>(p);
    else ifc::if_contract_kernel<false><<<ctas, ifc::THREADS, smem>>>(p);
    CUDA_OK(cudaEventRecord(e1, 0));
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaEventSynchronize(e1));
  }
  float ms = 0.f;
  CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
  if (elapsed_ms) *elapsed_ms = ms;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  if (dbg) {
    std::vector<long long> h(8 * 256);
    cudaMemcpy(h.data(), dbg, sizeof(long long) * 8 * 256, cudaMemcpyDeviceToHost);
    const char* names[8] = {"conv.wait_landed", "conv.wait_empty", "conv.wait_st", "conv.total", "load.wait_consumed",
                            "epi.wait_acc_full", "mma.wait_full", "mma.wait_acc_free"};
    for (int k = 0; k < 8; ++k) {
      double mean = 0;
      for (unsigned b = 0; b < ctas; ++b) mean += (double)h[b * 8 + k];
      fprintf(stderr, "[ifc trace] %-20s mean %.0f cycles per CTA (CTA0 %lld)\n", names[k], mean / ctas, h[k]);
    }
    cudaFree(dbg);
  }
  cudaFree(spec); cudaFree(taps);
  return 0;
}

int baysar_debug_pix_contract(const float* spectra_dev, int64_t n_rows, int F, const double* taps_host, int KI,
                              const int* pix_idx_host, const double* pix_frac_host, int P, int tile_n, float* out_dev,
                              int64_t out_capacity, double* tsum_dev, int64_t tsum_capacity, int64_t* info,
                              int* pix_col_host, float* elapsed_ms) {
  // Unit-test hook for the folded contraction: out[r][c] for the plan's columns (see include/baysar_b200.h).
  if (!spectra_dev || !taps_host || !pix_idx_host || !pix_frac_host || !info) return fail("null argument");
  if (n_rows % ifc::TILE_M != 0) return fail("n_rows must be a multiple of 128");
  double wmax = 0.0;
  for (int k = 0; k < KI; ++k) wmax = fabs(taps_host[k]) > wmax ? fabs(taps_host[k]) : wmax;
  int k_lo = KI, k_hi = 0;
  for (int k = 0; k < KI; ++k)
    if (fabs(taps_host[k]) > 1e-17 * wmax) { k_lo = k < k_lo ? k : k_lo; k_hi = k + 1; }
  PxcPlan pl = make_pxc_plan(taps_host, KI, k_lo, k_hi, F, P, pix_idx_host, pix_frac_host, tile_n == 64 ? 64 : 128);
  if (!pl.on) return fail("no folded-contraction plan for this instrument function / grid");
  int64_t kb_total = 0;
  for (int t = 0; t < pl.n_col_tiles; ++t) kb_total += pl.tile_nkb[t];
  info[0] = pl.out_pitch; info[1] = pl.col_left0; info[2] = pl.x_l; info[3] = pl.col_right0; info[4] = pl.x_r;
  info[5] = pl.n_col_tiles; info[6] = kb_total; info[7] = pl.ring;
  if (pix_col_host) memcpy(pix_col_host, pl.pix_col.data(), sizeof(int) * P);
  if (!out_dev) return 0;  // layout query only
  if (out_capacity < n_rows * pl.out_pitch) return fail("output buffer too small");
  if (upload_pxc_plan(pl)) { free_pxc_plan(pl); return 1; }
  float* spec = nullptr;
  if (tsum_dev && tsum_capacity < n_rows * pl.n_col_tiles * 2) return fail("tsum buffer too small");
  cudaError_t e = cudaMalloc(&spec, sizeof(float) * n_rows * pl.pitch);
  if (e == cudaSuccess) e = cudaMemset(spec, 0, sizeof(float) * n_rows * pl.pitch);
  if (e != cudaSuccess) { free_pxc_plan(pl); return fail(std::string("debug spectra: ") + cudaGetErrorString(e)); }
  ifc::pad_rows_kernel<<<dim3(8, (unsigned)n_rows), 256>>>(spectra_dev, n_rows, F, spec, pl.pitch, pl.pad_left);
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  int rc = 0;
  long long* dbg = nullptr;
  if (getenv("BAYSAR_PXC_TRACE")) {
    cudaMalloc(&dbg, sizeof(long long) * 16 * 256);
    cudaMemset(dbg, 0, sizeof(long long) * 16 * 256);
  }
  for (int rep = 0; rep < 2 && rc == 0; ++rep) {
    cudaEventRecord(e0, 0);
    rc = launch_pxc(pl, spec, (int)(n_rows / ifc::TILE_M), out_dev, tsum_dev, sms, 0, dbg);
    cudaEventRecord(e1, 0);
    if (rc == 0 && cudaEventSynchronize(e1) != cudaSuccess) rc = fail(std::string("folded contraction: ") + cudaGetErrorString(cudaGetLastError()));
  }
  float ms = 0.f;
  if (rc == 0) cudaEventElapsedTime(&ms, e0, e1);
  if (elapsed_ms) *elapsed_ms = ms;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  if (dbg && rc == 0) {
    std::vector<long long> h(16 * 256);
    cudaMemcpy(h.data(), dbg, sizeof(long long) * 16 * 256, cudaMemcpyDeviceToHost);
    const char* names[13] = {"conv.wait_landed", "conv.wait_empty", "conv.total", "load.wait_consumed", "build.wait_bempty",
                             "build.total", "epi.wait_acc_full", "epi.drain", "mma.wait_full", "mma.wait_bfull",
                             "mma.wait_acc_free", "mma.total", "mma.k_blocks"};
    const int ctas = (int)std::min<int64_t>(sms, (n_rows / ifc::TILE_M) * pl.n_col_tiles);
    for (int k = 0; k < 13; ++k) {
      double mean = 0, mx = 0;
      for (int b = 0; b < ctas; ++b) { mean += (double)h[b * 16 + k]; mx = std::max(mx, (double)h[b * 16 + k]); }
      fprintf(stderr, "[pxc trace] %-20s mean %.0f max %.0f (CTA0 %lld)\n", names[k], mean / ctas, mx, h[k]);
    }
  }
  cudaFree(dbg);
  cudaFree(spec);
  free_pxc_plan(pl);
  return rc;
}

int baysar_debug_pix_plan_check(const double* spectrum_host, int F, const double* taps_host, int KI,
                                const int* pix_idx_host, const double* pix_frac_host, int P, int tile_n, double* result) {
  // Host-only self check of the folded-contraction plan (no GPU needed): evaluates every column through the plan's
  // tables with the kernel's index arithmetic (float64 products) against the definition, and the area identity of
  // pixel_fold_kernel against the direct trapz of the clipped convolution.
  if (!spectrum_host || !taps_host || !pix_idx_host || !pix_frac_host || !result) return fail("null argument");
  double wmax = 0.0;
  for (int k = 0; k < KI; ++k) wmax = fabs(taps_host[k]) > wmax ? fabs(taps_host[k]) : wmax;
  int k_lo = KI, k_hi = 0;
  for (int k = 0; k < KI; ++k)
    if (fabs(taps_host[k]) > 1e-17 * wmax) { k_lo = k < k_lo ? k : k_lo; k_hi = k + 1; }
  PxcPlan pl = make_pxc_plan(taps_host, KI, k_lo, k_hi, F, P, pix_idx_host, pix_frac_host, tile_n == 64 ? 64 : 128);
  if (!pl.on) return fail("no folded-contraction plan for this instrument function / grid");
  const int o = (KI - 1) / 2, tn = pl.tn;
  std::vector<double> conv(F);
  double mn = spectrum_host[0], cmax = 0.0;
  for (int j = 0; j < F; ++j) {
    mn = spectrum_host[j] < mn ? spectrum_host[j] : mn;
    double acc = 0.0;
    for (int k = k_lo; k < k_hi; ++k) {
      const int sidx = j + o - k;
      if (sidx >= 0 && sidx < F) acc += (double)(float)taps_host[k] * spectrum_host[sidx];
    }
    conv[j] = acc;
    cmax = fabs(acc) > cmax ? fabs(acc) : cmax;
  }
  auto padded = [&](int64_t col) {  // the zero-padded row the kernel reads
    const int64_t sidx = col - pl.pad_left;
    return (sidx >= 0 && sidx < F) ? spectrum_host[sidx] : 0.0;
  };
  std::vector<double> out(pl.out_pitch);
  for (int t = 0; t < pl.n_col_tiles; ++t)
    for (int c = t * tn; c < (t + 1) * tn; ++c) {
      double acc = 0.0;
      for (int kb = 0; kb < pl.tile_nkb[t]; ++kb)
        for (int e = 0; e < ifc::TILE_K; ++e) {
          const int x = pl.col_idx[c] - ifc::TILE_K * kb - e;
          if (x - 3 < 0 || x + 1 >= pl.wlen) return fail("tap table index out of range");
          const double b = (double)pl.wtab[x] + (double)pl.col_frac[c] * ((double)pl.wtab[x + 1] - (double)pl.wtab[x]);
          acc += padded((int64_t)pl.tile_kcol0[t] + ifc::TILE_K * kb + e) * b;
        }
      out[c] = acc;
    }
  double err_pix = 0.0, err_edge = 0.0;
  int n_folded = 0;
  for (int p = 0; p < P; ++p) {
    if (pl.pix_col[p] < 0) continue;
    ++n_folded;
    const int i = pix_idx_host[p];
    const double ref = conv[i] + (double)(float)pix_frac_host[p] * (conv[i + 1] - conv[i]);
    err_pix = std::max(err_pix, fabs(out[pl.pix_col[p]] - ref) / cmax);
  }
  for (int j = 0; j < pl.x_l; ++j) err_edge = std::max(err_edge, fabs(out[pl.col_left0 + j] - conv[j]) / cmax);
  for (int j = F - pl.x_r; j < F; ++j) err_edge = std::max(err_edge, fabs(out[pl.col_right0 + j - (F - pl.x_r)] - conv[j]) / cmax);
  // area identity
  double pre = 0.0, post_direct = 0.0, part = 0.0;
  for (int j = 0; j < F; ++j) {
    const double cj = (j == 0 || j == F - 1) ? 0.5 : 1.0;
    pre += cj * spectrum_host[j];
    post_direct += cj * (conv[j] < mn ? mn : conv[j]);
  }
  for (int t = 0; t < pl.x_l + pl.x_r; ++t) {
    const int j = t < pl.x_l ? t : F - pl.x_r + (t - pl.x_l);
    const double v = t < pl.x_l ? out[pl.col_left0 + t] : out[pl.col_right0 + (t - pl.x_l)];
    if (v < mn) part += ((j == 0 || j == F - 1) ? 0.5 : 1.0) * (mn - v);
  }
  for (int t = 0; t < pl.n_dl + pl.n_dr; ++t) {
    const int sidx = t < pl.n_dl ? t : F - pl.n_dr + (t - pl.n_dl);
    part += spectrum_host[sidx] * pl.dtab[t];
  }
  // column-sum identity per folded tile
  double err_cs = 0.0;
  for (int t = 0; t < pl.n_fold_tiles && pl.tile_boff[t] >= 0; ++t) {
    double col_sum = 0.0, t_id = 0.0;
    for (int c = t * tn; c < (t + 1) * tn; ++c) col_sum += out[c];
    for (int x = 0; x < ifc::TILE_K * pl.tile_nkb[t]; ++x)
      t_id += padded((int64_t)pl.tile_kcol0[t] + x) * (double)pl.bbar[pl.tile_boff[t] + x];
    err_cs = std::max(err_cs, fabs(t_id - col_sum) / fabs(col_sum));
  }
  result[8] = err_cs;
  // interior columns must not need the clip
  double clip_violation = 0.0;
  for (int j = pl.e_l; j < F - pl.e_r; ++j) clip_violation = std::max(clip_violation, (mn - conv[j]) / cmax);
  int64_t kb_total = 0;
  for (int t = 0; t < pl.n_col_tiles; ++t) kb_total += pl.tile_nkb[t];
  result[0] = err_pix; result[1] = err_edge; result[2] = fabs((pl.wsum * pre + part) - post_direct) / fabs(post_direct);
  result[3] = (double)kb_total; result[4] = (double)n_folded; result[5] = clip_violation; result[6] = (double)pl.wlen;
  result[7] = (double)pl.ring;
  return 0;
}

int baysar_stretch_propose(const double* walkers, const double* complement, int n_params, int64_t n_walkers,
                           int64_t n_complement, double a, uint64_t seed, uint64_t step, double* proposal,
                           double* log_factor, void* stream_) {
  if (!walkers || !complement || !proposal || !log_factor) return fail("null argument");
  if (n_params < 1 || n_complement < 1 || !(a > 1.0)) return fail("bad stretch-move arguments");
  if (n_walkers <= 0) return 0;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ens::stretch_propose_kernel<<<(unsigned)((n_walkers + 127) / 128), 128, 0, stream>>>(
      walkers, complement, n_params, n_walkers, n_complement, a, seed, step, proposal, log_factor);
  CUDA_OK(cudaGetLastError());
  return 0;
}

int baysar_stretch_accept(double* walkers, double* logp, const double* proposal, const double* logp_new,
                          const double* log_factor, const uint32_t* status_new, int n_params, int64_t n_walkers,
                          uint64_t seed, uint64_t step, unsigned long long* n_accepted, void* stream_) {
  if (!walkers || !logp || !proposal || !logp_new || !log_factor) return fail("null argument");
  if (n_walkers <= 0) return 0;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ens::stretch_accept_kernel<<<(unsigned)((n_walkers + 127) / 128), 128, 0, stream>>>(
      walkers, logp, proposal, logp_new, log_factor, status_new, n_params, n_walkers, seed, step, n_accepted);
  CUDA_OK(cudaGetLastError());
  return 0;
}

int baysar_debug_mma_probe(int device, int tile_n, int form, int n_acc_hi, int b_stages, int pattern, int iters,
                           double* cycles_per_mma) {
  if (!cycles_per_mma) return fail("null argument");
  if (tile_n < 8 || tile_n > 256 || tile_n % 8 || n_acc_hi < 1 || (n_acc_hi + 1) * tile_n > 512 || b_stages < 1 ||
      b_stages > 4 || iters < 1)
    return fail("bad probe arguments");
  CUDA_OK(cudaSetDevice(device));
  int sms = 0;
  CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  long long* out = nullptr;
  CUDA_OK(cudaMalloc(&out, sizeof(long long) * sms));
  CUDA_OK(cudaMemset(out, 0, sizeof(long long) * sms));
  const size_t smem = (size_t)4 * 256 * 128 + 128 * 128 + 1024;
  CUDA_OK(cudaFuncSetAttribute(mmaprobe::mma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  for (int rep = 0; rep < 2; ++rep) {
    mmaprobe::mma_probe_kernel<<<sms, 128, smem>>>(tile_n, form, n_acc_hi, b_stages, pattern, iters, out);
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaDeviceSynchronize());
  }
  std::vector<long long> h(sms);
  CUDA_OK(cudaMemcpy(h.data(), out, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
  cudaFree(out);
  double worst = 0.0;
  for (long long c : h) {
    if (c < 0) return fail("probe timed out");
    worst = (double)c > worst ? (double)c : worst;
  }
  *cycles_per_mma = worst / ((double)iters * (pattern == 2 ? 2.0 : 6.0));
  return 0;
}

int baysar_microbench(int device, int what, double* result) {
  if (!result) return fail("null argument");
  CUDA_OK(cudaSetDevice(device));
  int sms = 0;
  CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  float* out = nullptr;
  CUDA_OK(cudaMalloc(&out, sizeof(float) * 1024));
  const int iters = 4096, blocks = sms * 8, threads = 256;
  cudaEvent_t e0, e1;
  CUDA_OK(cudaEventCreate(&e0));
  CUDA_OK(cudaEventCreate(&e1));
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    CUDA_OK(cudaEventRecord(e0, 0));
    if (what == 0) microbench_ffma_kernel<<<blocks, threads>>>(out, iters, 1.0001f);
    else if (what == 1) microbench_mufu_kernel<<<blocks, threads>>>(out, iters, 1.0001f);
    else if (what == 3) microbench_ffma_reg_kernel<<<blocks, threads>>>(out, iters, 1.0001f, 0.5f);
    else if (what == 4) microbench_ffma2_kernel<<<blocks, threads>>>(out, iters, 1.0001f, 0.5f);
    else microbench_dfma_kernel<<<blocks, threads>>>(reinterpret_cast<double*>(out), iters, 1.0001);
    CUDA_OK(cudaEventRecord(e1, 0));
    CUDA_OK(cudaEventSynchronize(e1));
    float ms = 0.f;
    CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
    // ops per thread per iteration: 16 independent chains x 1 instruction
    double ops = (double)blocks * threads * (double)iters * 16.0 * (what == 4 ? 2.0 : 1.0);
    double rate = ops / (ms * 1e-3);
    if (rep > 0 && rate > best) best = rate;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  *result = what == 1 ? best : 2.0 * best;  // FMA = 2 flop; MUFU counted as ops
  return 0;
}

}  // extern "C"
