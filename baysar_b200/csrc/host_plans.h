// Host-side plans of the two tcgen05 contractions (if_contract.cuh, pix_contract.cuh): packing of one chord's instrument
// function, column tiles of the folded pixel contraction, per-batch-size schedules, and the launch helpers.  Shared by
// the product library (baysar_abi.cu) and the measurement / unit-test library (baysar_tools.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <new>
#include <string>
#include <utility>
#include <vector>

#include "if_contract.cuh"
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

// Kernel launch with programmatic dependent launch (PDL): the grid may be scheduled while the previous kernel of the
// stream is still draining its last wave, so launch latency and the prologue of this kernel overlap with that tail.
// Every kernel of the evaluation executes griddepcontrol.wait (pdl_wait) before its first access to global memory, which
// blocks until the previous grid has completed and its writes are visible: stream order is unchanged, only the gaps
// between the kernels close.  `pdl = false` launches plainly (options.serial_launches, for A/B timing).
template <class... KArgs, class... Args>
inline cudaError_t launch_k(bool pdl, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                            Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
}

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
  bool no_extrapolation = false;  // every pixel of the static wavelength map lies between two fine-grid points
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
  // B operand as a precomputed image (pix_contract.cuh b_image_kernel): one stage-sized tile per k-block of every column tile
  std::vector<int> tile_bimg;
  int* d_tile_bimg = nullptr;
  unsigned char* d_bimg = nullptr;
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

namespace {

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
  if (pl.n_fold_tiles <= pxc::MAX_FOLD_TILES) {
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
  // the pixel stage keeps one correction / scale per folded tile in 64-entry shared-memory tables: wider chords (P > 8192
  // at 128-column tiles) stay on the every-column Toeplitz contraction
  pl.on = pl.ring >= 3 && pl.n_col_tiles < 32768 && pl.n_fold_tiles <= pxc::MAX_FOLD_TILES;
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
  cudaFree(pl.d_tile_bimg); cudaFree(pl.d_bimg);
  pl.d_tile_bimg = nullptr; pl.d_bimg = nullptr;
  for (auto& kv : pl.schedules) { cudaFree(kv.second.d_sched); cudaFree(kv.second.d_cnt); }
  pl.schedules.clear();
}

// The theta-independent B operand of the folded contraction, written once per plan: 16 KB (TN = 128) per k-block of every
// column tile, in the bytes a ring stage holds (Balmer demo: 790 tiles, 12.9 MB -- it stays in L2 while the 32 theta blocks
// of a 4096-theta step read it).  Plans whose image would exceed `max_bytes` keep the builder warps.
int build_pxc_b_image(PxcPlan& pl, size_t max_bytes = (size_t)512 << 20) {
  int64_t total = 0;
  int max_kb = 0;
  pl.tile_bimg.resize(pl.n_col_tiles);
  for (int t = 0; t < pl.n_col_tiles; ++t) {
    pl.tile_bimg[t] = (int)total;
    total += pl.tile_nkb[t];
    max_kb = std::max(max_kb, pl.tile_nkb[t]);
  }
  const size_t stage = (size_t)pl.tn * 128;
  if (total == 0 || (size_t)total * stage > max_bytes || max_kb > 65535) return 0;
  if (upload(&pl.d_tile_bimg, pl.tile_bimg)) return 1;
  CUDA_OK(cudaMalloc(&pl.d_bimg, (size_t)total * stage));
  const dim3 grid((unsigned)pl.n_col_tiles, (unsigned)max_kb);
  if (pl.tn == 64)
    pxc::b_image_kernel<64><<<grid, 64>>>(pl.d_tile_nkb, pl.d_tile_bimg, pl.d_col_idx, pl.d_col_frac, pl.d_wtab, pl.wlen, pl.d_bimg);
  else
    pxc::b_image_kernel<128><<<grid, 64>>>(pl.d_tile_nkb, pl.d_tile_bimg, pl.d_col_idx, pl.d_col_frac, pl.d_wtab, pl.wlen, pl.d_bimg);
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaDeviceSynchronize());
  return 0;
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

// Opt-in shared-memory limit of a kernel (227 KB): set once per device and library at model creation, not per launch
#define BAYSAR_SMEM_OPT_IN 232448
template <class K>
int allow_max_smem(K kern) {
  cudaFuncAttributes fa;
  CUDA_OK(cudaFuncGetAttributes(&fa, kern));  // the opt-in limit covers static + dynamic shared memory
  CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, BAYSAR_SMEM_OPT_IN - (int)fa.sharedSizeBytes));
  return 0;
}
int contraction_kernel_attributes() {
  return allow_max_smem(pxc::pix_contract_kernel<64, false>) || allow_max_smem(pxc::pix_contract_kernel<64, true>) ||
         allow_max_smem(pxc::pix_contract_kernel<128, false>) || allow_max_smem(pxc::pix_contract_kernel<128, true>) ||
         allow_max_smem(ifc::if_contract_kernel<false>) || allow_max_smem(ifc::if_contract_kernel<true>);
}

template <int TN, bool TRACE>
int launch_pxc_t(const pxc::Params& p, const CUtensorMap& tmap, int ctas, size_t smem, cudaStream_t stream, bool pdl) {
  CUDA_OK(launch_k(pdl, pxc::pix_contract_kernel<TN, TRACE>, dim3(ctas), dim3(pxc::THREADS), smem, stream, p, tmap));
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
               long long* dbg_cycles = nullptr, bool pdl = false) {
  if (n_rb > 65535) return fail("too many theta blocks for one folded contraction launch");
  const PxcSchedule* sc = nullptr;
  if (pxc_schedule(pl, n_rb, sms, &sc)) return 1;
  pxc::Params p{};
  p.s = spec; p.pitch = pl.pitch; p.tile_kcol0 = pl.d_tile_kcol0; p.tile_nkb = pl.d_tile_nkb; p.col_idx = pl.d_col_idx;
  p.col_frac = pl.d_col_frac; p.wtab = pl.d_wtab; p.wlen = pl.wlen; p.sched = sc->d_sched; p.sched_cnt = sc->d_cnt;
  p.sched_slots = sc->slots; p.out = out; p.out_pitch = pl.out_pitch; p.ring = pl.ring; p.dbg_cycles = dbg_cycles;
  p.bbar = tsum ? pl.d_bbar : nullptr; p.tile_boff = pl.d_tile_boff; p.tsum = tsum;
  p.n_col_tiles = pl.n_col_tiles;
  p.bimg = pl.d_bimg; p.tile_bimg = pl.d_tile_bimg;
  // A loader: cp.async by two loader warps.  The TMA variant (one cp.async.bulk.tensor box per k-block) is compiled only
  // with -DBAYSAR_PXC_TMA.  Its data hazard is root-caused (round 2): the converters read a ring stage through the generic
  // proxy and the TMA refill writes it through the async proxy; the mbarrier hand-over alone does not order the two, so
  // a lagging lane could read the next cycle's data (single elements, a few launches in a thousand:
  // profiles/r2_pxc_tma_loader_unfenced_determinism.log).  With fence.proxy.async before the hand-over it is bit-
  // reproducible (profiles/r2_pxc_tma_loader_fenced_determinism.log) but the fence costs more than the TMA saves
  // (contraction 0.1076 ms against 0.1038 ms with cp.async), so cp.async stays the loader.
  p.use_tma = 0;
  p.sleep_epi = 0; p.sleep_build = 0;
  CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
#ifdef BAYSAR_PXC_TMA
  p.use_tma = 1;
  if (make_spec_tmap(spec, pl.pitch, (int64_t)n_rb * ifc::TILE_M, &tmap)) return 1;
#endif
  if (pl.tn == 64) {
    const size_t smem = pxc::smem_bytes<64>(pl.ring, pl.wlen);
    return dbg_cycles ? launch_pxc_t<64, true>(p, tmap, sc->ctas, smem, stream, pdl)
                      : launch_pxc_t<64, false>(p, tmap, sc->ctas, smem, stream, pdl);
  }
  const size_t smem = pxc::smem_bytes<128>(pl.ring, pl.wlen);
  return dbg_cycles ? launch_pxc_t<128, true>(p, tmap, sc->ctas, smem, stream, pdl)
                    : launch_pxc_t<128, false>(p, tmap, sc->ctas, smem, stream, pdl);
}

}  // namespace
