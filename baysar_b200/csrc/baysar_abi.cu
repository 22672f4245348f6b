// C ABI of the B200-native BaySAR posterior path (include/baysar_b200.h): model upload, workspace, launches.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

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
  double* logp_stage = nullptr;
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
  const size_t per_warp = los_smem_per_warp(L, (int)m->I[I_N_PARAMS]);
  if (per_warp > 200 * 1024) return fail("line-of-sight grid too long for the shared-memory LOS stage");
  static const int minb = getenv("BAYSAR_LOS_MINB") ? atoi(getenv("BAYSAR_LOS_MINB")) : 4;  // register cap experiment
  if (4 * per_warp <= 200 * 1024) {
    size_t smem = 4 * per_warp;
    auto kern = minb >= 6 ? los_stage_kernel<4, 6> : (minb == 5 ? los_stage_kernel<4, 5> : los_stage_kernel<4, 4>);
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
int launch_chord_kernel(baysar_model* m, const ChordArgs& a, dim3 grid, size_t smem, cudaStream_t stream) {
  constexpr int NT = 256;
  static const int cap = getenv("BAYSAR_CHORD_MINB") ? atoi(getenv("BAYSAR_CHORD_MINB")) : 5;
  const int by_smem = (int)((size_t)232448 / (smem + 1024));
  const int minb = by_smem < cap ? by_smem : cap;
  auto kern = minb >= 5 ? chord_stage_kernel<NT, MODE, 5> : (minb == 4 ? chord_stage_kernel<NT, MODE, 4> : chord_stage_kernel<NT, MODE, 3>);
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
                  uint32_t* status, cudaStream_t stream) {
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
      a.spec = m->spec; a.split_pitch = pl.pitch; a.split_pad = pl.pad_left;
      a.aux = m->aux;
      if (launch_chord_kernel<1>(m, a, dim3((unsigned)rows, 1), m->chord_smem[c], stream)) return 1;
      if (mark(m, ST_CHORD, stream)) return 1;
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
  cudaFree(m->theta_stage); cudaFree(m->logp_stage);
  cudaFree(m->spec); cudaFree(m->conv); cudaFree(m->conv_c); cudaFree(m->trapz_part); cudaFree(m->aux);
  for (float* t : m->ifc_taps_dev) cudaFree(t);
  for (uint32_t* t : m->ifc_cmask_dev) cudaFree(t);
  for (int* t : m->ifc_cbase_dev) cudaFree(t);
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
  if (launch_chords(m, theta, n, 0, (int)m->I[I_N_CHORDS], m->lines, comps, nullptr, nullptr, nullptr, st, stream))
    return 1;
  const int ncols = (int)(m->I[I_N_CHORDS] + m->I[I_N_PRIORS]);
  finalize_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(comps, ncols, n, logp, st);
  CUDA_OK(cudaGetLastError());
  m->last_launches += 1;
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
    cudaFree(m->theta_stage); cudaFree(m->logp_stage);
    m->theta_stage = nullptr; m->logp_stage = nullptr; m->stage_cap = 0;
    int64_t cap = n < 1024 ? 1024 : n;
    CUDA_OK(cudaMalloc(&m->theta_stage, sizeof(double) * cap * np));
    CUDA_OK(cudaMalloc(&m->logp_stage, sizeof(double) * cap));
    m->stage_cap = cap;
  }
  CUDA_OK(cudaMemcpyAsync(m->theta_stage, theta_host, sizeof(double) * n * np, cudaMemcpyHostToDevice, 0));
  if (baysar_eval_logp(m, m->theta_stage, n, m->logp_stage, nullptr, nullptr, nullptr)) return 1;
  CUDA_OK(cudaMemcpyAsync(logp_host, m->logp_stage, sizeof(double) * n, cudaMemcpyDeviceToHost, 0));
  if (status_host)
    CUDA_OK(cudaMemcpyAsync(status_host, m->status, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, 0));
  CUDA_OK(cudaStreamSynchronize(0));
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
