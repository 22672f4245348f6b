// C ABI of the B200-native BaySAR posterior path (include/baysar_b200.h): model upload, workspace, launches.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include <new>
#include <string>
#include <vector>

#include "../../include/baysar_b200.h"
#include "chord_stage.cuh"
#include "finalize.cuh"
#include "los_stage.cuh"

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
  size_t chord_smem_max = 0;
  // optional per-stage timing (bench.py roofline): 4 events per eval_logp call while profiling is on
  bool profiling = false;
  std::vector<cudaEvent_t> prof_events;
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
    for (int k = 0; k < 16; ++k) asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(v[k]) : "f"(v[k]));
  }
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 16; ++k) s += v[k];
  if (s == 12345.678f) out[threadIdx.x] = s * a;
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
  const size_t per_warp = los_smem_per_warp(L);
  if (per_warp > 200 * 1024) return fail("line-of-sight grid too long for the shared-memory LOS stage");
  if (4 * per_warp <= 200 * 1024) {
    size_t smem = 4 * per_warp;
    CUDA_OK(cudaFuncSetAttribute(los_stage_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    los_stage_kernel<4><<<(unsigned)((n + 3) / 4), 128, smem, stream>>>(m->dev, a);
  } else {
    size_t smem = per_warp;
    CUDA_OK(cudaFuncSetAttribute(los_stage_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    los_stage_kernel<1><<<(unsigned)n, 32, smem, stream>>>(m->dev, a);
  }
  CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_chords(baysar_model* m, const double* theta, int64_t n, int chord_begin, int n_chords,
                  const double* lines, double* comps, double* spectra, float* fine_pre, float* fine_post,
                  uint32_t* status, cudaStream_t stream) {
  ChordArgs a;
  a.theta = theta; a.n = n; a.chord_begin = chord_begin; a.lines = lines; a.components = comps;
  a.spectra = spectra; a.fine_pre = fine_pre; a.fine_post = fine_post; a.status = status;
  constexpr int NT = 256;
  CUDA_OK(cudaFuncSetAttribute(chord_stage_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)m->chord_smem_max));
  dim3 grid((unsigned)n, (unsigned)n_chords);
  chord_stage_kernel<NT><<<grid, NT, m->chord_smem_max, stream>>>(m->dev, a);
  CUDA_OK(cudaGetLastError());
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
  }
  if (smem_max > 227 * 1024) {
    baysar_model_destroy(m);
    return fail("fine grid too long for the shared-memory chord stage (needs " + std::to_string(smem_max) + " B)");
  }
  m->chord_smem_max = smem_max;
  *out = m;
  return 0;
}

void baysar_model_destroy(baysar_model_t* m) {
  if (!m) return;
  cudaSetDevice(m->device);
  cudaFree(m->blob_dev); cudaFree(m->lines); cudaFree(m->comps); cudaFree(m->hdr); cudaFree(m->status);
  cudaFree(m->theta_stage); cudaFree(m->logp_stage);
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
    case 9: return 3;
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
  auto mark = [&]() -> int {
    if (!m->profiling) return 0;
    cudaEvent_t ev;
    CUDA_OK(cudaEventCreate(&ev));
    CUDA_OK(cudaEventRecord(ev, stream));
    m->prof_events.push_back(ev);
    return 0;
  };
  if (mark()) return 1;
  if (launch_los(m, theta, n, m->lines, comps, nullptr, st, stream)) return 1;
  if (mark()) return 1;
  if (launch_chords(m, theta, n, 0, (int)m->I[I_N_CHORDS], m->lines, comps, nullptr, nullptr, nullptr, st, stream))
    return 1;
  if (mark()) return 1;
  const int ncols = (int)(m->I[I_N_CHORDS] + m->I[I_N_PRIORS]);
  finalize_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(comps, ncols, n, logp, st);
  CUDA_OK(cudaGetLastError());
  if (mark()) return 1;
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
  m->profiling = on != 0;
  return 0;
}

int baysar_model_get_profile(baysar_model_t* m, double* stage_ms, int64_t* calls) {
  if (!m || !stage_ms || !calls) return fail("null argument");
  CUDA_OK(cudaSetDevice(m->device));
  stage_ms[0] = stage_ms[1] = stage_ms[2] = 0.0;
  *calls = (int64_t)(m->prof_events.size() / 4);
  for (size_t c = 0; c + 3 < m->prof_events.size(); c += 4) {
    CUDA_OK(cudaEventSynchronize(m->prof_events[c + 3]));
    for (int k = 0; k < 3; ++k) {
      float ms = 0.f;
      CUDA_OK(cudaEventElapsedTime(&ms, m->prof_events[c + k], m->prof_events[c + k + 1]));
      stage_ms[k] += ms;
    }
  }
  for (cudaEvent_t ev : m->prof_events) cudaEventDestroy(ev);
  m->prof_events.clear();
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
    else microbench_dfma_kernel<<<blocks, threads>>>(reinterpret_cast<double*>(out), iters, 1.0001);
    CUDA_OK(cudaEventRecord(e1, 0));
    CUDA_OK(cudaEventSynchronize(e1));
    float ms = 0.f;
    CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
    // ops per thread per iteration: 16 independent chains x 1 instruction
    double ops = (double)blocks * threads * (double)iters * 16.0;
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
