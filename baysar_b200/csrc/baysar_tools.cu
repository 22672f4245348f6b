// Measurement and unit-test entry points (include/baysar_b200_tools.h): peak-rate micro-benchmarks for the roofline
// denominators, the two tcgen05 contractions on caller-supplied spectra, the host-only plan check and the MMA issue
// probe.  Built into libbaysar_b200_tools.so; nothing here is on the posterior path.
#include "../../include/baysar_b200_tools.h"
#include "host_plans.h"
#include "mma_probe.cuh"

namespace {

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

}  // namespace

extern "C" {

const char* baysar_tools_last_error(void) { return g_last_error.c_str(); }

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
  if (contraction_kernel_attributes()) return 1;
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
  if (upload_pxc_plan(pl) || (!getenv("BAYSAR_PXC_BUILDERS") && build_pxc_b_image(pl))) { free_pxc_plan(pl); return 1; }
  if (contraction_kernel_attributes()) { free_pxc_plan(pl); return 1; }
  float* spec = nullptr;
  if (tsum_dev && tsum_capacity < n_rows * pl.n_col_tiles * pxc::TSUM_SLOTS) return fail("tsum buffer too small");
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
    const char* names[15] = {"conv.wait_landed", "conv.wait_empty", "conv.total", "load.wait_consumed", "build.wait_bempty",
                             "build.total", "epi.wait_acc_full", "epi.drain", "mma.wait_full", "mma.wait_bfull",
                             "mma.wait_acc_free", "mma.total", "mma.k_blocks", "cta.prologue", "cta.total"};
    const int ctas = (int)std::min<int64_t>(sms, (n_rows / ifc::TILE_M) * pl.n_col_tiles);
    for (int k = 0; k < 15; ++k) {
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
