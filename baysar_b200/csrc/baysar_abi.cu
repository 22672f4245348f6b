// C ABI of the B200-native BaySAR posterior path (include/baysar_b200.h): model upload, workspace, launches.
#include "../../include/baysar_b200.h"
#include "chord_stage.cuh"
#include "ensemble.cuh"
#include "finalize.cuh"
#include "host_plans.h"
#include "los_stage.cuh"
#include "los_tables.h"

struct baysar_model {
  int device = 0;
  unsigned char* blob_dev = nullptr;
  unsigned char* los_tables_dev = nullptr;  // derived tables of the LOS stage (los_tables.h): doubles, then bucket tables
  size_t blob_bytes = 0;
  std::vector<unsigned char> blob_host;  // header + small tables are read on the host for launch configuration
  ModelDev dev;
  const int64_t* I = nullptr;   // host view
  const int* ch_i = nullptr;    // host view
  const double* ch_d = nullptr;
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
  // layout (pitch, pad, fine points) of the spectra rows as the last spectral-stage launch left them, and the rows it
  // covered: a launch with the same layout finds its zero padding in place
  int64_t spec_pitch_last = -1; int spec_pad_last = -1, spec_f_last = -1; int64_t spec_rows_last = 0;
  baysar_options_t opt{};  // tuning options given at creation (0 = default everywhere)
  // optional per-stage timing (bench.py roofline): events tagged with the stage that ends at them.  The events are a pool
  // that only grows (first profiled calls) and is reused after every baysar_model_get_profile
  bool profiling = false;
  std::vector<cudaEvent_t> prof_events;
  std::vector<int> prof_slots;  // stage tags of the prof_used recorded events
  size_t prof_used = 0;
  // host-buffer entry point: its own non-blocking stream (the legacy default stream would serialise with every other
  // stream of the process)
  cudaStream_t host_stream = nullptr;
  cudaStream_t copy_stream = nullptr;       // baysar_eval_logp_host: theta chunks are copied here while earlier chunks are evaluated
  std::vector<cudaEvent_t> copy_done;       // one event per chunk in flight
  // one workspace per model: every evaluation waits for the previous one (whatever stream it ran on) before it starts
  cudaEvent_t last_done = nullptr;
};

namespace {

struct SectionTable {
  const int64_t* hdr;
  const unsigned char* base;
  const void* ptr(int s) const { return base + hdr[3 + 2 * s]; }
  int64_t count(int s) const { return hdr[3 + 2 * s + 1]; }
};

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
  const bool pdl = !m->opt.serial_launches;
  const int L = (int)m->I[I_L];
  const size_t per_warp = los_smem_per_warp(L, (int)m->I[I_N_PARAMS], (int)m->I[I_N_SPECIES]);
  if (per_warp > 200 * 1024) return fail("line-of-sight grid too long for the shared-memory LOS stage");
  // Register budget by batch size.  The kernel is latency-bound at 16 warps per SM (4 CTAs, 128 registers, no spills):
  // the best choice for many waves.  A batch of a few thousand thetas, though, is one or two waves, and then fitting
  // every CTA into ONE wave beats the spills of a tighter budget (4096 thetas: 1024 CTAs = 1.7 waves of 592 at 0.080 ms,
  // one wave of 7 CTAs per SM at 72 registers at 0.068 ms).  options.los_min_blocks overrides.
  // (many waves: 5 CTAs per SM at 96 registers when the per-point cache leaves room for them -- multi-species chord
  // 1.88 ms at 4, 1.76 at 5, 1.82 at 6 with its spills)
  int minb = (size_t)232448 / (4 * per_warp + 1024) >= 5 ? 5 : 4;
  if (m->opt.los_min_blocks > 0) {
    minb = m->opt.los_min_blocks;
  } else {
    const int64_t ctas = (n + 3) / 4;
    if (ctas > (int64_t)minb * m->sm_count)
      for (int b = minb + 1; b <= 7; ++b)
        if (ctas <= (int64_t)b * m->sm_count) { minb = b; break; }
  }
  // Long lines of sight: the per-point cache (84 B per point and theta) starves the warp-per-theta form of occupancy, so
  // from a few hundred points on TW warps share ONE theta (one CTA per theta, the points spread over its threads).
  // Measured LOS time per 32768 thetas (20 lines, tools/los_team_probe.py): L = 100: 1.92 ms warp-per-theta, 2.66 team
  // of four at 128 registers, 2.13 at 96; 160: 2.91 / 2.87; 200: 4.35 / 3.32; 300: 8.88 / 4.66; 500: 18.0 / 7.87;
  // 700: 10.8 with eight warps, two CTAs per SM (12.4 with four); 1000: 15.2 (23.0).  Default: warp-per-theta up to
  // 160 points, then the team size (four or eight warps) and register budget that keep the most warps on an SM.
  // options.los_team_points (teams from that length on), los_team_warps (2, 4, 8) and los_min_blocks override.
  const size_t smem_sm = 227 * 1024;
  int tw = m->opt.los_team_warps;
  if (tw != 2 && tw != 4 && tw != 8) {
    tw = 0;
    const int team_l = m->opt.los_team_points;
    if ((team_l > 0 && L >= team_l) || 4 * per_warp > 200 * 1024 || L > 160) {
      const int n4 = (int)std::min<size_t>(smem_sm / (per_warp + los_team_xch_bytes(4) + 1024), 5);
      const int n8 = (int)std::min<size_t>(smem_sm / (per_warp + los_team_xch_bytes(8) + 1024), 2);
      tw = 8 * n8 > 4 * n4 ? 8 : 4;
    }
  }
  if (tw && per_warp + los_team_xch_bytes(tw) > 200 * 1024) tw = 0;
  if (tw) {
    const size_t smem = per_warp + los_team_xch_bytes(tw);
    const int fit = (int)(smem_sm / (smem + 1024));
    const int want = m->opt.los_min_blocks > 0 ? m->opt.los_min_blocks : fit;
    if (tw == 2) {
      if (want >= 16) CUDA_OK(launch_k(pdl, los_stage_kernel<2, 16, 1, 2>, dim3((unsigned)n), dim3(64), smem, stream, m->dev, a));
      else CUDA_OK(launch_k(pdl, los_stage_kernel<2, 12, 1, 2>, dim3((unsigned)n), dim3(64), smem, stream, m->dev, a));
    } else if (tw == 4) {
      if (want >= 5) CUDA_OK(launch_k(pdl, los_stage_kernel<4, 5, 1, 4>, dim3((unsigned)n), dim3(128), smem, stream, m->dev, a));
      else if (want == 4) CUDA_OK(launch_k(pdl, los_stage_kernel<4, 4, 1, 4>, dim3((unsigned)n), dim3(128), smem, stream, m->dev, a));
      else CUDA_OK(launch_k(pdl, los_stage_kernel<4, 1, 1, 4>, dim3((unsigned)n), dim3(128), smem, stream, m->dev, a));
    } else {
      if (want >= 2) CUDA_OK(launch_k(pdl, los_stage_kernel<8, 2, 1, 8>, dim3((unsigned)n), dim3(256), smem, stream, m->dev, a));
      else CUDA_OK(launch_k(pdl, los_stage_kernel<8, 1, 1, 8>, dim3((unsigned)n), dim3(256), smem, stream, m->dev, a));
    }
  } else if (4 * per_warp <= 200 * 1024) {
    const size_t smem = 4 * per_warp;
    auto kern = minb >= 6 ? los_stage_kernel<4, 6> : (minb == 5 ? los_stage_kernel<4, 5> : los_stage_kernel<4, 4>);
    if (minb == 7) kern = los_stage_kernel<4, 7>;
    if (minb >= 8) kern = los_stage_kernel<4, 8>;
    CUDA_OK(launch_k(pdl, kern, dim3((unsigned)((n + 3) / 4)), dim3(128), smem, stream, m->dev, a));
  } else {
    CUDA_OK(launch_k(pdl, los_stage_kernel<1>, dim3((unsigned)n), dim3(32), per_warp, stream, m->dev, a));
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
  // the memset runs on the legacy stream, the evaluation on a non-blocking one: without this (allocation-time only)
  // synchronisation nothing orders the zero fill before the first spectral-stage launch that writes the same rows
  CUDA_OK(cudaDeviceSynchronize());
  m->ifc_rows = rows;
  m->spec_pitch_last = -1; m->spec_rows_last = 0;
  return 0;
}

// Launch the chord kernel with the register budget that matches how many CTAs the shared-memory footprint lets an SM hold
template <int MODE, int TAPS>
int launch_chord_kernel_t(baysar_model* m, const ChordArgs& a, dim3 grid, size_t smem, cudaStream_t stream) {
  // 128 threads per (theta, chord) on short fine grids: with a few thousand points 256 threads have a dozen points each
  // and the per-thread set-up and the barriers dominate (multi-species chord, 3000 points: 2.84 -> 2.62 ms); long grids
  // keep 256 (Balmer demo, 8000 points: 0.119 vs 0.128 ms).  options.chord_threads overrides.
  const bool pdl = !m->opt.serial_launches;
  const bool nt128 = m->opt.chord_threads == 128 || (m->opt.chord_threads == 0 && m->I[I_F_MAX] <= 4096);
  if (nt128) {
    // register budget by how many CTAs the shared-memory footprint lets an SM hold (6 by default, options.chord_min_blocks
    // raises it to 7 or 8 where the footprint allows)
    const int by_smem128 = (int)((size_t)232448 / (smem + 1024));
    const int want128 = m->opt.chord_min_blocks > 6 ? m->opt.chord_min_blocks : 6;
    const int minb128 = by_smem128 < want128 ? by_smem128 : want128;
    auto kern128 = minb128 >= 8 ? chord_stage_kernel<128, MODE, 8, TAPS>
                                : (minb128 == 7 ? chord_stage_kernel<128, MODE, 7, TAPS> : chord_stage_kernel<128, MODE, 6, TAPS>);
    CUDA_OK(launch_k(pdl, kern128, grid, dim3(128), smem, stream, m->dev, a));
    return 0;
  }
  constexpr int NT = 256;
  const int cap = m->opt.chord_min_blocks > 0 ? m->opt.chord_min_blocks : 5;
  const int by_smem = (int)((size_t)232448 / (smem + 1024));
  const int minb = by_smem < cap ? by_smem : cap;
  auto kern = minb >= 6 ? chord_stage_kernel<NT, MODE, 6, TAPS>
                        : (minb == 5 ? chord_stage_kernel<NT, MODE, 5, TAPS>
                                     : (minb == 4 ? chord_stage_kernel<NT, MODE, 4, TAPS> : chord_stage_kernel<NT, MODE, 3, TAPS>));
  CUDA_OK(launch_k(pdl, kern, grid, dim3(NT), smem, stream, m->dev, a));
  return 0;
}

// Launch the chord kernel with the register budget that matches how many CTAs the shared-memory footprint lets an SM
// hold, in the build that matches what the host guarantees about the chord's Doppler kernels (chord_stage.cuh TAPS)
template <int MODE>
int launch_chord_kernel(baysar_model* m, const ChordArgs& a, dim3 grid, size_t smem, cudaStream_t stream, int chord) {
  const int* ch = m->ch_i + chord * CH_I_COUNT;
  if (ch[CH_B_OPTIONAL] || ch[CH_NARROW]) {
    ChordArgs b = a;
    b.kd_max = 0;  // no Doppler tap table in these builds (the shared-memory size was computed without it)
    if (MODE == 1 && a.skip_edge64 && ch[CH_B_OPTIONAL]) {
      // Spectral stage of the folded contraction with single-tap lines only: nothing is convolved in this kernel, so the
      // spectrum buffer needs neither haloes nor the instrument-tap table (Balmer demo: 38.2 -> 34.6 KB, a sixth CTA fits)
      b.ch_i[CH_HALO] = 0;
      b.ch_i[CH_TAPS_MAX] = 0;
      smem = chord_smem_layout(ch[CH_F], 0, 0, 0, 0, ch[CH_N_LINES]).total;
    }
    return ch[CH_B_OPTIONAL] ? launch_chord_kernel_t<MODE, 0>(m, b, grid, smem, stream)
                             : launch_chord_kernel_t<MODE, 1>(m, b, grid, smem, stream);
  }
  return launch_chord_kernel_t<MODE, 2>(m, a, grid, smem, stream);
}

// stage tags for the optional event log
enum { ST_BEGIN = -1, ST_LOS = 0, ST_CHORD = 1, ST_CONTRACT = 2, ST_PIXEL = 3, ST_FINAL = 4, ST_COUNT = 5 };

int mark(baysar_model* m, int slot, cudaStream_t stream) {
  if (!m->profiling) return 0;
  if (m->prof_used == m->prof_events.size()) {  // pool growth: first profiled calls only
    cudaEvent_t ev;
    CUDA_OK(cudaEventCreate(&ev));
    m->prof_events.push_back(ev);
  }
  CUDA_OK(cudaEventRecord(m->prof_events[m->prof_used], stream));
  if (m->prof_slots.size() <= m->prof_used) m->prof_slots.resize(m->prof_used + 1);
  m->prof_slots[m->prof_used++] = slot;
  return 0;
}

// Opt-in shared-memory limits of every kernel variant the launchers can pick: once per model creation
int set_kernel_attributes() {
  if (allow_max_smem(los_stage_kernel<2, 12, 1, 2>) || allow_max_smem(los_stage_kernel<2, 16, 1, 2>) ||
      allow_max_smem(los_stage_kernel<4, 4, 1, 4>) || allow_max_smem(los_stage_kernel<4, 5, 1, 4>) ||
      allow_max_smem(los_stage_kernel<8, 2, 1, 8>) ||
      allow_max_smem(los_stage_kernel<8, 1, 1, 8>) || allow_max_smem(los_stage_kernel<4, 1, 1, 4>) ||
      allow_max_smem(los_stage_kernel<4, 4>) || allow_max_smem(los_stage_kernel<4, 5>) ||
      allow_max_smem(los_stage_kernel<4, 6>) || allow_max_smem(los_stage_kernel<4, 7>) ||
      allow_max_smem(los_stage_kernel<4, 8>) || allow_max_smem(los_stage_kernel<1>))
    return 1;
  if (allow_max_smem(chord_stage_kernel<256, 0, 3, 2>) || allow_max_smem(chord_stage_kernel<256, 0, 4, 2>) ||
      allow_max_smem(chord_stage_kernel<256, 0, 5, 2>) || allow_max_smem(chord_stage_kernel<256, 1, 3, 2>) ||
      allow_max_smem(chord_stage_kernel<256, 1, 4, 2>) || allow_max_smem(chord_stage_kernel<256, 1, 5, 2>) ||
      allow_max_smem(chord_stage_kernel<256, 0, 3, 1>) || allow_max_smem(chord_stage_kernel<256, 0, 4, 1>) ||
      allow_max_smem(chord_stage_kernel<256, 0, 5, 1>) || allow_max_smem(chord_stage_kernel<256, 1, 3, 1>) ||
      allow_max_smem(chord_stage_kernel<256, 1, 4, 1>) || allow_max_smem(chord_stage_kernel<256, 1, 5, 1>) ||
      allow_max_smem(chord_stage_kernel<256, 0, 3, 0>) || allow_max_smem(chord_stage_kernel<256, 0, 4, 0>) ||
      allow_max_smem(chord_stage_kernel<256, 0, 5, 0>) || allow_max_smem(chord_stage_kernel<256, 1, 3, 0>) ||
      allow_max_smem(chord_stage_kernel<256, 1, 4, 0>) || allow_max_smem(chord_stage_kernel<256, 1, 5, 0>) ||
      allow_max_smem(chord_stage_kernel<256, 0, 6, 2>) || allow_max_smem(chord_stage_kernel<256, 1, 6, 2>) ||
      allow_max_smem(chord_stage_kernel<256, 0, 6, 1>) || allow_max_smem(chord_stage_kernel<256, 1, 6, 1>) ||
      allow_max_smem(chord_stage_kernel<256, 0, 6, 0>) || allow_max_smem(chord_stage_kernel<256, 1, 6, 0>) ||
      allow_max_smem(chord_stage_kernel<128, 0, 6, 2>) || allow_max_smem(chord_stage_kernel<128, 1, 6, 2>) ||
      allow_max_smem(chord_stage_kernel<128, 0, 6, 1>) || allow_max_smem(chord_stage_kernel<128, 1, 6, 1>) ||
      allow_max_smem(chord_stage_kernel<128, 0, 6, 0>) || allow_max_smem(chord_stage_kernel<128, 1, 6, 0>) ||
      allow_max_smem(chord_stage_kernel<128, 0, 7, 2>) || allow_max_smem(chord_stage_kernel<128, 1, 7, 2>) ||
      allow_max_smem(chord_stage_kernel<128, 0, 7, 1>) || allow_max_smem(chord_stage_kernel<128, 1, 7, 1>) ||
      allow_max_smem(chord_stage_kernel<128, 0, 7, 0>) || allow_max_smem(chord_stage_kernel<128, 1, 7, 0>) ||
      allow_max_smem(chord_stage_kernel<128, 0, 8, 2>) || allow_max_smem(chord_stage_kernel<128, 1, 8, 2>) ||
      allow_max_smem(chord_stage_kernel<128, 0, 8, 1>) || allow_max_smem(chord_stage_kernel<128, 1, 8, 1>) ||
      allow_max_smem(chord_stage_kernel<128, 0, 8, 0>) || allow_max_smem(chord_stage_kernel<128, 1, 8, 0>))
    return 1;
  return allow_max_smem(pixel_fold_kernel<128>) || contraction_kernel_attributes();
}

int launch_chords(baysar_model* m, const double* theta, int64_t n, int chord_begin, int n_chords,
                  const double* lines, double* comps, double* spectra, float* fine_pre, float* fine_post,
                  uint32_t* status, cudaStream_t stream, double* logp = nullptr, bool* finalized = nullptr) {
  const int n_params = (int)m->I[I_N_PARAMS];
  const int ncols = (int)(m->I[I_N_CHORDS] + m->I[I_N_PRIORS]);
  bool any = false;
  for (int c = chord_begin; c < chord_begin + n_chords; ++c) any = any || !m->ifc[c].taps_rev.empty();
  if (any && ensure_ifc_workspace(m, n)) return 1;
  for (int c = chord_begin; c < chord_begin + n_chords; ++c) {
    const IfcPlan& pl = m->ifc[c];
    const int P = m->ch_i[c * CH_I_COUNT + CH_P], F = m->ch_i[c * CH_I_COUNT + CH_F];
    if (pl.taps_rev.empty()) {
      ChordArgs a{};
      chord_args_set_chord(a, m->I, m->ch_i, m->ch_d, m->cl_i, c);
      a.theta = theta; a.n = n; a.lines = lines; a.components = comps;
      a.spectra = spectra; a.fine_pre = fine_pre; a.fine_post = fine_post; a.status = status;
      if (launch_chord_kernel<0>(m, a, dim3((unsigned)n, 1), m->chord_smem[c], stream, c)) return 1;
      m->last_launches += 1;
      if (mark(m, ST_CHORD, stream)) return 1;
      continue;
    }
    for (int64_t r0 = 0; r0 < n; r0 += m->ifc_rows) {
      const int64_t rows = (n - r0) < m->ifc_rows ? (n - r0) : m->ifc_rows;
      const int64_t rows128 = (rows + ifc::TILE_M - 1) / ifc::TILE_M * ifc::TILE_M;
      // K2a: spectrum -> TF32-split rows + per-theta scalars
      ChordArgs a{};
      chord_args_set_chord(a, m->I, m->ch_i, m->ch_d, m->cl_i, c);
      a.theta = theta + r0 * n_params; a.n = rows;
      a.lines = lines + r0 * m->I[I_N_ULINES] * BAYSAR_LINE_STRIDE;
      a.components = comps + r0 * ncols; a.spectra = nullptr;
      a.fine_pre = fine_pre ? fine_pre + r0 * F : nullptr; a.fine_post = nullptr; a.status = status + r0;
      PxcPlan& fx = m->pxc[c];
      const bool folded = fx.on && !fine_post;  // the fine-grid diagnostic needs every convolved column
      a.spec = m->spec; a.split_pitch = folded ? fx.pitch : pl.pitch; a.split_pad = folded ? fx.pad_left : pl.pad_left;
      a.aux = m->aux;
      // the folded pixel stage reads the float64 end outputs only when a pixel extrapolates beyond the fine grid
      a.skip_edge64 = folded && fx.no_extrapolation ? 1 : 0;
      a.keep_padding = (m->spec_pitch_last == a.split_pitch && m->spec_pad_last == a.split_pad && m->spec_f_last == F &&
                        m->spec_rows_last >= rows) ? 1 : 0;
      if (!a.keep_padding) { m->spec_pitch_last = a.split_pitch; m->spec_pad_last = a.split_pad; m->spec_f_last = F; m->spec_rows_last = rows; }
      if (launch_chord_kernel<1>(m, a, dim3((unsigned)rows, 1), m->chord_smem[c], stream, c)) return 1;
      if (mark(m, ST_CHORD, stream)) return 1;
      if (folded) {
        // K2b': instrument function and wavelength map as one tcgen05 contraction; K2c': area identity + likelihood
        if (launch_pxc(fx, m->spec, (int)(rows128 / ifc::TILE_M), m->conv, m->trapz_part, m->sm_count, stream, nullptr, !m->opt.serial_launches)) return 1;
        if (mark(m, ST_CONTRACT, stream)) return 1;
        PixelFoldArgs pf{};
        pf.theta = theta + r0 * n_params; pf.n = rows; pf.chord = c; pf.cp = m->conv; pf.cp_pitch = fx.out_pitch;
        pf.pix_col = fx.d_pix_col; pf.col_left0 = fx.col_left0; pf.x_l = fx.x_l; pf.col_right0 = fx.col_right0; pf.x_r = fx.x_r;
        pf.spec = m->spec; pf.spec_pitch = fx.pitch; pf.spec_pad = fx.pad_left; pf.dtab = fx.d_dtab; pf.n_dl = fx.n_dl;
        pf.n_dr = fx.n_dr; pf.wsum = fx.wsum; pf.inv_err = fx.d_inv_err; pf.tsum = m->trapz_part;
        pf.use_edge64 = fx.no_extrapolation ? 0 : 1; pf.n_col_tiles = fx.n_col_tiles; pf.n_fold_tiles = fx.n_fold_tiles; pf.tile_n = fx.tn; pf.aux = m->aux; pf.components = comps + r0 * ncols;
        pf.spectra = spectra ? spectra + r0 * P : nullptr; pf.status = status + r0;
        {
          // a single-chord model gets its logP from the same kernel (no finalize launch)
          const bool fuse = logp && !pf.spectra && pf.inv_err && m->I[I_N_CHORDS] == 1;
          pf.logp = fuse ? logp + r0 : nullptr;
          if (fuse && finalized) *finalized = true;
          const size_t row_bytes = sizeof(float) * (size_t)fx.out_pitch;
          if (row_bytes > 160 * 1024) return fail("contraction row too long for the pixel stage's shared-memory copy");
          CUDA_OK(launch_k(!m->opt.serial_launches, pixel_fold_kernel<128>, dim3((unsigned)rows), dim3(128), row_bytes, stream, m->dev, pf));
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

// Structural validation of a descriptor blob before anything indexes it (a truncated or stale blob must fail here, not
// as an out-of-bounds read on the host or the device).
int validate_blob(const void* desc_blob, size_t nbytes) {
  if (!desc_blob || nbytes < (size_t)8 * (3 + 2 * SEC_COUNT)) return fail("bad descriptor blob");
  const int64_t* hdr = static_cast<const int64_t*>(desc_blob);
  if (hdr[0] != BAYSAR_BLOB_MAGIC) return fail("descriptor magic mismatch");
  if (hdr[1] != BAYSAR_BLOB_VERSION) return fail("descriptor version mismatch (rebuild the model with this library)");
  if (hdr[2] != SEC_COUNT) return fail("descriptor section count mismatch");
  auto elem = [](int sec) -> size_t {
    switch (sec) {
      case SEC_SP_I: case SEC_ELEM_DENS_IDX: case SEC_UL_I: case SEC_PR_I: case SEC_CSO_OFF: case SEC_CSO_UL:
      case SEC_CH_I: case SEC_CL_I: case SEC_PIX_IDX: case SEC_MESH_LO: case SEC_PR_AUX: case SEC_GR_I: case SEC_GR_UL: return 4;
      default: return 8;
    }
  };
  for (int sc = 0; sc < SEC_COUNT; ++sc) {
    const int64_t off = hdr[3 + 2 * sc], cnt = hdr[3 + 2 * sc + 1];
    if (off < 0 || cnt < 0 || (off & 15) != 0 || (size_t)off > nbytes || (size_t)cnt > (nbytes - (size_t)off) / elem(sc))
      return fail("descriptor section " + std::to_string(sc) + " out of range or misaligned");
  }
  SectionTable t{hdr, static_cast<const unsigned char*>(desc_blob)};
  if (t.count(SEC_I) != I_COUNT) return fail("descriptor I_COUNT mismatch");
  const int64_t* I = static_cast<const int64_t*>(t.ptr(SEC_I));
  const int64_t np = I[I_N_PARAMS], L = I[I_L], nc = I[I_N_CHORDS], nsp = I[I_N_SPECIES], nul = I[I_N_ULINES],
                npr = I[I_N_PRIORS], ncl = I[I_N_CLINES];
  if (np < 1 || L < 2 || nc < 0 || nsp < 0 || nul < 0 || npr < 0 || ncl < 0) return fail("descriptor scalar out of range");
  auto need = [&](int sec, int64_t cnt, const char* what) {
    return t.count(sec) < cnt ? fail(std::string("descriptor section too short: ") + what) : 0;
  };
  if (need(SEC_LOS_X, L, "LOS_X") || need(SEC_LOS_W, L, "LOS_W") || need(SEC_PROF_D, PROF_D_COUNT, "PROF_D") ||
      need(SEC_SP_I, nsp * SP_I_COUNT, "SP_I") || need(SEC_SP_D, nsp * SP_D_COUNT, "SP_D") ||
      need(SEC_ELEM_DENS_IDX, I[I_N_ELEM], "ELEM_DENS_IDX") || need(SEC_UL_I, nul * UL_I_COUNT, "UL_I") ||
      need(SEC_UL_D, nul * UL_D_COUNT, "UL_D") || need(SEC_PR_I, npr * PR_I_COUNT, "PR_I") ||
      need(SEC_PR_D, npr * PR_D_COUNT, "PR_D") || need(SEC_CSO_OFF, I[I_N_ELEM] + 1, "CSO_OFF") ||
      need(SEC_CH_I, nc * CH_I_COUNT, "CH_I") || need(SEC_CH_D, nc * CH_D_COUNT, "CH_D") ||
      need(SEC_CL_I, ncl * CL_I_COUNT, "CL_I") || need(SEC_CL_D, ncl * CL_D_COUNT, "CL_D") ||
      need(SEC_A11_NE, I[I_A11_NNE], "A11_NE") || need(SEC_A11_TE, I[I_A11_NTE], "A11_TE") ||
      need(SEC_A11_SCD, I[I_A11_NNE] * I[I_A11_NTE], "A11_SCD") ||
      need(SEC_A11_ACD, I[I_RECOMBINING] ? I[I_A11_NNE] * I[I_A11_NTE] : 0, "A11_ACD") || need(SEC_H_NE, I[I_H_NNE], "H_NE") ||
      need(SEC_H_TE, I[I_H_NTE], "H_TE") || need(SEC_H_TAB, I[I_N_HTAB] * I[I_H_NNE] * I[I_H_NTE], "H_TAB") ||
      need(SEC_TAU_GRID, I[I_N_TAU], "TAU_GRID") || need(SEC_SPL_TX, I[I_SPL_NX], "SPL_TX") ||
      need(SEC_SPL_TY, I[I_SPL_NY], "SPL_TY") ||
      need(SEC_GR_I, I[I_N_GROUPS] * GR_I_COUNT, "GR_I") ||
      need(SEC_MESH_X, I[I_MESH_NE_N] + I[I_MESH_TE_N], "MESH_X") ||
      need(SEC_MESH_LO, I[I_PROFILE_KIND] == 2 ? 2 * L : 0, "MESH_LO") || need(SEC_PIX_FRAC, t.count(SEC_PIX_IDX), "PIX_FRAC") ||
      need(SEC_ERR, t.count(SEC_Y), "ERR") || need(SEC_PIX_IDX, t.count(SEC_Y), "PIX_IDX"))
    return 1;
  {
    // ADAS line groups: members are unique-line ids, coefficient blocks lie inside SEC_SPL_C (32-byte entries, read with
    // 256-bit loads: the section must start on a 32-byte boundary)
    const int* gr_i = static_cast<const int*>(t.ptr(SEC_GR_I));
    const int* gr_ul = static_cast<const int*>(t.ptr(SEC_GR_UL));
    const int* ul_i = static_cast<const int*>(t.ptr(SEC_UL_I));
    const int64_t cells = (I[I_SPL_NX] - 4) * (I[I_SPL_NY] - 4), pairs = I[I_N_TAU] - 1;
    if (I[I_N_GROUPS] < 0 || (I[I_N_GROUPS] > 0 && (cells < 1 || pairs < 1 || (hdr[3 + 2 * SEC_SPL_C] & 31) != 0)))
      return fail("descriptor ADAS line groups inconsistent with the spline tables");
    for (int64_t g = 0; g < I[I_N_GROUPS]; ++g) {
      const int* gr = gr_i + g * GR_I_COUNT;
      if (gr[GR_SPECIES] < 0 || gr[GR_SPECIES] >= nsp || gr[GR_NK] < 1 || gr[GR_FIRST] < 0 ||
          (int64_t)gr[GR_FIRST] + gr[GR_NK] > t.count(SEC_GR_UL) || gr[GR_COEF_OFF] < 0 ||
          ((int64_t)gr[GR_COEF_OFF] + pairs * cells * gr[GR_NK]) * 4 > t.count(SEC_SPL_C))
        return fail("descriptor ADAS line group " + std::to_string(g) + " out of range");
      for (int k = 0; k < gr[GR_NK]; ++k) {
        const int u = gr_ul[gr[GR_FIRST] + k];
        if (u < 0 || u >= nul || ul_i[u * UL_I_COUNT + UL_KIND] != 1)
          return fail("descriptor ADAS line group " + std::to_string(g) + " lists a line that is not an ADAS line");
      }
    }
  }
  const int* ch_i = static_cast<const int*>(t.ptr(SEC_CH_I));
  const int* cl_i = static_cast<const int*>(t.ptr(SEC_CL_I));
  const int64_t n_comps = t.count(SEC_COMP_D) / 2;
  auto theta_ok = [&](int idx) { return idx >= -1 && idx < np; };
  for (int64_t c = 0; c < nc; ++c) {
    const int* ch = ch_i + c * CH_I_COUNT;
    if (ch[CH_P] < 1 || ch[CH_F] < 8 || ch[CH_KI] < 1 || ch[CH_N_LINES] < 0 || ch[CH_LINE_OFF] < 0 ||
        (int64_t)ch[CH_LINE_OFF] + ch[CH_N_LINES] > ncl || ch[CH_PIX_OFF] < 0 ||
        (int64_t)ch[CH_PIX_OFF] + ch[CH_P] > t.count(SEC_Y) || ch[CH_IF_OFF] < 0 ||
        (int64_t)ch[CH_IF_OFF] + ch[CH_KI] > t.count(SEC_IF) || ch[CH_IF_KLO] < 0 || ch[CH_IF_KLO] >= ch[CH_IF_KHI] ||
        ch[CH_IF_KHI] > ch[CH_KI] || ch[CH_HALO] < 0 || ch[CH_HALO] % 4 != 0 || ch[CH_TAPS_MAX] < 0 ||
        ch[CH_BG_IDX] < 0 || !theta_ok(ch[CH_BG_IDX]) || !theta_ok(ch[CH_CAL_IDX]) || !theta_ok(ch[CH_CALWAVE_IDX] + (ch[CH_CALWAVE_IDX] >= 0)) ||
        !theta_ok(ch[CH_CALINT_IDX]))
      return fail("descriptor chord record " + std::to_string(c) + " inconsistent with the section sizes");
  }
  for (int64_t k = 0; k < ncl; ++k) {
    const int* cl = cl_i + k * CL_I_COUNT;
    if (cl[CL_KIND] < 0 || cl[CL_KIND] > 2 || cl[CL_ULINE] < 0 || cl[CL_ULINE] >= nul || cl[CL_NCOMP] < 0 ||
        cl[CL_COMP_OFF] < 0 || (int64_t)cl[CL_COMP_OFF] + cl[CL_NCOMP] > n_comps || cl[CL_KD] < 0 || cl[CL_SPECIES] >= nsp)
      return fail("descriptor line record " + std::to_string(k) + " inconsistent with the section sizes");
  }
  return 0;
}

}  // namespace

extern "C" {

const char* baysar_last_error(void) { return g_last_error.c_str(); }
const char* baysar_version(void) { return "baysar_b200 0.1 (sm_100a)"; }

int baysar_model_create(const void* desc_blob, size_t nbytes, int device, baysar_model_t** out) {
  return baysar_model_create_ex(desc_blob, nbytes, device, nullptr, out);
}

int baysar_model_create_ex(const void* desc_blob, size_t nbytes, int device, const baysar_options_t* options,
                           baysar_model_t** out) {
  if (!out) return fail("null output pointer");
  if (validate_blob(desc_blob, nbytes)) return 1;
  CUDA_OK(cudaSetDevice(device));
  baysar_model* m = new (std::nothrow) baysar_model();
  if (!m) return fail("out of host memory");
  m->device = device;
  if (options) m->opt = *options;
  m->blob_bytes = nbytes;
  m->blob_host.assign(static_cast<const unsigned char*>(desc_blob), static_cast<const unsigned char*>(desc_blob) + nbytes);
  cudaError_t e = cudaMalloc(&m->blob_dev, nbytes);
  if (e != cudaSuccess) { delete m; return fail(std::string("cudaMalloc(blob): ") + cudaGetErrorString(e)); }
  e = cudaMemcpy(m->blob_dev, desc_blob, nbytes, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { cudaFree(m->blob_dev); delete m; return fail(std::string("cudaMemcpy(blob): ") + cudaGetErrorString(e)); }

  SectionTable host{reinterpret_cast<const int64_t*>(m->blob_host.data()), m->blob_host.data()};
  SectionTable dev{reinterpret_cast<const int64_t*>(m->blob_host.data()), m->blob_dev};
  if (set_kernel_attributes()) { baysar_model_destroy(m); return 1; }
  {
    cudaError_t es = cudaStreamCreateWithFlags(&m->host_stream, cudaStreamNonBlocking);
    if (es == cudaSuccess) es = cudaEventCreateWithFlags(&m->last_done, cudaEventDisableTiming);
    if (es == cudaSuccess) es = cudaStreamCreateWithFlags(&m->copy_stream, cudaStreamNonBlocking);
    if (es != cudaSuccess) { baysar_model_destroy(m); return fail(std::string("cudaStreamCreate: ") + cudaGetErrorString(es)); }
  }
  m->I = static_cast<const int64_t*>(host.ptr(SEC_I));
  m->ch_i = static_cast<const int*>(host.ptr(SEC_CH_I));
  m->cl_i = static_cast<const int*>(host.ptr(SEC_CL_I));
  m->ch_d = static_cast<const double*>(host.ptr(SEC_CH_D));
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
  d.a11_acd = static_cast<const double*>(dev.ptr(SEC_A11_ACD));
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
  d.gr_i = static_cast<const int*>(dev.ptr(SEC_GR_I));
  d.gr_ul = static_cast<const int*>(dev.ptr(SEC_GR_UL));
  {
    // reciprocal spacings, search buckets and the 2^(k/64) table of the LOS stage, derived from the axes in the blob
    const los_tables::Host lt = los_tables::build(
        m->I, static_cast<const double*>(host.ptr(SEC_A11_NE)), static_cast<const double*>(host.ptr(SEC_A11_TE)),
        static_cast<const double*>(host.ptr(SEC_H_NE)), static_cast<const double*>(host.ptr(SEC_H_TE)),
        static_cast<const double*>(host.ptr(SEC_SPL_TX)), static_cast<const double*>(host.ptr(SEC_SPL_TY)));
    const size_t nd = lt.d.size() * sizeof(double), ni = lt.i.size() * sizeof(int);
    cudaError_t el = cudaMalloc(&m->los_tables_dev, nd + ni);
    if (el == cudaSuccess) el = cudaMemcpy(m->los_tables_dev, lt.d.data(), nd, cudaMemcpyHostToDevice);
    if (el == cudaSuccess) el = cudaMemcpy(m->los_tables_dev + nd, lt.i.data(), ni, cudaMemcpyHostToDevice);
    if (el != cudaSuccess) { baysar_model_destroy(m); return fail(std::string("LOS tables: ") + cudaGetErrorString(el)); }
    d.lt = los_tables::view(lt, reinterpret_cast<const double*>(m->los_tables_dev),
                            reinterpret_cast<const int*>(m->los_tables_dev + nd));
  }

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
    // the single- and three-tap builds (chord_stage.cuh TAPS < 2) never tabulate a Doppler kernel: no tap table for them
    if (ch[CH_B_OPTIONAL] || ch[CH_NARROW]) kd_max = 0;
    ChordSmem s = chord_smem_layout(ch[CH_F], ch[CH_HALO], ch[CH_TAPS_MAX], kd_max, 1, ch[CH_N_LINES]);
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
  const bool ifc_enabled = m->opt.tensor_cores >= 0;
  const int min_taps = m->opt.tensor_cores > 0 ? 1 : (m->opt.ifc_min_taps > 0 ? m->opt.ifc_min_taps : 160);
  if (m->opt.ifc_budget_mb > 0) m->ifc_budget_bytes = (int64_t)m->opt.ifc_budget_mb << 20;
  const bool pxc_enabled = m->opt.folded_tile >= 0;  // -1: keep the every-column contraction
  const int pxc_tn = m->opt.folded_tile == 64 ? 64 : 128;
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
      m->chord_smem[c] = chord_smem_layout(ch[CH_F], ch[CH_HALO], ch[CH_TAPS_MAX], 0, 0, ch[CH_N_LINES]).total;
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
        fx.no_extrapolation = true;
        for (int px = 0; px < ch[CH_P]; ++px) {
          const double fr = pix_frac_host[ch[CH_PIX_OFF] + px];
          if (!(fr >= -1e-6 && fr <= 1.0 + 1e-6)) fx.no_extrapolation = false;
        }
        if (fx.x_l < 2 || fx.x_r < 2) fx.no_extrapolation = false;
        fx.inv_err.resize(ch[CH_P]);
        for (int px = 0; px < ch[CH_P]; ++px) fx.inv_err[px] = (float)(1.0 / err_host[px]);
        if (upload_pxc_plan(fx) || (m->opt.contraction_b_image >= 0 && build_pxc_b_image(fx))) {
          free_pxc_plan(fx); baysar_model_destroy(m); return 1;
        }
        m->pxc[c] = fx;
        m->split_pitch_max = fx.pitch > m->split_pitch_max ? fx.pitch : m->split_pitch_max;
        m->conv_pitch_max = fx.out_pitch > m->conv_pitch_max ? fx.out_pitch : m->conv_pitch_max;
        m->tiles_max = pxc::TSUM_SLOTS * fx.n_col_tiles > m->tiles_max ? pxc::TSUM_SLOTS * fx.n_col_tiles : m->tiles_max;  // tsum lives in trapz_part
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
  cudaFree(m->blob_dev); cudaFree(m->los_tables_dev); cudaFree(m->lines); cudaFree(m->comps); cudaFree(m->hdr); cudaFree(m->status);
  cudaFree(m->theta_stage); cudaFree(m->logp_stage); cudaFreeHost(m->out_pinned);
  if (m->host_stream) cudaStreamDestroy(m->host_stream);
  if (m->copy_stream) cudaStreamDestroy(m->copy_stream);
  for (cudaEvent_t e : m->copy_done) cudaEventDestroy(e);
  if (m->last_done) cudaEventDestroy(m->last_done);
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
  CUDA_OK(cudaStreamWaitEvent(stream, m->last_done, 0));
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
    CUDA_OK(launch_k(!m->opt.serial_launches, finalize_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, stream, comps, ncols, n, logp, st));
    CUDA_OK(cudaGetLastError());
    m->last_launches += 1;
  }
  if (mark(m, ST_FINAL, stream)) return 1;
  CUDA_OK(cudaEventRecord(m->last_done, stream));
  return 0;
}

int baysar_eval_spectra(baysar_model_t* m, const double* theta, int64_t n, int chord, double* spectra,
                        float* fine_pre, float* fine_post, uint32_t* status, void* stream_) {
  if (!m || !theta || !spectra) return fail("null argument");
  if (chord < 0 || chord >= m->I[I_N_CHORDS]) return fail("chord index out of range");
  if (n <= 0) return 0;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CUDA_OK(cudaSetDevice(m->device));
  CUDA_OK(cudaStreamWaitEvent(stream, m->last_done, 0));
  if (ensure_workspace(m, n)) return 1;
  uint32_t* st = status ? status : m->status;
  if (launch_los(m, theta, n, m->lines, m->comps, nullptr, st, stream)) return 1;
  if (launch_chords(m, theta, n, chord, 1, m->lines, m->comps, spectra, fine_pre, fine_post, st, stream)) return 1;
  CUDA_OK(cudaEventRecord(m->last_done, stream));
  return 0;
}

int baysar_eval_lines(baysar_model_t* m, const double* theta, int64_t n, double* lines, double* profiles,
                      uint32_t* status, void* stream_) {
  if (!m || !theta || !lines) return fail("null argument");
  if (n <= 0) return 0;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CUDA_OK(cudaSetDevice(m->device));
  CUDA_OK(cudaStreamWaitEvent(stream, m->last_done, 0));
  if (ensure_workspace(m, n)) return 1;
  uint32_t* st = status ? status : m->status;
  if (launch_los(m, theta, n, lines, m->comps, profiles, st, stream)) return 1;
  CUDA_OK(cudaEventRecord(m->last_done, stream));
  return 0;
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
  // Large batches: the copy of theta (multi-species, 65 536 thetas: 8.9 MB, a few hundred microseconds over PCIe) is
  // pipelined with the evaluation in chunks -- chunk c + 1 crosses the bus on a second stream while chunk c is evaluated.
  // Small batches (the 4096-theta Balmer step is one wave of the line-of-sight kernel) stay one copy and one evaluation.
  const int64_t HOST_CHUNK = m->opt.host_chunk_thetas > 0 ? m->opt.host_chunk_thetas : 32768;
  if (m->opt.host_chunk_thetas >= 0 && n >= 2 * HOST_CHUNK) {
    const int64_t n_chunks = (n + HOST_CHUNK - 1) / HOST_CHUNK;
    while ((int64_t)m->copy_done.size() < n_chunks) {
      cudaEvent_t ev;
      CUDA_OK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
      m->copy_done.push_back(ev);
    }
    // (the previous call ended with a synchronisation of host_stream: nothing still reads the staging buffer)
    for (int64_t c = 0; c < n_chunks; ++c) {
      const int64_t off = c * HOST_CHUNK, nc = n - off < HOST_CHUNK ? n - off : HOST_CHUNK;
      CUDA_OK(cudaMemcpyAsync(m->theta_stage + off * np, theta_host + off * np, sizeof(double) * nc * np, cudaMemcpyHostToDevice,
                              m->copy_stream));
      CUDA_OK(cudaEventRecord(m->copy_done[c], m->copy_stream));
    }
    int launches = 0;
    for (int64_t c = 0; c < n_chunks; ++c) {
      const int64_t off = c * HOST_CHUNK, nc = n - off < HOST_CHUNK ? n - off : HOST_CHUNK;
      CUDA_OK(cudaStreamWaitEvent(m->host_stream, m->copy_done[c], 0));
      if (baysar_eval_logp(m, m->theta_stage + off * np, nc, m->logp_stage + off, status_dev + off, nullptr, m->host_stream)) return 1;
      launches += m->last_launches;
    }
    m->last_launches = launches;
  } else {
    CUDA_OK(cudaMemcpyAsync(m->theta_stage, theta_host, sizeof(double) * n * np, cudaMemcpyHostToDevice, m->host_stream));
    if (baysar_eval_logp(m, m->theta_stage, n, m->logp_stage, status_dev, nullptr, m->host_stream)) return 1;
  }
  const size_t out_bytes = (sizeof(double) + (status_host ? sizeof(uint32_t) : 0)) * (size_t)n;
  CUDA_OK(cudaMemcpyAsync(m->out_pinned, m->logp_stage, out_bytes, cudaMemcpyDeviceToHost, m->host_stream));
  CUDA_OK(cudaStreamSynchronize(m->host_stream));
  memcpy(logp_host, m->out_pinned, sizeof(double) * n);
  if (status_host) memcpy(status_host, static_cast<unsigned char*>(m->out_pinned) + sizeof(double) * n, sizeof(uint32_t) * n);
  return 0;
}

int baysar_model_set_profiling(baysar_model_t* m, int on) {
  if (!m) return fail("null model");
  m->prof_used = 0;
  m->profiling = on != 0;
  return 0;
}

int baysar_model_get_profile(baysar_model_t* m, double* stage_ms, int64_t* calls) {
  if (!m || !stage_ms || !calls) return fail("null argument");
  CUDA_OK(cudaSetDevice(m->device));
  for (int k = 0; k < ST_COUNT; ++k) stage_ms[k] = 0.0;
  *calls = 0;
  for (size_t e = 0; e < m->prof_used; ++e) {
    if (m->prof_slots[e] == ST_BEGIN) { *calls += 1; continue; }
    CUDA_OK(cudaEventSynchronize(m->prof_events[e]));
    float ms = 0.f;
    CUDA_OK(cudaEventElapsedTime(&ms, m->prof_events[e - 1], m->prof_events[e]));
    stage_ms[m->prof_slots[e]] += ms;
  }
  m->prof_used = 0;  // the events stay in the pool
  return 0;
}

int baysar_stretch_propose(const double* walkers, const double* complement, int n_params, int64_t n_walkers,
                           int64_t n_complement, double a, uint64_t seed, uint64_t step, int64_t walker_offset,
                           double* proposal, double* log_factor, void* stream_) {
  if (!walkers || !complement || !proposal || !log_factor) return fail("null argument");
  if (n_params < 1 || n_complement < 1 || !(a > 1.0)) return fail("bad stretch-move arguments");
  if (n_walkers <= 0) return 0;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ens::stretch_propose_kernel<<<(unsigned)((n_walkers + 127) / 128), 128, 0, stream>>>(
      walkers, complement, n_params, n_walkers, n_complement, a, seed, step, walker_offset, proposal, log_factor);
  CUDA_OK(cudaGetLastError());
  return 0;
}

int baysar_stretch_accept(double* walkers, double* logp, const double* proposal, const double* logp_new,
                          const double* log_factor, const uint32_t* status_new, int n_params, int64_t n_walkers,
                          uint64_t seed, uint64_t step, int64_t walker_offset, unsigned long long* n_accepted, void* stream_) {
  if (!walkers || !logp || !proposal || !logp_new || !log_factor) return fail("null argument");
  if (n_walkers <= 0) return 0;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ens::stretch_accept_kernel<<<(unsigned)((n_walkers + 127) / 128), 128, 0, stream>>>(
      walkers, logp, proposal, logp_new, log_factor, status_new, n_params, n_walkers, seed, step, walker_offset, n_accepted);
  CUDA_OK(cudaGetLastError());
  return 0;
}

}  // extern "C"
