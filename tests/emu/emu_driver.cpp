// TEST INFRASTRUCTURE ONLY — runs the product's kernel sources on CPU threads (see cpu_emu.h) for one model blob.
#define BAYSAR_CPU_EMU 1
#include <cstdlib>
#include "../../baysar_b200/csrc/chord_stage.cuh"
#include "../../baysar_b200/csrc/finalize.cuh"
#include "../../baysar_b200/csrc/los_stage.cuh"
#include "../../baysar_b200/csrc/los_tables.h"

static ModelDev view(const unsigned char* blob) {
  const int64_t* hdr = reinterpret_cast<const int64_t*>(blob);
  auto P = [&](int s) { return (const void*)(blob + hdr[3 + 2 * s]); };
  ModelDev d;
  d.I = (const int64_t*)P(SEC_I); d.D = (const double*)P(SEC_D); d.los_x = (const double*)P(SEC_LOS_X);
  d.los_w = (const double*)P(SEC_LOS_W); d.prof_d = (const double*)P(SEC_PROF_D); d.sp_i = (const int*)P(SEC_SP_I);
  d.sp_d = (const double*)P(SEC_SP_D); d.elem_dens_idx = (const int*)P(SEC_ELEM_DENS_IDX);
  d.a11_ne = (const double*)P(SEC_A11_NE); d.a11_te = (const double*)P(SEC_A11_TE);
  d.a11_scd = (const double*)P(SEC_A11_SCD); d.a11_acd = (const double*)P(SEC_A11_ACD); d.h_ne = (const double*)P(SEC_H_NE); d.h_te = (const double*)P(SEC_H_TE);
  d.h_tab = (const double*)P(SEC_H_TAB); d.tau_grid = (const double*)P(SEC_TAU_GRID);
  d.spl_tx = (const double*)P(SEC_SPL_TX); d.spl_ty = (const double*)P(SEC_SPL_TY); d.spl_c = (const double*)P(SEC_SPL_C);
  d.ul_i = (const int*)P(SEC_UL_I); d.ul_d = (const double*)P(SEC_UL_D); d.pr_i = (const int*)P(SEC_PR_I); d.pr_d = (const double*)P(SEC_PR_D);
  d.cso_off = (const int*)P(SEC_CSO_OFF); d.cso_ul = (const int*)P(SEC_CSO_UL); d.ch_i = (const int*)P(SEC_CH_I);
  d.ch_d = (const double*)P(SEC_CH_D); d.cl_i = (const int*)P(SEC_CL_I); d.cl_d = (const double*)P(SEC_CL_D);
  d.comp_d = (const double*)P(SEC_COMP_D); d.y = (const double*)P(SEC_Y); d.err = (const double*)P(SEC_ERR);
  d.ifn = (const double*)P(SEC_IF); d.pix_idx = (const int*)P(SEC_PIX_IDX); d.pix_frac = (const double*)P(SEC_PIX_FRAC);
  d.mesh_x = (const double*)P(SEC_MESH_X); d.mesh_lo = (const int*)P(SEC_MESH_LO); d.pr_aux = (const int*)P(SEC_PR_AUX);
  d.gr_i = (const int*)P(SEC_GR_I); d.gr_ul = (const int*)P(SEC_GR_UL);
  // derived LOS tables: one set per process is enough for the tests (kept alive for the duration of the call chain)
  static thread_local los_tables::Host lt;
  lt = los_tables::build(d.I, d.a11_ne, d.a11_te, d.h_ne, d.h_te, d.spl_tx, d.spl_ty);
  d.lt = los_tables::view(lt, lt.d.data(), lt.i.data());
  return d;
}

// Tensor-core path in emulation: the spectral stage (MODE 1) and the pixel stage run through the thread emulator; the
// tcgen05 contraction itself cannot, so a plain loop multiplies the TF32-split operands the way 3xTF32 does
// (hi*hi + hi*lo + lo*hi, float accumulation).  Checks the hand-over logic (padding, aux, partial sums, fetch).
extern "C" int emu_eval_split(const unsigned char* blob, const double* theta, long n, int chord, double* logp,
                              unsigned* status, double* comps, double* lines, double* spectra) {
  ModelDev m = view(blob);
  const int64_t* I = m.I;
  const int L = (int)I[I_L], n_chords = (int)I[I_N_CHORDS];
  std::vector<double> hdr(n * BAYSAR_HDR_STRIDE);
  LosArgs la{theta, n, lines, comps, hdr.data(), nullptr, status};
  EmuDim3 g;
  g.x = (unsigned)((n + 3) / 4);
  emu_launch(los_stage_kernel<4>, g, 128u, 4 * los_smem_per_warp(L, (int)I[I_N_PARAMS], (int)I[I_N_SPECIES]), m, la);
  for (int c = 0; c < n_chords; ++c) {
    const int* ch = m.ch_i + c * CH_I_COUNT;
    const int F = ch[CH_F], KI = ch[CH_KI], k_lo = ch[CH_IF_KLO], k_hi = ch[CH_IF_KHI], o = (KI - 1) / 2;
    int kd_max = 0;
    for (int k = 0; k < ch[CH_N_LINES]; ++k) {
      int kd = m.cl_i[(ch[CH_LINE_OFF] + k) * CL_I_COUNT + CL_KD];
      kd_max = kd > kd_max ? kd : kd_max;
    }
    ChordSmem s = chord_smem_layout(F, ch[CH_HALO], ch[CH_TAPS_MAX], kd_max, ch[CH_B_OPTIONAL] ? 0 : 1, ch[CH_N_LINES]);
    const int pad = 64, n_tiles = (F + 127) / 128;
    const int64_t pitch = pad + F + 96, cpitch = (int64_t)n_tiles * 128;
    std::vector<float> spec(n * pitch, 7.f), conv(n * cpitch, 0.f);
    std::vector<double> aux(n * BAYSAR_AUX_STRIDE), part(n * n_tiles, 0.0);
    ChordArgs ca{};
    chord_args_set_chord(ca, m.I, m.ch_i, m.ch_d, m.cl_i, c);
    ca.theta = theta; ca.n = n; ca.lines = lines; ca.components = comps; ca.status = status;
    ca.spec = spec.data(); ca.split_pitch = pitch; ca.split_pad = pad; ca.aux = aux.data();
    EmuDim3 gc;
    gc.x = (unsigned)n; gc.y = 1;
    emu_launch(chord_stage_kernel<256, 1>, gc, 256u, s.total, m, ca);
    const double* w = m.ifn + ch[CH_IF_OFF];
    for (long r = 0; r < n; ++r) {
      const float* sp = spec.data() + r * pitch + pad;
      const float clip = (float)aux[r * BAYSAR_AUX_STRIDE + 1];
      for (int j = 0; j < F; ++j) {
        float acc0 = 0.f, acc1 = 0.f;
        for (int k = k_lo; k < k_hi; ++k) {
          int kp = j + o - k;
          if (kp < -pad || kp >= F + 96) continue;
          float wh = emu_tf32((float)w[k]), wl = emu_tf32((float)w[k] - wh);
          float h, l;
          split_tf32_pair(sp[kp], h, l);  // what the contraction's producer threads do before tcgen05.st
          acc0 += h * wh;
          acc1 += h * wl + l * wh;
        }
        float v = acc0 + acc1;
        v = v < clip ? clip : v;
        conv[r * cpitch + j] = v;
        part[r * n_tiles + j / 128] += ((j == 0 || j == F - 1) ? 0.5 : 1.0) * (double)v;
      }
    }
    PixelArgs pa{};
    pa.theta = theta; pa.n = n; pa.chord = c; pa.conv = conv.data(); pa.conv_pitch = cpitch; pa.trapz_part = part.data();
    pa.n_tiles = n_tiles; pa.aux = aux.data(); pa.components = comps; pa.spectra = (c == chord) ? spectra : nullptr;
    pa.status = status;
    emu_launch(pixel_stage_kernel<128>, gc, 128u, 64 * sizeof(double), m, pa);
  }
  EmuDim3 gf;
  gf.x = (unsigned)((n + 255) / 256);
  emu_launch(finalize_kernel, gf, 256u, 0, (const double*)comps, (int)(n_chords + I[I_N_PRIORS]), (int64_t)n, logp, status);
  return 0;
}

extern "C" int emu_eval(const unsigned char* blob, const double* theta, long n, int chord_for_spectra, double* logp,
                        unsigned* status, double* comps, double* lines, double* profiles, double* spectra,
                        float* fine_pre, float* fine_post) {
  ModelDev m = view(blob);
  const int64_t* I = m.I;
  const int L = (int)I[I_L], n_chords = (int)I[I_N_CHORDS];
  std::vector<double> hdr(n * BAYSAR_HDR_STRIDE);
  LosArgs la{theta, n, lines, comps, hdr.data(), profiles, status};
  EmuDim3 g;
  if (getenv("BAYSAR_EMU_LOS_TEAM")) {  // the one-CTA-per-theta form the product uses for long lines of sight
    g.x = (unsigned)n;
    emu_launch(los_stage_kernel<4, 1, 1, 4>, g, 128u, los_smem_per_warp(L, (int)I[I_N_PARAMS], (int)I[I_N_SPECIES]) + los_team_xch_bytes(4), m, la);
  } else {
    g.x = (unsigned)((n + 3) / 4);
    emu_launch(los_stage_kernel<4>, g, 128u, 4 * los_smem_per_warp(L, (int)I[I_N_PARAMS], (int)I[I_N_SPECIES]), m, la);
  }
  for (int c = 0; c < n_chords; ++c) {
    const int* ch = m.ch_i + c * CH_I_COUNT;
    int kd_max = 0;
    for (int k = 0; k < ch[CH_N_LINES]; ++k) {
      int kd = m.cl_i[(ch[CH_LINE_OFF] + k) * CL_I_COUNT + CL_KD];
      kd_max = kd > kd_max ? kd : kd_max;
    }
    ChordSmem s = chord_smem_layout(ch[CH_F], ch[CH_HALO], ch[CH_TAPS_MAX], kd_max, 1, ch[CH_N_LINES]);
    bool want = c == chord_for_spectra;
    ChordArgs ca{};
    chord_args_set_chord(ca, m.I, m.ch_i, m.ch_d, m.cl_i, c);
    ca.theta = theta; ca.n = n; ca.lines = lines; ca.components = comps;
    ca.spectra = want ? spectra : nullptr; ca.fine_pre = want ? fine_pre : nullptr;
    ca.fine_post = want ? fine_post : nullptr; ca.status = status;
    EmuDim3 gc;
    gc.x = (unsigned)n; gc.y = 1;
    emu_launch(chord_stage_kernel<256, 0>, gc, 256u, s.total, m, ca);
  }
  EmuDim3 gf;
  gf.x = (unsigned)((n + 255) / 256);
  emu_launch(finalize_kernel, gf, 256u, 0, (const double*)comps, (int)(n_chords + I[I_N_PRIORS]), (int64_t)n, logp, status);
  return 0;
}
