// TEST INFRASTRUCTURE ONLY — runs the product's kernel sources on CPU threads (see cpu_emu.h) for one model blob.
#define BAYSAR_CPU_EMU 1
#include "../../baysar_b200/csrc/chord_stage.cuh"
#include "../../baysar_b200/csrc/finalize.cuh"
#include "../../baysar_b200/csrc/los_stage.cuh"

static ModelDev view(const unsigned char* blob) {
  const int64_t* hdr = reinterpret_cast<const int64_t*>(blob);
  auto P = [&](int s) { return (const void*)(blob + hdr[3 + 2 * s]); };
  ModelDev d;
  d.I = (const int64_t*)P(SEC_I); d.D = (const double*)P(SEC_D); d.los_x = (const double*)P(SEC_LOS_X);
  d.los_w = (const double*)P(SEC_LOS_W); d.prof_d = (const double*)P(SEC_PROF_D); d.sp_i = (const int*)P(SEC_SP_I);
  d.sp_d = (const double*)P(SEC_SP_D); d.elem_dens_idx = (const int*)P(SEC_ELEM_DENS_IDX);
  d.a11_ne = (const double*)P(SEC_A11_NE); d.a11_te = (const double*)P(SEC_A11_TE);
  d.a11_scd = (const double*)P(SEC_A11_SCD); d.h_ne = (const double*)P(SEC_H_NE); d.h_te = (const double*)P(SEC_H_TE);
  d.h_tab = (const double*)P(SEC_H_TAB); d.tau_grid = (const double*)P(SEC_TAU_GRID);
  d.spl_tx = (const double*)P(SEC_SPL_TX); d.spl_ty = (const double*)P(SEC_SPL_TY); d.spl_c = (const double*)P(SEC_SPL_C);
  d.ul_i = (const int*)P(SEC_UL_I); d.ul_d = (const double*)P(SEC_UL_D); d.pr_i = (const int*)P(SEC_PR_I); d.pr_d = (const double*)P(SEC_PR_D);
  d.cso_off = (const int*)P(SEC_CSO_OFF); d.cso_ul = (const int*)P(SEC_CSO_UL); d.ch_i = (const int*)P(SEC_CH_I);
  d.ch_d = (const double*)P(SEC_CH_D); d.cl_i = (const int*)P(SEC_CL_I); d.cl_d = (const double*)P(SEC_CL_D);
  d.comp_d = (const double*)P(SEC_COMP_D); d.y = (const double*)P(SEC_Y); d.err = (const double*)P(SEC_ERR);
  d.ifn = (const double*)P(SEC_IF); d.pix_idx = (const int*)P(SEC_PIX_IDX); d.pix_frac = (const double*)P(SEC_PIX_FRAC);
  return d;
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
  g.x = (unsigned)((n + 3) / 4);
  emu_launch(los_stage_kernel<4>, g, 128u, 4 * los_smem_per_warp(L), m, la);
  for (int c = 0; c < n_chords; ++c) {
    const int* ch = m.ch_i + c * CH_I_COUNT;
    int kd_max = 0;
    for (int k = 0; k < ch[CH_N_LINES]; ++k) {
      int kd = m.cl_i[(ch[CH_LINE_OFF] + k) * CL_I_COUNT + CL_KD];
      kd_max = kd > kd_max ? kd : kd_max;
    }
    ChordSmem s = chord_smem_layout(ch[CH_F], ch[CH_HALO], ch[CH_TAPS_MAX], kd_max);
    bool want = c == chord_for_spectra;
    ChordArgs ca{theta, n, c, lines, comps, want ? spectra : nullptr, want ? fine_pre : nullptr,
                 want ? fine_post : nullptr, status};
    EmuDim3 gc;
    gc.x = (unsigned)n; gc.y = 1;
    emu_launch(chord_stage_kernel<256>, gc, 256u, s.total, m, ca);
  }
  EmuDim3 gf;
  gf.x = (unsigned)((n + 255) / 256);
  emu_launch(finalize_kernel, gf, 256u, 0, (const double*)comps, (int)(n_chords + I[I_N_PRIORS]), (int64_t)n, logp, status);
  return 0;
}
