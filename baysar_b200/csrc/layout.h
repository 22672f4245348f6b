/*
 * Descriptor-blob layout shared by the host-side model compiler (baysar_b200/model.py parses the enums below with a
 * regex — keep one enumerator per line, no explicit values) and the CUDA side.
 *
 * blob = [ int64 magic, int64 version, int64 n_sections, { int64 offset_bytes, int64 count } x n_sections ] payload...
 * Every section is 16-byte aligned.  Element types are fixed per section (see comments).
 */
#ifndef BAYSAR_LAYOUT_H
#define BAYSAR_LAYOUT_H

#define BAYSAR_BLOB_MAGIC 0x4252415359534142LL /* "BAYSYSAB" */
#define BAYSAR_BLOB_VERSION 13

enum BaysarSection {
  SEC_I,              /* int64  [I_COUNT]           model-wide integer scalars                         */
  SEC_D,              /* double [D_COUNT]           model-wide double scalars                          */
  SEC_LOS_X,          /* double [L]                 line-of-sight coordinate                           */
  SEC_LOS_W,          /* double [L]                 trapz weights of the LOS grid                      */
  SEC_PROF_D,         /* double [PROF_D_COUNT]      profile-family constants                           */
  SEC_SP_I,           /* int32  [n_species][SP_I_COUNT]                                                 */
  SEC_SP_D,           /* double [n_species][SP_D_COUNT]                                                 */
  SEC_ELEM_DENS_IDX,  /* int32  [n_elem]            theta index of the density of each impurity element */
  SEC_A11_NE,         /* double [a11_nne]           adf11 (scd) density grid                           */
  SEC_A11_TE,         /* double [a11_nte]                                                              */
  SEC_A11_SCD,        /* double [a11_nne][a11_nte]  ionisation rate, raw values                        */
  SEC_A11_ACD,        /* double [a11_nne][a11_nte]  recombination rate on the same grid (used when I_RECOMBINING) */
  SEC_H_NE,           /* double [h_nne]             adf15 density grid                                 */
  SEC_H_TE,           /* double [h_nte]                                                                */
  SEC_H_TAB,          /* double [n_htab][h_nne][h_nte]  log(PEC)                                       */
  SEC_TAU_GRID,       /* double [n_tau]             log10 tau grid ("magical_tau")                     */
  SEC_SPL_TX,         /* double [spl_nx]            bicubic spline knots in ne                         */
  SEC_SPL_TY,         /* double [spl_ny]            bicubic spline knots in Te                         */
  SEC_SPL_C,          /* double, per ADAS line group g (GR_COEF_OFF): [n_tau-1][spl_nx-4][spl_ny-4][n_k][4]  B-spline
                         coefficients of log(ne*TEC) of the group's n_k lines side by side; the 4 doubles (32 bytes, one
                         256-bit load) are (tau slice i, slice i+1) of column iy and of column iy+1 (clipped at the last) */
  SEC_UL_I,           /* int32  [n_ulines][UL_I_COUNT]  unique (line-of-sight) lines                    */
  SEC_UL_D,           /* double [n_ulines][UL_D_COUNT]                                                  */
  SEC_PR_I,           /* int32  [n_priors][PR_I_COUNT]                                                  */
  SEC_PR_D,           /* double [n_priors][PR_D_COUNT]                                                  */
  SEC_CSO_OFF,        /* int32  [n_elem + 1]        ChargeStateOrderPrior: per element, offsets into CSO_UL */
  SEC_CSO_UL,         /* int32  [..]                unique-line ids ordered by charge state             */
  SEC_CH_I,           /* int32  [n_chords][CH_I_COUNT]                                                  */
  SEC_CH_D,           /* double [n_chords][CH_D_COUNT]                                                  */
  SEC_CL_I,           /* int32  [n_clines][CL_I_COUNT]  per-chord line objects                          */
  SEC_CL_D,           /* double [n_clines][CL_D_COUNT]                                                  */
  SEC_COMP_D,         /* double [n_comps][2]        multiplet members: centre wavelength, fraction      */
  SEC_Y,              /* double [sum P]             measured spectra                                    */
  SEC_ERR,            /* double [sum P]             error model                                         */
  SEC_IF,             /* double [sum KI]            fine-grid instrument function                       */
  SEC_PIX_IDX,        /* int32  [sum P]             default wavelength map: lower fine-grid index       */
  SEC_PIX_FRAC,       /* double [sum P]             default wavelength map: interpolation weight        */
  SEC_MESH_X,         /* double [mesh_ne_n + mesh_te_n]  MeshLine knot positions (density, then temperature)  */
  SEC_MESH_LO,        /* int32  [2][L]              MeshLine: lower knot of every LOS point (density, temperature) */
  SEC_PR_AUX,         /* int32  [..]                index lists of the optional priors (theta indices; PR_I0 = offset, PR_I1 = count) */
  SEC_GR_I,           /* int32  [n_groups][GR_I_COUNT]  ADAS line groups: the unique lines of one species (one tau), evaluated together */
  SEC_GR_UL,          /* int32  [sum n_k]           unique-line id of every group member                 */
  SEC_COUNT
};

enum BaysarI {
  I_N_PARAMS,
  I_L,
  I_N_CHORDS,
  I_N_SPECIES,
  I_N_ELEM,
  I_N_ULINES,
  I_N_PRIORS,
  I_N_CLINES,
  I_PROFILE_KIND,       /* 0 = EsymmtricCauchyPlasma, 1 = BowmanTeePlasma, 2 = MeshPlasma, 3 = closed-form pair (I_NE_FN, I_TE_FN) */
  I_NE_IDX,             /* first theta index of the electron_density slice */
  I_TE_IDX,
  I_HAS_HYDROGEN,
  I_H_SPECIES,          /* index into the species table */
  I_COLD_NEUTRALS,
  I_COLD_IONS,
  I_THERMALISED,
  I_CAL_WAVE,
  I_CAL_INT,
  I_CAL_IF,
  I_A11_NNE,
  I_A11_NTE,
  I_H_NNE,
  I_H_NTE,
  I_N_HTAB,
  I_N_TAU,
  I_SPL_NX,
  I_SPL_NY,
  I_N_TEC,
  I_P_MAX,
  I_F_MAX,
  I_MESH_NE_N,          /* MeshLine: number of density / temperature knots */
  I_MESH_TE_N,
  I_N_GROUPS,           /* ADAS line groups */
  I_RECOMBINING,        /* plasma.recombining: the neutral profile gains n1 ne acd tau before the ionisation loss (plasmas.py:735-737) */
  I_REVERSE_EN,         /* plasma.reverse_electroneutrality: the sampled profile is the main-ion density (plasmas.py:667-671) */
  I_NE_FN,              /* I_PROFILE_KIND 3: BaysarProfileFn of the density / temperature profile */
  I_TE_FN,
  I_GAUSS_LIKELIHOOD,   /* chord likelihood -0.5 sum r^2 instead of the Cauchy form (spectrometers.py:185-188) */
  I_COUNT
};

enum BaysarD {
  D_RESERVED,
  D_COUNT
};

enum BaysarProfD {
  PROF_NE_CENTRE_FIXED, /* 1.0 if the density profile has a fixed centre (dr given), else the centre is sampled */
  PROF_NE_CENTRE,
  PROF_TE_CENTRE,
  PROF_NE_FLOOR,        /* 1e-3 */
  PROF_BACKGROUND,      /* BowmanTee: 1.0 if the last parameter of each profile is log10 of a background level */
  PROF_MESH_NE_LOG,     /* MeshLine: 1.0 if the knot parameters are log10 values */
  PROF_MESH_TE_LOG,
  PROF_ESYM_GAUSSIAN,   /* EsymmtricCauchyPlasma(ptype="gaussian"): exp(-0.5 z^2) instead of 1 / (1 + z^2) */
  PROF_NE_AUX,          /* closed-form families: a constant of the density function (fixed Gaussian centre, top-hat grid step) */
  PROF_NE_AUX_SET,      /* 1.0 if PROF_NE_AUX is in use (GaussianPlasma with a fixed centre) */
  PROF_TE_AUX,
  PROF_TE_AUX_SET,
  PROF_D_COUNT
};

/* closed-form profile functions of one coordinate (reference lineshapes.py:483-591, :855-929) */
/* appended values keep the blob layout: BAYSAR_BLOB_VERSION is unchanged by PF_POISSON */
enum BaysarProfileFn {
  PF_EXP_DECAY,           /* theta = [log10 a, log10 b]:        max(a (1 + b)^-x, 0.01)                                   */
  PF_DOUBLE_EXP_DECAY,    /* [log10 a, log10 b, x0, log10 f, k]: max(a ((1 + b) exp(f tanh(k (x - x0))))^(x0 - x), 0.01)    */
  PF_GAUSSIAN,            /* [log10 I, log10 fwhm (, centre)]:   max(I exp(-0.5 ((x - c) / sigma)^2), 0.01); fixed centre = aux */
  PF_FLAT,                /* [log10 v]:                          v                                                          */
  PF_LEFT_TOP_HAT,        /* [log10 v, edge]:                    v where x < edge + aux (the grid step), 1e10 elsewhere     */
  PF_LEFT_RIGHT_TOP_HAT,  /* [log10 l, log10 r (, centre)]:      l left of the centre, r right of it, 0 on it; the temperature
                             of a DoubleSlabPlasma takes the centre from the density's third parameter                      */
  PF_POISSON,             /* [log10 a, log10 nu]:                a p / max_LOS(p), p = nu^x e^-nu / Gamma(x + 1) (lineshapes.py:537-550;
                             the kernel returns p, the profile stage scales it by a / max)                                  */
  PF_COUNT
};

enum BaysarSpI { SP_DENS_IDX, SP_TAU_IDX, SP_VEL_IDX, SP_TI_IDX, SP_IS_H, SP_IS_IMPURITY, SP_I_COUNT };
enum BaysarSpD { SP_MASS, SP_D_COUNT };

enum BaysarGrI {
  GR_SPECIES,   /* index into the species table (density, tau) */
  GR_NK,        /* lines in the group */
  GR_FIRST,     /* offset of the group's members in SEC_GR_UL */
  GR_COEF_OFF,  /* offset of the group's coefficients in SEC_SPL_C, in units of 4 doubles */
  GR_I_COUNT
};

enum BaysarUlI {
  UL_KIND,     /* 0 Balmer, 1 ADAS406, 2 XLine */
  UL_SPECIES,
  UL_TAB_A,    /* Balmer: excitation log-PEC table; ADAS406: TEC table; XLine: theta index */
  UL_TAB_B,    /* Balmer: recombination log-PEC table */
  UL_I_COUNT
};

enum BaysarUlD {
  UL_CWL0,     /* Balmer: rest wavelength */
  UL_LOMAN_A,  /* Balmer: Stark-width parameterisation a_ij, b_ij, c_ij */
  UL_LOMAN_B,
  UL_LOMAN_C,
  UL_MASS,
  UL_D_COUNT
};

enum BaysarPriorKind {
  PRIOR_STATIC_PRESSURE,
  PRIOR_NEUTRAL_FRACTION,
  PRIOR_STARK_TO_PEAK,
  PRIOR_ANTIPROTON,
  PRIOR_MAIN_ION_FRACTION,
  PRIOR_CHARGE_STATE_ORDER,
  PRIOR_WAVELENGTH_CAL,
  PRIOR_INTENSITY_CAL,
  PRIOR_GAUSSIAN_IF,
  PRIOR_CURVATURE,
  PRIOR_TDV,              /* optional priors (passed in priors=[...], reference priors.py): ElectronDensityTDVCost */
  PRIOR_PEAK_TE,          /* PeakTePrior                                                                          */
  PRIOR_TAU,              /* TauPrior: aux = pairs of tau theta indices (consecutive ions of an element)          */
  PRIOR_WALL_CONDITIONS,  /* WallConditionsPrior                                                                  */
  PRIOR_SIMPLE_WALL,      /* SimpleWallPrior                                                                      */
  PRIOR_TE_MIN,           /* TeMinPrior                                                                           */
  PRIOR_TE_TAU,           /* TeTauPrior: aux = tau theta index of every impurity species                          */
  PRIOR_IMPURITY_TAU,     /* ImpurityTauPrior: aux = pairs of tau theta indices (equal charge, consecutive elements) */
  PRIOR_BOWMAN_TEE,       /* BowmanTeePrior                                                                       */
  PRIOR_KIND_COUNT
};
enum BaysarPrI { PR_KIND, PR_I0, PR_I1, PR_I_COUNT };
enum BaysarPrD { PR_D0, PR_D1, PR_D2, PR_D3, PR_D_COUNT };

enum BaysarChI {
  CH_P,
  CH_F,
  CH_KI,         /* instrument-function taps */
  CH_IF_KLO,     /* first / one-past-last significant tap */
  CH_IF_KHI,
  CH_N_LINES,
  CH_LINE_OFF,
  CH_BG_IDX,
  CH_CAL_IDX,
  CH_CALWAVE_IDX,
  CH_CALINT_IDX,
  CH_PIX_OFF,    /* offset into SEC_Y / SEC_ERR / SEC_PIX_* */
  CH_IF_OFF,
  CH_HALO,       /* zero padding (multiple of 4) on each side of the fine-grid shared-memory buffers */
  CH_TAPS_MAX,   /* capacity of the shared-memory tap buffer */
  CH_B_OPTIONAL, /* 1: no Gaussian components and every Balmer line is a guaranteed single Doppler tap, so the
                    spectral stage of the tensor-core path needs no second fine-grid buffer */
  CH_NARROW,     /* 1: every Balmer line of the chord has at most three Doppler taps for every theta inside the bounds
                    (cold neutrals): the chord kernel's multi-tap path is not needed */
  CH_I_COUNT
};
enum BaysarChD {
  CH_X0,            /* fine grid = X0 + j * DELTA (numpy arange), bit-exact */
  CH_DELTA,
  CH_DL,            /* mean spacing, the trapz weight */
  CH_CALWAVE_CWL,   /* default wavelength map constants */
  CH_CALWAVE_DISP,
  CH_PIX_MEAN,      /* mean(arange(P)) */
  CH_X_MEAN,        /* mean of the fine grid: centre of the sampled Gaussian instrument function */
  CH_D_COUNT
};

enum BaysarClI {
  CL_KIND,
  CL_ULINE,
  CL_SPECIES,
  CL_NCOMP,
  CL_COMP_OFF,
  CL_KD,
  CL_BRANK,      /* Balmer lines: rank among the chord's Balmer lines (slot of the line's shape record in shared memory) */
  CL_PAD,
  CL_I_COUNT
};
enum BaysarClD {
  CL_CWL0,
  CL_LOMAN_A,
  CL_LOMAN_B,
  CL_LOMAN_C,
  CL_MASS,
  CL_DOP_START,   /* Doppler tap grid = DOP_START + k * DOP_DELTA (numpy arange) */
  CL_DOP_DELTA,
  CL_FWHM,        /* XLine: fixed FWHM */
  CL_D_COUNT
};

#endif /* BAYSAR_LAYOUT_H */
