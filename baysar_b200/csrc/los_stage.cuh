// K1 — line-of-sight stage (float64).  One warp per theta; LOS points across lanes; warp-shuffle reductions.
//
// Replaces, per theta (reference file:line):
//   PlasmaLine.update_plasma_state            baysar/plasmas.py:642-694   theta -> ne, Te, main-ion density
//   EsymmtricProfile.__call__                 baysar/lineshapes.py:779-817
//   PlasmaLine.calc_total_impurity_electrons  baysar/plasmas.py:876-889   (concstar)
//   PlasmaLine.update_neutral_profile         baysar/plasmas.py:696-741   (scd only; acd/plt/prb feed side outputs)
//   BalmerHydrogenLine.__call__ (LOS part)    baysar/linemodels.py:758-836
//   ADAS406Lines.__call__ / get_tec (LOS part) baysar/linemodels.py:386-426, :445-494
//   XLine.__call__ (amplitude)                baysar/linemodels.py:228-230
//   default priors                            baysar/priors.py:14-101, :509-556, :646-667, :1046-1089
#pragma once
#include "device_common.cuh"

struct LosArgs {
  const double* theta;  // [N][n_params]
  int64_t n;
  double* lines;       // [N][n_ulines][BAYSAR_LINE_STRIDE]
  double* components;  // [N][n_chords + n_priors]; this kernel fills the prior columns
  double* hdr;         // [N][BAYSAR_HDR_STRIDE]
  double* profiles;    // optional [N][4][L]
  uint32_t* status;    // [N] (overwritten)
};

// bytes of dynamic shared memory one warp needs for L line-of-sight points
__host__ __device__ static inline size_t los_smem_per_warp(int L, int n_params, int n_species) {
  return (size_t)L * 10 * sizeof(double) + (size_t)(L + (L & 1)) * sizeof(int) +
         (size_t)(2 * n_params + 3 * n_species) * sizeof(double);
}

// exchange slots of a team of TW warps: four sums for each of a warp's 32 lanes (the ADAS line groups)
__host__ __device__ static inline size_t los_team_xch_bytes(int tw) { return (size_t)tw * 32 * 4 * sizeof(double); }

// Closed-form profile functions (layout.h BaysarProfileFn; reference lineshapes.py:483-591, :855-929).  th / p10: the
// function's slice of theta and of 10^theta; aux / aux_set: PROF_*_AUX(_SET); linked: the centre a DoubleSlabPlasma's
// density hands to its temperature.
BAYSAR_DEV double profile_fn(int fn, double x, const double* th, const double* p10, double aux, double aux_set, double linked) {
  switch (fn) {
    case PF_EXP_DECAY:
      return clip_lo(p10[0] * pow(1.0 + p10[1], -x), 0.01);
    case PF_DOUBLE_EXP_DECAY: {
      const double base = (1.0 + p10[1]) * exp(p10[3] * tanh(th[4] * (x - th[2])));
      return clip_lo(p10[0] * pow(base, th[2] - x), 0.01);
    }
    case PF_GAUSSIAN: {
      const double c = aux_set != 0.0 ? aux : th[2];
      const double z = (x - c) / (p10[1] * 0.42466090014400953);
      return clip_lo(p10[0] * exp(-0.5 * (z * z)), 0.01);
    }
    case PF_FLAT:
      return p10[0] + 0.0;
    case PF_LEFT_TOP_HAT:
      return x < th[1] + aux ? p10[0] : 1e10;
    case PF_POISSON:  // unscaled: nu^x e^-nu / x! (scipy's factorial of a float is Gamma(x + 1), 0 for x < 0)
      return pow(p10[1], x) * (x < 0.0 ? 0.0 : exp(-p10[1]) / tgamma(x + 1.0));
    default: {  // PF_LEFT_RIGHT_TOP_HAT
      const double c = aux_set == 1.0 ? th[2] : (aux_set == 2.0 ? aux : linked);
      return x < c ? p10[0] : (c < x ? p10[1] : 0.0);
    }
  }
}

// TW warps cooperate on one theta.  TW = 1 is the warp-per-theta form (reductions are warp shuffles); TW = 4 runs one CTA
// per theta with the line-of-sight points spread over 128 threads, for long lines of sight whose per-point cache
// (128 bytes per point) would otherwise leave a single warp per SM (L = 1000: 128 KB per theta).
template <int TW>
struct LosTeam {
  double* xch;  // exchange slots in shared memory (TW > 1): los_team_xch_bytes(TW)
  BAYSAR_DEV void sync() const {
    if (TW == 1) __syncwarp();
    else __syncthreads();
  }
  template <class Combine>
  BAYSAR_DEV double reduce(double v, Combine comb) const {  // v: already reduced over the warp
    if (TW == 1) return v;
    __syncthreads();  // the slots' previous use
    if ((threadIdx.x & 31) == 0) xch[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = xch[0];
#pragma unroll
    for (int i = 1; i < TW; ++i) r = comb(r, xch[i]);
    return r;
  }
  // NV sums with one exchange (TW > 1: a single pair of barriers instead of one pair per sum); xch holds at least 8 TW doubles
  template <int NV>
  BAYSAR_DEV void sum_n(double* v) const {
    static_assert(NV <= 8, "exchange buffer holds eight sums");
#pragma unroll
    for (int q = 0; q < NV; ++q) v[q] = warp_sum(v[q]);
    if (TW == 1) return;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
      for (int q = 0; q < NV; ++q) xch[q * TW + (threadIdx.x >> 5)] = v[q];
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < NV; ++q) {
      double r = xch[q * TW];
#pragma unroll
      for (int i = 1; i < TW; ++i) r += xch[q * TW + i];
      v[q] = r;
    }
  }
  BAYSAR_DEV double sum(double v) const { return reduce(warp_sum(v), [](double a, double b) { return a + b; }); }
  BAYSAR_DEV double max(double v) const { return reduce(warp_max(v), [](double a, double b) { return a > b ? a : b; }); }
  BAYSAR_DEV uint32_t or_bits(uint32_t v) const {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v |= __shfl_xor_sync(0xffffffffu, v, o);
    if (TW == 1) return v;
    // exact in a double: 32 bits
    return (uint32_t)reduce((double)v, [](double a, double b) { return (double)((uint32_t)a | (uint32_t)b); });
  }
  BAYSAR_DEV int any(int pred) const { return or_bits(pred ? 1u : 0u) != 0u; }
  // largest value, lowest index among equals (numpy.argmax)
  BAYSAR_DEV void argmax(double& best, int& bi) const {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      double ob = __shfl_xor_sync(0xffffffffu, best, o);
      int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if (TW == 1) return;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) { xch[threadIdx.x >> 5] = best; xch[TW + (threadIdx.x >> 5)] = (double)bi; }
    __syncthreads();
    best = xch[0]; bi = (int)xch[TW];
#pragma unroll
    for (int i = 1; i < TW; ++i) {
      const double ob = xch[i];
      const int oi = (int)xch[TW + i];
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
  }
};

// UNR: unroll factor of the per-line loops over line-of-sight points (two points of a lane in flight give the FP64
// dependency chains some instruction-level parallelism; the kernel is latency-bound at 16 warps per SM)
template <int WARPS, int MINB = 1, int UNR = 1, int TW = 1>
__global__ void __launch_bounds__(WARPS * 32, MINB) los_stage_kernel(ModelDev m, LosArgs a) {
  static_assert(TW == 1 || TW == WARPS, "team mode: the whole CTA works on one theta");
  pdl_trigger();
  pdl_wait();  // theta comes from the caller's previous kernel; the workspaces are read by the previous evaluation
#ifdef BAYSAR_CPU_EMU
  unsigned char* los_smem_raw = emu_dynamic_smem();
#else
  extern __shared__ __align__(16) unsigned char los_smem_raw[];
#endif
  // `lane`: index of this thread in its team (TW = 1: the warp lane), TS: threads per team
  constexpr int TS = 32 * TW;
  const int lane = TW == 1 ? (threadIdx.x & 31) : threadIdx.x, warp = TW == 1 ? (threadIdx.x >> 5) : 0;
  const int64_t n = TW == 1 ? (int64_t)blockIdx.x * WARPS + warp : (int64_t)blockIdx.x;
  if (n >= a.n) return;  // whole teams exit together (TW = 1: no block-level barrier is used in this kernel)

  const int64_t* I = m.I;
  const int L = (int)I[I_L], n_params = (int)I[I_N_PARAMS], n_elem = (int)I[I_N_ELEM];
  const int n_species = (int)I[I_N_SPECIES], n_ul = (int)I[I_N_ULINES], n_pr = (int)I[I_N_PRIORS];
  const int n_chords = (int)I[I_N_CHORDS];
  const double* th = a.theta + n * n_params;

  unsigned char* base = los_smem_raw + (size_t)warp * los_smem_per_warp(L, n_params, n_species);
  double* __restrict__ p10 = (double*)base;        // 10^theta[k]: every pow(10, theta) of the reference, one lane per parameter
  double* __restrict__ l10 = p10 + n_params;       // log10(10^theta[k]) (the reference takes log10 of the transformed tau)
  double* __restrict__ spw = l10 + n_params;       // per species: ADAS tau-slice pair index and its two weights
  double* __restrict__ ne_s = spw + 3 * n_species;
  double* __restrict__ te_s = ne_s + L;
  // The rest of the per-point cache is a union (its size bounds the thetas resident on an SM, and with them the warps
  // that hide the latency of this kernel): first the main-ion and neutral densities and the adf15 cell of every point
  // (neutral profile, Balmer lines, the priors on the profiles), then -- those are dead -- the spline cell and basis
  // (ADAS line groups).
  double* __restrict__ un_s = te_s + L;
  double* __restrict__ nm_s = un_s;       // main ion density
  double* __restrict__ n0_s = nm_s + L;   // neutral density profile
  double* __restrict__ ha_s = n0_s + L;   // adf15-grid cell weights
  double* __restrict__ hb_s = ha_s + L;
  double* __restrict__ bx_s = un_s;       // [4][L] spline basis in ne (structure of arrays: conflict-free across lanes)
  double* __restrict__ by_s = bx_s + 4 * L;
  int* __restrict__ ho_s = (int*)(un_s + 8 * L);  // adf15 cell: element offset of its (ne, Te) corner in a table
  int* __restrict__ co_s = ho_s;                  // spline cell: index of coefficient entry (lx - 3, ly - 3) in a group's table
  const LosTablesDev& lt = m.lt;
  LosTeam<TW> team{reinterpret_cast<double*>(los_smem_raw + (size_t)(TW == 1 ? WARPS : 1) * los_smem_per_warp(L, n_params, n_species))};

  uint32_t st = 0;
  for (int k = lane; k < n_params; k += TS) {
    double v = th[k];
    if (v != v) st |= 1u << 0;
    if (v - v != 0.0 && v == v) st |= 1u << 1;  // +-inf
    const double p = exp10(v);  // numpy's 10 ** theta; pow(10, v) is three times the instructions for the same ~1 ulp
    p10[k] = p;
    l10[k] = log10(p);
  }
  team.sync();

  // ---- electron density / temperature profiles
  const int ne_idx = (int)I[I_NE_IDX], te_idx = (int)I[I_TE_IDX];
  const int profile_kind = (int)I[I_PROFILE_KIND];
  double nemax = -1.7976931348623157e308;
  if (profile_kind == 0) {
    // EsymmtricProfile (lineshapes.py:779-817)
    double ne_A, ne_c, ne_f, ne_floor;
    if (m.prof_d[PROF_NE_CENTRE_FIXED] != 0.0) {
      ne_A = th[ne_idx]; ne_f = th[ne_idx + 1]; ne_floor = th[ne_idx + 2]; ne_c = m.prof_d[PROF_NE_CENTRE];
    } else {
      ne_A = th[ne_idx]; ne_c = th[ne_idx + 1]; ne_f = th[ne_idx + 2]; ne_floor = m.prof_d[PROF_NE_FLOOR];
    }
    (void)ne_A;
    const double te_f = th[te_idx + 1], te_floor = th[te_idx + 2], te_c = m.prof_d[PROF_TE_CENTRE];
    const double ne_pk = p10[ne_idx], te_pk = p10[te_idx];
    const bool esym_gauss = m.prof_d[PROF_ESYM_GAUSSIAN] != 0.0;  // ptype="gaussian" (lineshapes.py:769-776)
    for (int l = lane; l < L; l += TS) {
      // sigma = 0.2 f + f / (1 + e), e = exp(-dx): dx / sigma = dx (1 + e) / (f (1.2 + 0.2 e)), one division instead of two
      double x = m.los_x[l];
      double dx = x - ne_c;
      double e = exp_fast(-dx);
      double r = (dx * (1.0 + e)) / (ne_f * fma(0.2, e, 1.2));
      double v = clip_lo(esym_gauss ? ne_pk * exp(-0.5 * (r * r)) : ne_pk / fma(r, r, 1.0), ne_floor);
      ne_s[l] = v;
      nemax = v > nemax ? v : nemax;
      dx = x - te_c;
      e = exp_fast(-dx);
      r = (dx * (1.0 + e)) / (te_f * fma(0.2, e, 1.2));
      te_s[l] = clip_lo(esym_gauss ? te_pk * exp(-0.5 * (r * r)) : te_pk / fma(r, r, 1.0), te_floor);
    }
  } else if (profile_kind == 1) {
    // BowmanTeeNe / BowmanTeeTe (lineshapes.py:332-376): A (1 + |((x - x0) / sigma)|^q / nu)^(-(nu + 1) / 2) + b,
    // sigma = sigma0 exp(f tanh(k (x - x0))); the temperature profile is centred on 0
    const int bgf = m.prof_d[PROF_BACKGROUND] != 0.0;
    const double nA = p10[ne_idx], nx0 = th[ne_idx + 1], ns0 = th[ne_idx + 2], nq = th[ne_idx + 3], nnu = th[ne_idx + 4],
                 nk = th[ne_idx + 5], nf = th[ne_idx + 6], nb = bgf ? p10[ne_idx + 7] : 0.0;
    const double tA = p10[te_idx], ts0 = th[te_idx + 1], tq = th[te_idx + 2], tnu = th[te_idx + 3], tk = th[te_idx + 4],
                 tf = th[te_idx + 5], tb = bgf ? p10[te_idx + 6] : 0.0;
    for (int l = lane; l < L; l += TS) {
      const double x = m.los_x[l];
      double sg = ns0 * exp(nf * tanh(nk * (x - nx0)));
      double z = pow(fabs((x - nx0) / sg), nq);
      const double v = nA * pow(1.0 + z / nnu, -(nnu + 1.0) * 0.5) + nb;
      ne_s[l] = v;
      nemax = v > nemax ? v : nemax;
      sg = ts0 * exp(tf * tanh(tk * (x - 0.0)));
      z = pow(fabs((x - 0.0) / sg), tq);
      te_s[l] = tA * pow(1.0 + z / tnu, -(tnu + 1.0) * 0.5) + tb;
    }
  } else if (profile_kind == 3) {
    // closed-form pair (SimplePlasma, LessSimplePlasma, SimpleGaussianPlasma, SlabPlasma, DoubleSlabPlasma)
    const int ne_fn = (int)I[I_NE_FN], te_fn = (int)I[I_TE_FN];
    const double linked = th[ne_idx + 2 < n_params ? ne_idx + 2 : ne_idx];  // DoubleSlabPlasma: the density's centre
    for (int l = lane; l < L; l += TS) {
      const double x = m.los_x[l];
      const double v = profile_fn(ne_fn, x, th + ne_idx, p10 + ne_idx, m.prof_d[PROF_NE_AUX], m.prof_d[PROF_NE_AUX_SET], 0.0);
      ne_s[l] = v;
      nemax = v > nemax ? v : nemax;
      te_s[l] = profile_fn(te_fn, x, th + te_idx, p10 + te_idx, m.prof_d[PROF_TE_AUX], m.prof_d[PROF_TE_AUX_SET], linked);
    }
    if (ne_fn == PF_POISSON || te_fn == PF_POISSON) {
      // Poisson (lineshapes.py:544-550): peak *= a / peak.max() -- one more pass once the maximum over the line of sight
      // is known
      for (int pass = 0; pass < 2; ++pass) {
        if ((pass == 0 ? ne_fn : te_fn) != PF_POISSON) continue;  // block-uniform
        double* prof = pass == 0 ? ne_s : te_s;
        double pmax = -1.7976931348623157e308;
        for (int l = lane; l < L; l += TS) pmax = prof[l] > pmax ? prof[l] : pmax;
        pmax = team.max(pmax);
        const double scale = p10[pass == 0 ? ne_idx : te_idx] / pmax;
        if (pass == 0) nemax = -1.7976931348623157e308;
        for (int l = lane; l < L; l += TS) {
          const double v = prof[l] * scale;
          prof[l] = v;
          if (pass == 0) nemax = v > nemax ? v : nemax;
        }
      }
    }
  } else {
    // MeshLine (lineshapes.py:282-301): values at the knots (10^theta if log), scipy interp1d "linear" with
    // extrapolation: slope (x - x_lo) + y_lo on the segment [lo, lo + 1] fixed per LOS point at model build
    const int n_ne = (int)I[I_MESH_NE_N];
    const int ne_log = m.prof_d[PROF_MESH_NE_LOG] != 0.0, te_log = m.prof_d[PROF_MESH_TE_LOG] != 0.0;
    const double* xk_ne = m.mesh_x;
    const double* xk_te = m.mesh_x + n_ne;
    for (int l = lane; l < L; l += TS) {
      const double x = m.los_x[l];
      int lo = m.mesh_lo[l];
      double y0 = ne_log ? p10[ne_idx + lo] : th[ne_idx + lo], y1 = ne_log ? p10[ne_idx + lo + 1] : th[ne_idx + lo + 1];
      double slope = (y1 - y0) / (xk_ne[lo + 1] - xk_ne[lo]);
      const double v = __dadd_rn(__dmul_rn(slope, x - xk_ne[lo]), y0);
      ne_s[l] = v;
      nemax = v > nemax ? v : nemax;
      lo = m.mesh_lo[L + l];
      y0 = te_log ? p10[te_idx + lo] : th[te_idx + lo]; y1 = te_log ? p10[te_idx + lo + 1] : th[te_idx + lo + 1];
      slope = (y1 - y0) / (xk_te[lo + 1] - xk_te[lo]);
      te_s[l] = __dadd_rn(__dmul_rn(slope, x - xk_te[lo]), y0);
    }
  }
  nemax = team.max(nemax);
  double inv_nemax = 1.0 / nemax;

  // ---- impurity electrons, main ion density (plasmas.py:876-889, :673)
  const int reverse_en = (int)I[I_REVERSE_EN];
  for (int l = lane; l < L; l += TS) {
    double ne = ne_s[l], tot = 0.0;
    for (int e = 0; e < n_elem; ++e) {
      double dens = p10[m.elem_dens_idx[e]];
      double conc = (dens * inv_nemax) * ne;
      tot += (conc * inv_nemax) * ne;
    }
    nm_s[l] = reverse_en ? nan_to_num(tot) : ne - nan_to_num(tot);
  }
  if (reverse_en) {
    // plasma.reverse_electroneutrality (plasmas.py:667-671): the sampled profile is the main-ion density and the impurity
    // electrons (formed above from it, :666) are added to the electron density every later stage sees
    double mx = -1.7976931348623157e308;
    for (int l = lane; l < L; l += TS) {
      const double n1 = ne_s[l], ne = n1 + nm_s[l];  // (the loop above parked the impurity electrons there)
      nm_s[l] = n1;
      ne_s[l] = ne;
      mx = ne > mx ? ne : mx;
    }
    nemax = team.max(mx);
    inv_nemax = 1.0 / nemax;
  }

  // ---- impurity tau must index the ADAS tau grid (plasmas.py:1410-1429 raise; linemodels.py:454-467)
  const int n_tau = (int)I[I_N_TAU];
  for (int s = lane; s < n_species; s += TS) {
    const int* sp = m.sp_i + s * SP_I_COUNT;
    if (sp[SP_IS_IMPURITY]) {
      double lt = l10[sp[SP_TAU_IDX]];
      if (!(lt > m.tau_grid[0] && lt <= m.tau_grid[n_tau - 1])) st |= 1u << 2;
      // tau slices (j - 1, j) and their weights (linemodels.py:450-451, :469-473), once per species, lanes side by side:
      // every line of the species reads them back instead of repeating the search and the two divisions
      int j = lower_bound(m.tau_grid, n_tau, lt);
      if (j >= n_tau) j = n_tau - 1;  // flagged above
      int i = j - 1;
      if (i < 0) i = n_tau - 1;  // python's negative index (flagged above)
      const double ti_ = m.tau_grid[i], tj_ = m.tau_grid[j];
      spw[3 * s] = (double)(j > 0 ? j - 1 : 0);  // the out-of-grid case j = 0 reads pair 0
      spw[3 * s + 1] = 1.0 + (lt - ti_) / (ti_ - tj_);
      spw[3 * s + 2] = 1.0 + (tj_ - lt) / (ti_ - tj_);
    }
  }

  // ---- neutral hydrogen profile (plasmas.py:696-741)
  const int has_h = (int)I[I_HAS_HYDROGEN];
  if (has_h) {
    const int* sp = m.sp_i + (int)I[I_H_SPECIES] * SP_I_COUNT;
    const int dens_idx = sp[SP_DENS_IDX];
    const double tau_h = p10[sp[SP_TAU_IDX]];
    const double n0_sampled = dens_idx >= 0 ? p10[dens_idx] : 0.0;
    const int nne = (int)I[I_A11_NNE], nte = (int)I[I_A11_NTE];
    const int recombining = (int)I[I_RECOMBINING];
    for (int l = lane; l < L; l += TS) {
      double ne = ne_s[l], te = te_s[l];
      int ci, cj; double ca, cb;
      linear_cell_fast(m.a11_ne, nne, ne, lt, AX_A11_NE, ci, ca);
      linear_cell_fast(m.a11_te, nte, te, lt, AX_A11_TE, cj, cb);
      double scd = bilinear(m.a11_scd, nte, ci, cj, ca, cb);
      double n0 = dens_idx >= 0 ? n0_sampled : nm_s[l];
      if (recombining) n0 += nm_s[l] * ne * bilinear(m.a11_acd, nte, ci, cj, ca, cb) * tau_h;  // plasmas.py:735-737
      n0 = n0 - n0 * ne * scd * tau_h;
      n0 = clip_lo(n0, 1.0);
      if (n0 != n0) st |= 1u << 3;
      n0_s[l] = n0;
    }
    // adf15 grid cells, shared by every Balmer line
    const int hne = (int)I[I_H_NNE], hte = (int)I[I_H_NTE];
    for (int l = lane; l < L; l += TS) {
      int ci, cj; double ca, cb;
      linear_cell_fast(m.h_ne, hne, ne_s[l], lt, AX_H_NE, ci, ca);
      linear_cell_fast(m.h_te, hte, te_s[l], lt, AX_H_TE, cj, cb);
      ho_s[l] = ci * hte + cj; ha_s[l] = ca; hb_s[l] = cb;
    }
  } else {
    for (int l = lane; l < L; l += TS) n0_s[l] = 0.0;
  }

  team.sync();

  if (a.profiles) {
    double* p = a.profiles + n * 4 * (int64_t)L;
    for (int l = lane; l < L; l += TS) {
      p[l] = ne_s[l]; p[L + l] = te_s[l]; p[2 * L + l] = nm_s[l]; p[3 * L + l] = n0_s[l];
    }
  }

  // ---- unique lines
  const double four_pi = 4.0 * 3.141592653589793;
  double* lines_n = a.lines + n * (int64_t)n_ul * BAYSAR_LINE_STRIDE;
  const int cold_neutrals = (int)I[I_COLD_NEUTRALS], cold_ions = (int)I[I_COLD_IONS];
  const int thermalised = (int)I[I_THERMALISED];
  for (int u = 0; u < n_ul; ++u) {
    const int* ul = m.ul_i + u * UL_I_COUNT;
    const int kind = ul[UL_KIND];
    double* rec = lines_n + u * BAYSAR_LINE_STRIDE;
    if (kind == 0) {  // Balmer (linemodels.py:782-836)
      const int hne = (int)I[I_H_NNE], hte = (int)I[I_H_NTE];
      const double* texc = m.h_tab + (int64_t)ul[UL_TAB_A] * hne * hte;
      const double* trec = m.h_tab + (int64_t)ul[UL_TAB_B] * hne * hte;
      // s[8..11] (the sums over recombination + excitation) are formed from the partial sums after the reduction
      double s[12];
#pragma unroll
      for (int q = 0; q < 8; ++q) s[q] = 0.0;
#pragma unroll UNR
      for (int l = lane; l < L; l += TS) {
        double ne = ne_s[l], te = te_s[l], w = m.los_w[l];
        double le, lr;
        bilinear2(texc, trec, hte, 0, ho_s[l], ha_s[l], hb_s[l], le, lr);
        double pe = exp_fast(le), pr = exp_fast(lr);
        if (pe != pe || pr != pr) st |= 1u << 4;
        double rp = clip_lo(nm_s[l], 1.0) * ne * pr;
        double ep = n0_s[l] * ne * pe;
        s[0] += rp; s[1] += ep; s[2] += w * rp; s[3] += w * ep;
        s[4] += rp * ne; s[5] += rp * te; s[6] += ep * ne; s[7] += ep * te;
      }
      team.template sum_n<8>(s);
      // The line's scalars (seven divisions, a square root) are uniform across the warp, so forming them here would run
      // them once per line, one line after the other; the raw sums wait in the record instead and one lane per line
      // forms all lines' scalars side by side after the loop.
      if (lane == 2) {
#pragma unroll
        for (int q = 0; q < 8; ++q) rec[q] = s[q];
      }
    } else if (kind == 1) {  // ADAS406: evaluated per group of lines below
    } else {  // XLine (linemodels.py:228-230)
      if (lane == 0) {
        rec[0] = p10[ul[UL_TAB_A]];
        for (int q = 1; q < BAYSAR_LINE_STRIDE; ++q) rec[q] = 0.0;
      }
    }
  }

  // ---- priors, in posterior_components order (posterior.py:108-161)
  // Two passes: the priors on the profiles run while the densities are still cached (before the spline basis takes their
  // place), the two that read line scalars (StarkToPeak, ChargeStateOrder) after the lines are done.
  double* comp = a.components + n * (int64_t)(n_chords + n_pr) + n_chords;
  auto priors_pass = [&](const bool line_priors) {
  for (int k = 0; k < n_pr; ++k) {
    const int* pi = m.pr_i + k * PR_I_COUNT;
    const double* pd = m.pr_d + k * PR_D_COUNT;
    const int kind = pi[PR_KIND];
    if ((kind == PRIOR_STARK_TO_PEAK || kind == PRIOR_CHARGE_STATE_ORDER) != line_priors) continue;
    double val = 0.0;
    if (kind == PRIOR_CURVATURE) {  // priors.py:104-127: curvature of the raw density-knot parameters
      // grad = numpy.gradient (unit spacing): central differences inside, one-sided at the two ends
      const int nk = (int)I[I_MESH_NE_N];
      double* g1 = by_s;  // scratch (the second half of the union: free while the densities are cached)
      double vmax = -1.7976931348623157e308;
      for (int q = 0; q < nk; ++q) vmax = th[ne_idx + q] > vmax ? th[ne_idx + q] : vmax;
      team.sync();
      for (int q = lane; q < nk; q += TS) {
        const double c = th[ne_idx + q] / vmax;
        if (nk == 1) g1[q] = 0.0;
        else if (q == 0) g1[q] = th[ne_idx + 1] / vmax - c;
        else if (q == nk - 1) g1[q] = c - th[ne_idx + nk - 2] / vmax;
        else g1[q] = (th[ne_idx + q + 1] / vmax - th[ne_idx + q - 1] / vmax) / 2.0;
      }
      team.sync();
      double cv = 0.0;
      for (int q = 0; q < nk; ++q) {  // sequential sum like Python's sum()
        double g2;
        if (nk == 1) g2 = 0.0;
        else if (q == 0) g2 = g1[1] - g1[0];
        else if (q == nk - 1) g2 = g1[nk - 1] - g1[nk - 2];
        else g2 = (g1[q + 1] - g1[q - 1]) / 2.0;
        const double den = 1.0 + g1[q] * g1[q];
        cv += (g2 * g2) / (den * den * den);
      }
      val = -cv * pd[PR_D0];
      team.sync();
    } else if (kind == PRIOR_STATIC_PRESSURE) {  // priors.py:51-64
      double s = 0.0;
      for (int l = lane; l < L; l += TS) {
        double pe = clip_lo(te_s[l], 0.01) * clip_lo(ne_s[l], 1.0);
        s += low_pass(pe * (1.0 / 5e15), 1.0, 0.1);
      }
      val = team.sum(s);
    } else if (kind == PRIOR_NEUTRAL_FRACTION) {  // priors.py:83-101
      double s = 0.0;
      for (int l = lane; l < L; l += TS) {
        double f0 = clip_lo(n0_s[l], 1.0) / ne_s[l];
        s += high_pass(f0 * (1.0 / 5e-3), 1.0, 0.1);
      }
      val = team.sum(s);
    } else if (kind == PRIOR_STARK_TO_PEAK) {  // priors.py:646-667
      const double* r = lines_n + pi[PR_I0] * BAYSAR_LINE_STRIDE;
      double exc = high_pass(r[4] / nemax, 0.65, 0.075) * (1 - r[7]);
      double rc = high_pass(r[2] / nemax, 0.65, 0.075) * r[7];
      val = exc + rc;
    } else if (kind == PRIOR_ANTIPROTON) {  // priors.py:14-24
      double s = 0.0; int any_neg = 0;
      for (int l = lane; l < L; l += TS) {
        double ni = nm_s[l];
        if (ni < 0) { any_neg = 1; double t = ni * (1.0 / 1e11); s += t * t; }
      }
      s = team.sum(s);
      any_neg = team.any(any_neg);
      val = any_neg ? -0.5 * s : 0.0;
    } else if (kind == PRIOR_MAIN_ION_FRACTION) {  // priors.py:27-48
      double best = -1.7976931348623157e308; int bi = 0x7fffffff;
      for (int l = lane; l < L; l += TS) {
        double v = nm_s[l];
        if (v > best) { best = v; bi = l; }
      }
      team.argmax(best, bi);
      if (bi >= L) bi = 0;
      const double inv_ref = 1.0 / clip_lo(ne_s[bi], 1.0);
      double s = 0.0;
      for (int l = lane; l < L; l += TS) {
        double nec = clip_lo(ne_s[l], 1.0);
        double factor = clip_lo(nec * inv_ref, 0.3);
        s += high_pass(nm_s[l] / nec, factor * 0.8, 0.1);
      }
      val = team.sum(s);
    } else if (kind == PRIOR_CHARGE_STATE_ORDER) {  // priors.py:509-556
      double cost = 0.0;
      for (int e = 0; e < n_elem; ++e) {
        for (int q = m.cso_off[e]; q + 1 < m.cso_off[e + 1]; ++q) {
          double t0 = lines_n[m.cso_ul[q] * BAYSAR_LINE_STRIDE + 3];
          double t1 = lines_n[m.cso_ul[q + 1] * BAYSAR_LINE_STRIDE + 3];
          cost += high_pass(t1 - t0, 2.0, 0.2);
        }
      }
      val = cost;
    } else if (kind == PRIOR_WAVELENGTH_CAL) {  // priors.py:1046-1060 (always calwave_0)
      double t0 = (pd[PR_D0] - th[pi[PR_I0]]) / pd[PR_D2];
      double t1 = (pd[PR_D1] - th[pi[PR_I0] + 1]) / pd[PR_D3];
      val = -0.5 * (t0 * t0) + -0.5 * (t1 * t1);
    } else if (kind == PRIOR_INTENSITY_CAL) {  // priors.py:1063-1076
      double t0 = (1.0 - p10[pi[PR_I0]]) / 0.2;
      val = -0.5 * (t0 * t0);
    } else if (kind == PRIOR_GAUSSIAN_IF) {  // priors.py:1079-1089
      double t0 = (p10[pi[PR_I0]] - 0.8) / 0.1;
      val = -0.5 * (t0 * t0);
    } else if (kind == PRIOR_TDV) {  // ElectronDensityTDVCost, priors.py:67-80: d0 threshold, d1 sigma
      double s = 0.0;
      for (int l = lane; l + 1 < L; l += TS) {
        const double ne_tdv = fabs(ne_s[l + 1] - ne_s[l]), ni_tdv = fabs(nm_s[l + 1] - nm_s[l]);
        s += (ne_tdv - ni_tdv) / ni_tdv;
      }
      val = low_pass(team.sum(s) / pd[PR_D0], 1.0, pd[PR_D1]);
    } else if (kind == PRIOR_PEAK_TE || kind == PRIOR_TE_MIN || kind == PRIOR_TE_TAU) {
      double tmax = -1.7976931348623157e308, tmin = 1.7976931348623157e308;
      for (int l = lane; l < L; l += TS) {
        tmax = te_s[l] > tmax ? te_s[l] : tmax;
        tmin = te_s[l] < tmin ? te_s[l] : tmin;
      }
      tmax = team.max(tmax);
      tmin = -team.max(-tmin);
      if (kind == PRIOR_PEAK_TE) {  // priors.py:241-251: d0 te_min, d1 te_err
        val = low_pass(tmax, pd[PR_D0], pd[PR_D1]);
      } else if (kind == PRIOR_TE_MIN) {  // priors.py:670-680: d0 ratio, d1 sigma
        val = high_pass(tmax / tmin, pd[PR_D0], pd[PR_D1]);
      } else {  // TeTauPrior, priors.py:696-727: log10 tau of every impurity species inside te_max-dependent bounds
        const double b0 = tmax > 5.0 ? -5.0 : (tmax < 3.0 ? 3.0 : -3.0), b1 = tmax > 5.0 ? 0.0 : 7.0;
        double cost = 0.0;
        for (int q = 0; q < pi[PR_I1]; ++q) {  // sequential, like the reference's loop
          const double lt = l10[m.pr_aux[pi[PR_I0] + q]];
          cost += 0.0 + high_pass(lt, b0, pd[PR_D0]) + low_pass(lt, b1, pd[PR_D0]);
        }
        val = cost;
      }
    } else if (kind == PRIOR_TAU) {  // TauPrior, priors.py:254-301: -diff(log10 tau) of consecutive ions; d0 mean, d1 sigma
      double cost = 0.0;
      for (int q = 0; q < pi[PR_I1]; ++q) {
        const int* pr = m.pr_aux + pi[PR_I0] + 2 * q;
        cost += low_pass(-(l10[pr[1]] - l10[pr[0]]), pd[PR_D0], pd[PR_D1]);
      }
      val = cost;
    } else if (kind == PRIOR_WALL_CONDITIONS) {  // priors.py:432-456: d0 te_min, d1 ne_min, d2 te_err, d3 ne_err
      double cost = 0.0;
      cost += low_pass(te_s[0] / pd[PR_D0], 1.0, pd[PR_D2]);
      cost += low_pass(te_s[L - 1] / pd[PR_D0], 1.0, pd[PR_D2]);
      cost += low_pass(ne_s[0] / pd[PR_D1], 1.0, pd[PR_D3]);
      cost += low_pass(ne_s[L - 1] / pd[PR_D1], 1.0, pd[PR_D3]);
      val = cost;
    } else if (kind == PRIOR_SIMPLE_WALL) {  // priors.py:459-481
      val = 0.0 + low_pass(te_s[L - 1], pd[PR_D0], pd[PR_D2]) + low_pass(ne_s[L - 1], pd[PR_D1], pd[PR_D3]);
    } else if (kind == PRIOR_IMPURITY_TAU) {  // priors.py:395-429: raw tau parameters of equal charge states; d0 sigma
      double cost = 0.0;
      int group = -1;
      double gsum = 0.0;
      for (int q = 0; q < pi[PR_I1]; ++q) {  // aux triples (group, idx_a, idx_b): -0.5 * sum over a group, added per group
        const int* pr = m.pr_aux + pi[PR_I0] + 3 * q;
        if (pr[0] != group) {
          if (group >= 0) cost += -0.5 * gsum;
          group = pr[0];
          gsum = 0.0;
        }
        const double t = (th[pr[2]] - th[pr[1]]) / pd[PR_D0];
        gsum += t * t;
      }
      if (group >= 0) cost += -0.5 * gsum;
      val = cost;
    } else if (kind == PRIOR_BOWMAN_TEE) {  // priors.py:484-506: Te width parameter not above the ne one; d0 sigma_err
      val = 0.0 + low_pass(th[te_idx + 1] - th[ne_idx + 2], 0.0, pd[PR_D0]);
    }
    if (lane == 0) comp[k] = val;
  }
  };

  priors_pass(false);
  team.sync();  // the densities are dead: the spline cell and basis take their place

  // ---- bicubic B-spline basis per LOS point, shared by every impurity line and tau slice
  const int n_tec = (int)I[I_N_TEC];
  const int snx = (int)I[I_SPL_NX], sny = (int)I[I_SPL_NY];
  if (n_tec > 0) {
    for (int l = lane; l < L; l += TS) {
      int lx, ly; double hx[4], hy[4];
      bspline_basis3_fast(m.spl_tx, snx, ne_s[l], lt, AX_SPL_X, lx, hx);
      bspline_basis3_fast(m.spl_ty, sny, te_s[l], lt, AX_SPL_Y, ly, hy);
      co_s[l] = (lx - 3) * (sny - 4) + (ly - 3);
#pragma unroll
      for (int q = 0; q < 4; ++q) { bx_s[q * L + l] = hx[q]; by_s[q * L + l] = hy[q]; }
    }
  }
  team.sync();

  // ---- ADAS406 lines (linemodels.py:386-426, :445-494), one group at a time: the lines of one species share the tau
  // slices, the density and, per line-of-sight point, the spline cell and basis.  The lanes of a warp are dealt
  // (line k, point) pairs, K = min(32, lines) lines x P = 32 / K consecutive points per trip; the group's coefficients
  // lie side by side (layout.h SEC_SPL_C), so the lanes of one cell read neighbouring bytes of the same cache lines
  // and a gather costs one wavefront per distinct cell among P points, for all K lines at once -- the stage is bound
  // by the L1 data pipe, and with one line at a time every line paid those wavefronts again.
  {
    const int n_groups = (int)I[I_N_GROUPS];
    const int ncx = snx - 4, ncy = sny - 4;
    const int wl = threadIdx.x & 31, wi = TW == 1 ? 0 : (int)(threadIdx.x >> 5);
    for (int g = 0; g < n_groups; ++g) {
      const int* gr = m.gr_i + g * GR_I_COUNT;
      const int n_k = gr[GR_NK];
      const int* sp = m.sp_i + gr[GR_SPECIES] * SP_I_COUNT;
      // n0 = dens * (ne / ne_max) and the concentration weight n0 / ne = dens / ne_max (linemodels.py:392-394, :411):
      // the two float64 divisions per point and line become one reciprocal per theta (relative difference ~1e-16)
      const double dens_over_max = p10[sp[SP_DENS_IDX]] * inv_nemax;
      const double* w3 = spw + 3 * gr[GR_SPECIES];
      const int pair = (int)w3[0];
      const double i_w = w3[1], j_w = w3[2];
      const size_t entry = (size_t)n_k * 32, row_bytes = (size_t)ncy * entry;
      const char* gbase = reinterpret_cast<const char*>(m.spl_c) +
                          ((size_t)gr[GR_COEF_OFF] + (size_t)pair * ncx * ncy * n_k) * 32;
      for (int k0 = 0; k0 < n_k; k0 += 32) {
        const int K = n_k - k0 < 32 ? n_k - k0 : 32, P = 32 / K;
        const int k = wl / P, pl = wl - k * P;
        const bool active = k < K;
        const char* lbase = gbase + (size_t)(k0 + (active ? k : 0)) * 32;
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s4 = 0.0;
        for (int p0 = wi * P; p0 < L; p0 += TW * P) {
          const int l = p0 + pl;
          if (active && l < L) {
            const double ne = ne_s[l], te = te_s[l];
            double vi, vj;
            bspline_eval_group(lbase + (size_t)co_s[l] * entry, entry, row_bytes, bx_s + l, by_s + l, L, vi, vj);
            double tec = exp_fast(vi) * i_w + exp_fast(vj) * j_w;
            if (!(fabs(tec) <= 1.7976931348623157e308)) tec = nan_to_num(tec);
            const double p = clip_lo((dens_over_max * ne) * tec, 1.0);
            s0 += p; s1 += m.los_w[l] * p; s2 += p * ne; s4 += p * te;
          }
        }
        // sums over the P lanes of every line: a shuffle tree that stays inside the segment
        for (int o = 1; o < P; o <<= 1) {
          const double t0 = __shfl_down_sync(0xffffffffu, s0, o), t1 = __shfl_down_sync(0xffffffffu, s1, o);
          const double t2 = __shfl_down_sync(0xffffffffu, s2, o), t4 = __shfl_down_sync(0xffffffffu, s4, o);
          if (pl + o < P) { s0 += t0; s1 += t1; s2 += t2; s4 += t4; }
        }
        // raw sums into the line records; the scalars follow below, one lane per line (as for the Balmer lines)
        if (TW == 1) {
          if (active && pl == 0) {
            double* rec = lines_n + m.gr_ul[gr[GR_FIRST] + k0 + k] * BAYSAR_LINE_STRIDE;
            rec[0] = s0; rec[1] = s1; rec[2] = s2; rec[3] = s0 * dens_over_max; rec[4] = s4;
          }
        } else {
          __syncthreads();  // the exchange slots' previous use
          if (active && pl == 0) {
            double* x = team.xch + (wi * 32 + k) * 4;
            x[0] = s0; x[1] = s1; x[2] = s2; x[3] = s4;
          }
          __syncthreads();
          if ((int)threadIdx.x < K) {
            double t[4] = {0.0, 0.0, 0.0, 0.0};
            for (int w = 0; w < TW; ++w) {
              const double* x = team.xch + (w * 32 + (int)threadIdx.x) * 4;
#pragma unroll
              for (int q = 0; q < 4; ++q) t[q] += x[q];
            }
            double* rec = lines_n + m.gr_ul[gr[GR_FIRST] + k0 + (int)threadIdx.x] * BAYSAR_LINE_STRIDE;
            rec[0] = t[0]; rec[1] = t[1]; rec[2] = t[2]; rec[3] = t[0] * dens_over_max; rec[4] = t[3];
          }
        }
      }
    }
  }
  team.sync();
  // line scalars from the raw sums, one lane per line (Balmer: linemodels.py:796-836)
  for (int u = lane; u < n_ul; u += TS) {
    const int* ul = m.ul_i + u * UL_I_COUNT;
    double* rec = lines_n + u * BAYSAR_LINE_STRIDE;
    if (ul[UL_KIND] == 1) {  // ADAS406 (linemodels.py:405-426)
      const int* sp = m.sp_i + ul[UL_SPECIES] * SP_I_COUNT;
      double s[5];
#pragma unroll
      for (int q = 0; q < 5; ++q) s[q] = rec[q];
      const double ems_te = s[4] / s[0];
      double ti;
      if (cold_ions) ti = 0.01;
      else if (thermalised) ti = ems_te;
      else ti = p10[sp[SP_TI_IDX]];
      rec[0] = s[1] / four_pi; rec[1] = s[2] / s[0]; rec[2] = s[3] / s[0]; rec[3] = ems_te; rec[4] = ti;
      for (int q = 5; q < BAYSAR_LINE_STRIDE; ++q) rec[q] = 0.0;
      continue;
    }
    if (ul[UL_KIND] != 0) continue;
    double s[12];
#pragma unroll
    for (int q = 0; q < 8; ++q) s[q] = rec[q];
    s[8] = s[0] + s[1]; s[9] = s[2] + s[3]; s[10] = s[4] + s[6]; s[11] = s[5] + s[7];
    double rec_sum = s[2] / four_pi, exc_sum = s[3] / four_pi, ems = s[9] / four_pi;
    double rec_ne = s[4] / s[0], rec_te = s[5] / s[0], exc_ne = s[6] / s[1], exc_te = s[7] / s[1];
    double f_rec = rec_sum / ems;
    double ti;
    const int* sp = m.sp_i + ul[UL_SPECIES] * SP_I_COUNT;
    if (cold_neutrals) ti = 0.01;
    else if (thermalised) ti = (1 - f_rec) * exc_te + f_rec * rec_te;
    else ti = p10[sp[SP_TI_IDX]];
    // shape constants shared by every chord that sees this line: Doppler-shifted centre (linemodels.py:239-241) and
    // Doppler sigma (:298, lineshapes.py:128); the pow-heavy Stark half-widths follow, one lane per (line,
    // recombination / excitation)
    const double* ud = m.ul_d + u * UL_D_COUNT;
    rec[0] = rec_sum; rec[1] = exc_sum; rec[2] = rec_ne; rec[3] = rec_te; rec[4] = exc_ne; rec[5] = exc_te;
    rec[6] = ti; rec[7] = f_rec; rec[8] = s[10] / s[8]; rec[9] = s[11] / s[8]; rec[10] = ems;
    const int vi = sp[SP_VEL_IDX];
    const double cwl = vi >= 0 ? ud[UL_CWL0] * ((1e3 * th[vi]) / 299792458.0 + 1.0) : ud[UL_CWL0];
    rec[13] = cwl;
    rec[14] = cwl * 7.715e-5 * sqrt(ti / ud[UL_MASS]) * 0.42466090014400953;
    rec[15] = 0.0;
  }
  team.sync();
  // Stark half-widths gamma^2.5 of every Balmer line (stehle_param, linemodels.py:554-566): lane 2q / 2q+1 take the
  // recombination / excitation width of the q-th Balmer line, so that all the pow() calls run side by side
  for (int idx = lane; idx < 2 * n_ul; idx += TS) {
    const int u = idx >> 1, which = idx & 1;
    if (m.ul_i[u * UL_I_COUNT + UL_KIND] != 0) continue;
    const double* ud = m.ul_d + u * UL_D_COUNT;
    double* rec = lines_n + u * BAYSAR_LINE_STRIDE;
    const double la = ud[UL_LOMAN_A], lb = ud[UL_LOMAN_B], lc = ud[UL_LOMAN_C];
    const double wne = rec[2 + 2 * which], wte = rec[3 + 2 * which];
    // 10 c (1e6 ne)^a / Te^b / 2 with the two general powers as one exponential of logarithms, and g^2.5 = g^2 sqrt(g):
    // ~1e-14 relative against three pow() calls, a third of their instructions (the widths end up in float32)
    const double g = (10.0 * lc * exp(la * log(1e6 * wne) - lb * log(wte))) / 2.0;
    rec[11 + which] = g * g * sqrt(g);
  }
  team.sync();

  priors_pass(true);

  st = team.or_bits(st);
  if (lane == 0) {
    a.status[n] = st;
    double* h = a.hdr + n * BAYSAR_HDR_STRIDE;
    h[0] = nemax; h[1] = 0.0; h[2] = 0.0; h[3] = 0.0;
  }
}
