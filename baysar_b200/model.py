"""Model compiler: flatten the host object graph (PlasmaLine, SpectrometerChords, priors) into the descriptor blob
the CUDA library consumes (`csrc/layout.h`).

The reference keeps this information in a web of mutable Python objects that every evaluation walks
(SURVEY.md §1); here it is resolved ONCE into index maps, constant tables and per-chord records so that a theta
batch can be evaluated by three kernel launches with no host involvement.
"""
import collections
import re
import struct

import numpy as np

from .runtime import CSRC
from .spectrometers import ADAS406Lines, BalmerHydrogenLine, XLine

Z_CUT = 8.85  # keep in sync with BAYSAR_Z_CUT (csrc/device_common.cuh)
FWHM_TO_SIGMA = 1 / np.sqrt(8 * np.log(2))


def _parse_layout():
    text = (CSRC / "layout.h").read_text()
    enums = {}
    for name, body in re.findall(r"enum\s+(\w+)\s*\{(.*?)\};", text, flags=re.S):
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        names = [tok.strip() for tok in body.replace("\n", " ").split(",") if tok.strip()]
        enums[name] = {n: i for i, n in enumerate(names)}
    consts = dict(re.findall(r"#define\s+(BAYSAR_BLOB_\w+)\s+(\S+)", text))
    magic = int(consts["BAYSAR_BLOB_MAGIC"].rstrip("L"), 16)
    return enums, magic, int(consts["BAYSAR_BLOB_VERSION"])


ENUMS, BLOB_MAGIC, BLOB_VERSION = _parse_layout()
SEC = ENUMS["BaysarSection"]
_DTYPES = {"SEC_GR_I": np.int32, "SEC_GR_UL": np.int32, "SEC_PR_AUX": np.int32, "SEC_MESH_LO": np.int32, "SEC_I": np.int64, "SEC_SP_I": np.int32, "SEC_ELEM_DENS_IDX": np.int32, "SEC_UL_I": np.int32,
           "SEC_PR_I": np.int32, "SEC_CSO_OFF": np.int32, "SEC_CSO_UL": np.int32, "SEC_CH_I": np.int32,
           "SEC_CL_I": np.int32, "SEC_PIX_IDX": np.int32}


IF_TAP_FLOOR = 1e-12  # instrument-function taps below this fraction of the largest tap are not convolved


def _is_arange(arr):
    """(start, delta) if ``arr`` is bit-identical to numpy arange's fill ``start + j * delta``, else None."""
    arr = np.asarray(arr, dtype=float)
    if len(arr) < 2:
        return None
    start, delta = arr[0], arr[1] - arr[0]
    regen = start + np.arange(len(arr), dtype=float) * delta
    return (float(start), float(delta)) if np.array_equal(regen, arr) else None


def _theta_index(plasma, tag):
    return plasma.slices[tag].start if tag in plasma.slices else -1


def _single_tap_guaranteed(plasma, chord):
    """True if the chord has only Balmer lines whose Doppler kernel is one tap for EVERY theta inside the bounds (cold
    neutrals: sigma depends on theta only through the Doppler shift of the centre).  Mirrors the kernel's block-uniform
    test (csrc/chord_stage.cuh): centre tap inside 6.9 sigma, both neighbours outside, with a 1e-3 margin on sigma."""
    if not chord.lines or not plasma.cold_neutrals:
        return False
    for line in chord.lines:
        if not isinstance(line, BalmerHydrogenLine):
            return False
        taps = np.asarray(line.wavelengths_doppler, dtype=float)
        KD = len(taps)
        if KD < 3:
            return False
        o = (KD - 1) // 2
        zc = np.abs(taps[o - 1:o + 2] - line.cwl)
        tag = line.species + "_velocity"
        vmax = 1e3 * np.abs(plasma.theta_bounds[plasma.slices[tag].start]).max() if tag in plasma.slices else 0.0
        cw = line.cwl * np.array([1 - vmax / 299792458.0, 1 + vmax / 299792458.0])
        sigma = cw * 7.715e-5 * np.sqrt(0.01 / line.atomic_mass) * FWHM_TO_SIGMA
        lo, hi = 6.9 * sigma.min() * (1 - 1e-3), 6.9 * sigma.max() * (1 + 1e-3)
        if not (zc[1] <= lo and zc[0] > hi and zc[2] > hi):
            return False
    return True


def _narrow_taps_guaranteed(plasma, chord):
    """True if every Balmer line of the chord has a Doppler kernel of one or three taps for EVERY theta inside the bounds
    (cold neutrals; the kernel's block-uniform tests of csrc/chord_stage.cuh with a 1e-3 margin on sigma): the chord
    kernel's multi-tap path can then be compiled out (CH_NARROW)."""
    if not plasma.cold_neutrals:
        return False
    for line in chord.lines:
        if not isinstance(line, BalmerHydrogenLine):
            continue
        taps = np.asarray(line.wavelengths_doppler, dtype=float)
        KD = len(taps)
        if KD < 5:
            return False
        o = (KD - 1) // 2
        zc = np.abs(taps[o - 2:o + 3] - line.cwl)
        tag = line.species + "_velocity"
        vmax = 1e3 * np.abs(plasma.theta_bounds[plasma.slices[tag].start]).max() if tag in plasma.slices else 0.0
        cw = line.cwl * np.array([1 - vmax / 299792458.0, 1 + vmax / 299792458.0])
        sigma = cw * 7.715e-5 * np.sqrt(0.01 / line.atomic_mass) * FWHM_TO_SIGMA
        lo, hi = 6.9 * sigma.min() * (1 - 1e-3), 6.9 * sigma.max() * (1 + 1e-3)
        one = zc[2] <= lo and zc[1] > hi and zc[3] > hi
        three = zc[1] <= lo and zc[2] <= lo and zc[3] <= lo and zc[0] > hi and zc[4] > hi
        if not (one or three):
            return False
    return True


def compile_model(plasma, chords, priors, gaussian_likelihood=False):
    """-> (blob bytes, info dict).  ``priors``: the prior descriptors in posterior_components order.
    ``gaussian_likelihood``: the chords return ``likelihood(cauchy=False)`` (reference spectrometers.py:187-188)."""
    e = ENUMS
    I = np.zeros(e["BaysarI"]["I_COUNT"], dtype=np.int64)
    iset = lambda name, v: I.__setitem__(e["BaysarI"][name], int(v))  # noqa: E731
    sec = {}

    los = np.asarray(plasma.los, dtype=float)
    L = len(los)
    w = np.zeros(L)
    w[1:-1] = (los[2:] - los[:-2]) / 2
    w[0], w[-1] = (los[1] - los[0]) / 2, (los[-1] - los[-2]) / 2
    sec["SEC_LOS_X"], sec["SEC_LOS_W"] = los, w

    # ---- profile family
    ne_fn, te_fn = plasma.profile_function.electron_density, plasma.profile_function.electron_temperature
    kinds = (getattr(ne_fn, "kind", None), getattr(te_fn, "kind", None))
    prof = np.zeros(e["BaysarProfD"]["PROF_D_COUNT"])
    prof[e["BaysarProfD"]["PROF_NE_FLOOR"]] = 1e-3
    sec["SEC_MESH_X"], sec["SEC_MESH_LO"] = np.zeros(2), np.zeros(2, dtype=np.int32)
    if kinds == ("esymmetric_cauchy", "esymmetric_cauchy"):
        if te_fn.centre is None:
            raise NotImplementedError("the temperature profile needs a fixed centre (reference default: 0)")
        profile_kind = 0
        if ne_fn.ptype != te_fn.ptype:
            raise NotImplementedError("density and temperature profiles must share one ptype")
        prof[e["BaysarProfD"]["PROF_ESYM_GAUSSIAN"]] = 1.0 if ne_fn.ptype == "gaussian" else 0.0
        prof[e["BaysarProfD"]["PROF_NE_CENTRE_FIXED"]] = 0.0 if ne_fn.centre is None else 1.0
        prof[e["BaysarProfD"]["PROF_NE_CENTRE"]] = 0.0 if ne_fn.centre is None else ne_fn.centre
        prof[e["BaysarProfD"]["PROF_TE_CENTRE"]] = te_fn.centre
    elif kinds == ("bowman_tee", "bowman_tee"):
        profile_kind = 1
        prof[e["BaysarProfD"]["PROF_BACKGROUND"]] = 1.0 if ne_fn.background else 0.0
    elif kinds == ("mesh", "mesh"):
        profile_kind = 2
        if ne_fn.zero_bounds is not None or te_fn.zero_bounds is not None:
            raise NotImplementedError("MeshLine with fixed end values is rejected by the reference's own bounds check")
        prof[e["BaysarProfD"]["PROF_MESH_NE_LOG"]] = 1.0 if ne_fn.log else 0.0
        prof[e["BaysarProfD"]["PROF_MESH_TE_LOG"]] = 1.0 if te_fn.log else 0.0
        lows = []
        for fn in (ne_fn, te_fn):  # scipy interp1d._call_linear: clip(searchsorted(x, x_new), 1, n - 1) - 1
            lows.append(np.clip(np.searchsorted(fn.x_points, los), 1, len(fn.x_points) - 1) - 1)
        sec["SEC_MESH_X"] = np.concatenate([ne_fn.x_points, te_fn.x_points])
        sec["SEC_MESH_LO"] = np.concatenate(lows).astype(np.int32)
        iset("I_MESH_NE_N", len(ne_fn.x_points)), iset("I_MESH_TE_N", len(te_fn.x_points))
    elif kinds == ("closed_form", "closed_form"):
        profile_kind = 3
        pf, pd = e["BaysarProfileFn"], e["BaysarProfD"]
        iset("I_NE_FN", pf[ne_fn.fn]), iset("I_TE_FN", pf[te_fn.fn])
        for fn, aux, aux_set in ((ne_fn, "PROF_NE_AUX", "PROF_NE_AUX_SET"), (te_fn, "PROF_TE_AUX", "PROF_TE_AUX_SET")):
            if fn.fn == "PF_GAUSSIAN" and fn.cwl is not None:
                prof[pd[aux]], prof[pd[aux_set]] = fn.cwl, 1.0
            elif fn.fn == "PF_LEFT_TOP_HAT":
                prof[pd[aux]] = los[1] - los[0]
            elif fn.fn == "PF_LEFT_RIGHT_TOP_HAT":
                # the density of a DoubleSlabPlasma carries the centre as its third parameter and hands it to the temperature
                prof[pd[aux_set]] = 1.0 if fn.alt is not None else 0.0
                if fn.alt is None and not (ne_fn.fn == "PF_LEFT_RIGHT_TOP_HAT" and ne_fn.alt is fn):
                    prof[pd[aux]], prof[pd[aux_set]] = fn.centre, 2.0  # a stand-alone top hat: fixed centre
    else:
        raise NotImplementedError(f"profile family {kinds} is not compiled for the GPU (EsymmtricCauchyPlasma, "
                                  "BowmanTeePlasma, MeshPlasma and the closed-form families of lineshapes.py are)")
    sec["SEC_PROF_D"] = prof

    # ---- species
    species = list(plasma.species)
    sp_i = np.full((max(len(species), 1), e["BaysarSpI"]["SP_I_COUNT"]), -1, dtype=np.int32)
    sp_d = np.zeros((max(len(species), 1), e["BaysarSpD"]["SP_D_COUNT"]))
    from .spectrometers import ATOMIC_MASSES

    for k, s in enumerate(species):
        row = sp_i[k]
        row[e["BaysarSpI"]["SP_DENS_IDX"]] = _theta_index(plasma, s + "_dens")
        row[e["BaysarSpI"]["SP_TAU_IDX"]] = _theta_index(plasma, s + "_tau")
        row[e["BaysarSpI"]["SP_VEL_IDX"]] = _theta_index(plasma, s + "_velocity")
        row[e["BaysarSpI"]["SP_TI_IDX"]] = _theta_index(plasma, s + "_Ti")
        row[e["BaysarSpI"]["SP_IS_H"]] = int(s in plasma.hydrogen_species)
        row[e["BaysarSpI"]["SP_IS_IMPURITY"]] = int(s in plasma.impurity_species)
        sp_d[k, e["BaysarSpD"]["SP_MASS"]] = ATOMIC_MASSES[s[: s.index("_")]]
    sec["SEC_SP_I"], sec["SEC_SP_D"] = sp_i, sp_d
    elem_idx = []
    for elem in plasma.impurities:  # first species of each element carries the density (plasmas.py:881-887)
        first = [s for s in plasma.impurity_species if s.startswith(elem + "_")][0]
        elem_idx.append(_theta_index(plasma, first + "_dens"))
    sec["SEC_ELEM_DENS_IDX"] = np.array(elem_idx or [0], dtype=np.int32)

    # ---- hydrogen tables
    htabs, htab_index = [], {}
    if plasma.contains_hydrogen:
        scd = plasma.neutral_adf11s["scd"]
        sec["SEC_A11_NE"], sec["SEC_A11_TE"], sec["SEC_A11_SCD"] = scd.ne, scd.te, scd.select(1)
        sec["SEC_A11_ACD"] = np.zeros(2)
        if getattr(plasma, "recombining", False) is True:  # plasmas.py:735-737 tests `is True`
            acd = plasma.neutral_adf11s["acd"]
            if not (np.array_equal(acd.ne, scd.ne) and np.array_equal(acd.te, scd.te)):
                raise NotImplementedError("recombining: the hydrogen acd and scd tables must share one (ne, Te) grid")
            sec["SEC_A11_ACD"] = acd.select(1)
            iset("I_RECOMBINING", 1)
        sec["SEC_H_NE"], sec["SEC_H_TE"] = plasma.hydrogen_pec_grid
        for key, tab in plasma.hydrogen_pecs.items():
            htab_index[key] = len(htabs)
            htabs.append(np.asarray(tab, dtype=float))
        iset("I_A11_NNE", len(scd.ne)), iset("I_A11_NTE", len(scd.te))
        iset("I_H_NNE", len(sec["SEC_H_NE"])), iset("I_H_NTE", len(sec["SEC_H_TE"]))
    else:
        for k in ("SEC_A11_NE", "SEC_A11_TE", "SEC_A11_SCD", "SEC_A11_ACD", "SEC_H_NE", "SEC_H_TE"):
            sec[k] = np.zeros(2)
    sec["SEC_H_TAB"] = np.array(htabs) if htabs else np.zeros(2)

    # ---- impurity TEC splines
    sec["SEC_TAU_GRID"] = np.asarray(plasma.adas_plasma_inputs["magical_tau"], dtype=float)
    tec_index, tecs = {}, []
    for key, coeffs in plasma.impurity_tecs.items():
        tec_index[key] = len(tecs)
        tecs.append(coeffs)
    if tecs:
        tx, ty = plasma.tec_knots
        sec["SEC_SPL_TX"], sec["SEC_SPL_TY"] = tx, ty
        iset("I_SPL_NX", len(tx)), iset("I_SPL_NY", len(ty))
        if np.array(tecs).shape[-1] != (len(tx) - 4) * (len(ty) - 4):
            raise ValueError("unexpected spline coefficient count")
    else:
        sec["SEC_SPL_TX"], sec["SEC_SPL_TY"] = np.zeros(8), np.zeros(8)
        iset("I_SPL_NX", 8), iset("I_SPL_NY", 8)
    # (SEC_SPL_C follows the unique-line table: the coefficients are laid out per group of lines that share a species)

    # ---- unique lines and per-chord line objects
    ul_rows, ul_index = [], {}
    cl_i_rows, cl_d_rows, comps = [], [], []
    ch_i = np.zeros((max(len(chords), 1), e["BaysarChI"]["CH_I_COUNT"]), dtype=np.int32)
    ch_d = np.zeros((max(len(chords), 1), e["BaysarChD"]["CH_D_COUNT"]))
    ys, errs, ifs, pix_idx, pix_frac = [], [], [], [], []
    E_CL_I, E_CL_D, E_CH_I, E_CH_D = e["BaysarClI"], e["BaysarClD"], e["BaysarChI"], e["BaysarChD"]
    info = {"uline_names": [], "chord_line_uline": []}

    ul_d_rows = []

    def unique_line(key, kind, sp, tab_a, tab_b, shape=None):
        if key not in ul_index:
            ul_index[key] = len(ul_rows)
            ul_rows.append([kind, sp, tab_a, tab_b])
            ul_d_rows.append(list(shape) if shape is not None else [0.0] * e["BaysarUlD"]["UL_D_COUNT"])
            info["uline_names"].append(key)
        return ul_index[key]

    calint_hi = None
    for c, chord in enumerate(chords):
        x_fm = np.asarray(chord.x_data_fm, dtype=float)
        F, P = len(x_fm), len(chord.x_data)
        grid = _is_arange(x_fm)
        if grid is None:
            raise NotImplementedError("the fine grid must be numpy.arange(low, up, refine) (refine=None is not compiled)")
        row, drow = ch_i[c], ch_d[c]
        row[E_CH_I["CH_P"]], row[E_CH_I["CH_F"]] = P, F
        drow[E_CH_D["CH_X0"]], drow[E_CH_D["CH_DELTA"]] = grid
        drow[E_CH_D["CH_DL"]] = np.diff(x_fm).mean()
        drow[E_CH_D["CH_CALWAVE_CWL"]], drow[E_CH_D["CH_CALWAVE_DISP"]] = chord.calwave_default
        drow[E_CH_D["CH_PIX_MEAN"]] = np.mean(np.arange(P, dtype=float))
        drow[E_CH_D["CH_X_MEAN"]] = x_fm.mean()
        row[E_CH_I["CH_BG_IDX"]] = _theta_index(plasma, f"background_{c}")
        row[E_CH_I["CH_CAL_IDX"]] = _theta_index(plasma, f"cal_{c}")
        row[E_CH_I["CH_CALWAVE_IDX"]] = _theta_index(plasma, f"calwave_{c}")
        row[E_CH_I["CH_CALINT_IDX"]] = _theta_index(plasma, f"calint_func_{c}")

        # instrument function: significant-tap window
        inst = np.asarray(chord.instrument_function_fm, dtype=float)
        KI = len(inst)
        # taps below 1e-12 of the peak are left out: on the demo chords they hold 1e-13 of the kernel's area (a pixel
        # 1e4 times fainter than a line two dozen pixels away moves by 1e-9 of itself), four orders below the float32
        # rounding of the convolution itself
        keep = np.where(np.abs(inst) > IF_TAP_FLOOR * np.abs(inst).max())[0]
        k_lo, k_hi = int(keep.min()), int(keep.max()) + 1
        o = (KI - 1) // 2
        half = max(k_hi - 1 - o, o - k_lo, 0)
        taps_needed = k_hi - k_lo
        row[E_CH_I["CH_KI"]], row[E_CH_I["CH_IF_KLO"]], row[E_CH_I["CH_IF_KHI"]] = KI, k_lo, k_hi
        row[E_CH_I["CH_IF_OFF"]] = sum(len(a) for a in ifs)
        ifs.append(inst)
        if row[E_CH_I["CH_CALINT_IDX"]] >= 0:
            # sampled Gaussian instrument function: size the halo from the upper bound of its FWHM (+25 %)
            fwhm_hi = 1.25 * 10 ** plasma.theta_bounds[row[E_CH_I["CH_CALINT_IDX"]], 1]
            ghalf = int(np.ceil(Z_CUT * fwhm_hi * FWHM_TO_SIGMA / drow[E_CH_D["CH_DL"]])) + 2
            ghalf = min(ghalf, F)
            half = max(half, ghalf)
            taps_needed = max(taps_needed, 2 * ghalf + 1)
            calint_hi = fwhm_hi

        # default wavelength map (spectrometers.py:302-310 with the constant calwave of :86-91)
        pix = np.arange(P, dtype=float)
        pix -= np.mean(pix)
        pix *= chord.calwave_default[1]
        pix += chord.calwave_default[0]
        idx = np.clip(np.searchsorted(x_fm, pix) - 1, 0, F - 2)
        frac = (pix - x_fm[idx]) / (x_fm[idx + 1] - x_fm[idx])
        row[E_CH_I["CH_PIX_OFF"]] = sum(len(a) for a in ys)
        ys.append(np.asarray(chord.y_data, dtype=float))
        errs.append(np.asarray(chord.error, dtype=float))
        pix_idx.append(idx.astype(np.int32))
        pix_frac.append(frac)

        row[E_CH_I["CH_LINE_OFF"]] = len(cl_i_rows)
        row[E_CH_I["CH_N_LINES"]] = len(chord.lines)
        kd_max = 0
        line_ulines = []
        n_balmer = 0
        for line in chord.lines:
            ci = np.zeros(E_CL_I["CL_I_COUNT"], dtype=np.int32)
            cd = np.zeros(E_CL_D["CL_D_COUNT"])
            ci[E_CL_I["CL_COMP_OFF"]] = len(comps)
            if isinstance(line, BalmerHydrogenLine):
                sp = species.index(line.species)
                u = unique_line(line.line, 0, sp, htab_index[line.line + "_exc"], htab_index[line.line + "_rec"],
                                shape=[line.cwl, *line.loman_ij_abc, line.atomic_mass])
                dop = _is_arange(line.wavelengths_doppler)
                if dop is None:
                    raise ValueError("Doppler tap grid is not an arange")
                KD = len(line.wavelengths_doppler)
                kd_max = max(kd_max, KD)
                half = max(half, KD - 1 - (KD - 1) // 2, (KD - 1) // 2)
                taps_needed = max(taps_needed, KD)
                ci[E_CL_I["CL_KIND"]], ci[E_CL_I["CL_ULINE"]], ci[E_CL_I["CL_SPECIES"]] = 0, u, sp
                ci[E_CL_I["CL_KD"]] = KD
                ci[E_CL_I["CL_BRANK"]] = n_balmer
                n_balmer += 1
                if n_balmer > 16:  # BAYSAR_MAX_BALMER_LINES (csrc/chord_stage.cuh): shape records held in shared memory
                    raise NotImplementedError("more than 16 hydrogen lines in one chord")
                cd[E_CL_D["CL_CWL0"]] = line.cwl
                cd[E_CL_D["CL_LOMAN_A"]], cd[E_CL_D["CL_LOMAN_B"]], cd[E_CL_D["CL_LOMAN_C"]] = line.loman_ij_abc
                cd[E_CL_D["CL_MASS"]] = line.atomic_mass
                cd[E_CL_D["CL_DOP_START"]], cd[E_CL_D["CL_DOP_DELTA"]] = dop
            elif isinstance(line, ADAS406Lines):
                sp = species.index(line.species)
                u = unique_line(line.line, 1, sp, tec_index[line.line], -1)
                members = line.cwls if plasma.include_doppler_shifts else line.lines
                ci[E_CL_I["CL_KIND"]], ci[E_CL_I["CL_ULINE"]], ci[E_CL_I["CL_SPECIES"]] = 1, u, sp
                ci[E_CL_I["CL_NCOMP"]] = len(members)
                cd[E_CL_D["CL_MASS"]] = line.atomic_mass
                for cw, fr in zip(members, line.jj_frac):
                    comps.append([cw, fr])
            elif isinstance(line, XLine):
                u = unique_line(line.line_tag, 2, -1, _theta_index(plasma, line.line_tag), -1)
                ci[E_CL_I["CL_KIND"]], ci[E_CL_I["CL_ULINE"]], ci[E_CL_I["CL_SPECIES"]] = 2, u, -1
                ci[E_CL_I["CL_NCOMP"]] = len(line.cwl)
                cd[E_CL_D["CL_FWHM"]] = line.fwhm
                for cw, fr in zip(line.cwl, line.fractions):
                    comps.append([cw, fr])
            else:
                raise TypeError(f"unknown line object {type(line)}")
            cl_i_rows.append(ci)
            cl_d_rows.append(cd)
            line_ulines.append(u)
        info["chord_line_uline"].append(line_ulines)
        row[E_CH_I["CH_B_OPTIONAL"]] = int(_single_tap_guaranteed(plasma, chord))
        row[E_CH_I["CH_NARROW"]] = int(_narrow_taps_guaranteed(plasma, chord))
        row[E_CH_I["CH_HALO"]] = (half + 4 + 3) // 4 * 4
        row[E_CH_I["CH_TAPS_MAX"]] = taps_needed + 8

    sec["SEC_CH_I"], sec["SEC_CH_D"] = ch_i, ch_d
    sec["SEC_CL_I"] = np.array(cl_i_rows, dtype=np.int32) if cl_i_rows else np.zeros((1, E_CL_I["CL_I_COUNT"]), np.int32)
    sec["SEC_CL_D"] = np.array(cl_d_rows) if cl_d_rows else np.zeros((1, E_CL_D["CL_D_COUNT"]))
    sec["SEC_COMP_D"] = np.array(comps, dtype=float) if comps else np.zeros((1, 2))
    sec["SEC_UL_I"] = np.array(ul_rows, dtype=np.int32) if ul_rows else np.zeros((1, 4), np.int32)

    # ---- ADAS line groups: the unique ADAS406 lines of one species share the tau slices, the density and, per
    # line-of-sight point, the spline cell; their coefficients are stored side by side so that one gather serves them all
    groups = collections.OrderedDict()
    for u, (kind, sp, tab_a, _) in enumerate(ul_rows):
        if kind == 1:
            groups.setdefault(sp, []).append((u, tab_a))
    gr_rows, gr_ul, coef_blocks, coef_off = [], [], [], 0
    if groups:
        ncx, ncy = len(sec["SEC_SPL_TX"]) - 4, len(sec["SEC_SPL_TY"]) - 4
        iy1 = np.minimum(np.arange(ncy) + 1, ncy - 1)
    for sp, members in groups.items():
        c = np.array([tecs[t] for _, t in members])  # [n_k][n_tau][cells]
        c = c.reshape(len(members), c.shape[1], ncx, ncy)
        lo, hi = c[:, :-1], c[:, 1:]  # tau slices (i, i + 1)
        block = np.stack([lo, hi, lo[..., iy1], hi[..., iy1]], axis=-1)  # [n_k][n_tau - 1][ncx][ncy][4]
        block = np.ascontiguousarray(np.moveaxis(block, 0, 3))  # [n_tau - 1][ncx][ncy][n_k][4]
        gr_rows.append([sp, len(members), len(gr_ul), coef_off // 4])
        gr_ul += [u for u, _ in members]
        coef_blocks.append(block.ravel())
        coef_off += block.size
    sec["SEC_GR_I"] = np.array(gr_rows or [[0, 0, 0, 0]], dtype=np.int32)
    sec["SEC_GR_UL"] = np.array(gr_ul or [0], dtype=np.int32)
    sec["SEC_SPL_C"] = np.concatenate(coef_blocks) if coef_blocks else np.zeros(4)
    iset("I_N_GROUPS", len(gr_rows))
    sec["SEC_UL_D"] = np.array(ul_d_rows, dtype=float) if ul_d_rows else np.zeros((1, e["BaysarUlD"]["UL_D_COUNT"]))
    cat = lambda arrs, dt: np.concatenate(arrs).astype(dt) if arrs else np.zeros(2, dtype=dt)  # noqa: E731
    sec["SEC_Y"], sec["SEC_ERR"], sec["SEC_IF"] = cat(ys, float), cat(errs, float), cat(ifs, float)
    sec["SEC_PIX_IDX"], sec["SEC_PIX_FRAC"] = cat(pix_idx, np.int32), cat(pix_frac, float)

    # ---- priors
    E_PR = e["BaysarPriorKind"]
    pr_i = np.zeros((max(len(priors), 1), e["BaysarPrI"]["PR_I_COUNT"]), dtype=np.int32)
    pr_d = np.zeros((max(len(priors), 1), e["BaysarPrD"]["PR_D_COUNT"]))
    cso_off, cso_ul = [0] * (len(plasma.impurities) + 1), []
    pr_aux = []  # index lists of the optional priors
    for k, p in enumerate(priors):
        pr_i[k, 0] = E_PR[p.kind]
        if p.kind == "PRIOR_STARK_TO_PEAK":
            pr_i[k, 1] = ul_index[p.line.line]
        elif p.kind == "PRIOR_CHARGE_STATE_ORDER":
            for ei, pairs in enumerate(p.ordered_lines(plasma)):
                for (c, li) in pairs:
                    cso_ul.append(info["chord_line_uline"][c][li])
                cso_off[ei + 1] = len(cso_ul)
        elif p.kind == "PRIOR_WAVELENGTH_CAL":
            pr_i[k, 1] = _theta_index(plasma, "calwave_0")
            pr_d[k, :4] = [p.mean[0], p.mean[1], p.std[0], p.std[1]]
        elif p.kind == "PRIOR_INTENSITY_CAL":
            pr_i[k, 1] = _theta_index(plasma, "cal_0")
        elif p.kind == "PRIOR_GAUSSIAN_IF":
            pr_i[k, 1] = _theta_index(plasma, "calint_func_0")
        elif p.kind == "PRIOR_CURVATURE":
            pr_d[k, 0] = p.scale
        elif p.kind == "PRIOR_TDV":
            pr_d[k, :2] = [p.threshold, p.sigma]
        elif p.kind == "PRIOR_PEAK_TE":
            pr_d[k, :2] = [p.te_min, p.te_err]
        elif p.kind == "PRIOR_TE_MIN":
            pr_d[k, :2] = [p.ratio, p.sigma]
        elif p.kind in ("PRIOR_WALL_CONDITIONS", "PRIOR_SIMPLE_WALL"):
            pr_d[k, :4] = [p.te_min, p.ne_min, p.te_err, p.ne_err]
        elif p.kind == "PRIOR_TAU":
            pairs = p.tau_pairs(plasma)
            pr_i[k, 1], pr_i[k, 2] = len(pr_aux), len(pairs)
            for a_tag, b_tag in pairs:
                pr_aux += [_theta_index(plasma, a_tag), _theta_index(plasma, b_tag)]
            pr_d[k, :2] = [p.mean, p.sigma]
        elif p.kind == "PRIOR_TE_TAU":
            tags = [f"{z}_tau" for z in plasma.impurity_species]
            pr_i[k, 1], pr_i[k, 2] = len(pr_aux), len(tags)
            pr_aux += [_theta_index(plasma, t) for t in tags]
            pr_d[k, 0] = p.sigma
        elif p.kind == "PRIOR_IMPURITY_TAU":
            triples = []
            for g, tags in enumerate(p.tau_groups(plasma)):
                triples += [[g, _theta_index(plasma, a_tag), _theta_index(plasma, b_tag)]
                            for a_tag, b_tag in zip(tags[:-1], tags[1:])]
            pr_i[k, 1], pr_i[k, 2] = len(pr_aux), len(triples)
            for t in triples:
                pr_aux += t
            pr_d[k, 0] = p.sigma
        elif p.kind == "PRIOR_BOWMAN_TEE":
            if profile_kind != 1:
                raise AttributeError("BowmanTeePrior needs the BowmanTee profile family")
            pr_d[k, 0] = p.sigma_err
    sec["SEC_PR_I"], sec["SEC_PR_D"] = pr_i, pr_d
    sec["SEC_PR_AUX"] = np.array(pr_aux or [0], dtype=np.int32)
    sec["SEC_CSO_OFF"] = np.array(cso_off, dtype=np.int32)
    sec["SEC_CSO_UL"] = np.array(cso_ul or [0], dtype=np.int32)

    # ---- scalars
    iset("I_N_PARAMS", plasma.n_params), iset("I_L", L), iset("I_N_CHORDS", len(chords))
    iset("I_N_SPECIES", len(species)), iset("I_N_ELEM", len(plasma.impurities)), iset("I_N_ULINES", len(ul_rows))
    iset("I_N_PRIORS", len(priors)), iset("I_N_CLINES", len(cl_i_rows)), iset("I_PROFILE_KIND", profile_kind)
    iset("I_NE_IDX", plasma.slices["electron_density"].start)
    iset("I_TE_IDX", plasma.slices["electron_temperature"].start)
    iset("I_HAS_HYDROGEN", plasma.contains_hydrogen)
    iset("I_H_SPECIES", species.index(plasma.hydrogen_species[0]) if plasma.contains_hydrogen else 0)
    iset("I_COLD_NEUTRALS", plasma.cold_neutrals), iset("I_COLD_IONS", plasma.cold_ions)
    iset("I_THERMALISED", plasma.thermalised), iset("I_CAL_WAVE", plasma.calibrate_wavelength)
    iset("I_CAL_INT", plasma.calibrate_intensity), iset("I_CAL_IF", plasma.calibrate_instrument_function)
    iset("I_REVERSE_EN", bool(getattr(plasma, "reverse_electroneutrality", False)))
    iset("I_GAUSS_LIKELIHOOD", bool(gaussian_likelihood))
    iset("I_N_HTAB", len(htabs)), iset("I_N_TAU", len(sec["SEC_TAU_GRID"])), iset("I_N_TEC", len(tecs))
    iset("I_P_MAX", max([len(c.x_data) for c in chords] or [0]))
    iset("I_F_MAX", max([len(c.x_data_fm) for c in chords] or [0]))
    sec["SEC_I"] = I
    sec["SEC_D"] = np.zeros(e["BaysarD"]["D_COUNT"])
    info["calint_fwhm_capacity"] = calint_hi

    # ---- pack
    n_sec = SEC["SEC_COUNT"]
    names = [n for n, i in sorted(SEC.items(), key=lambda kv: kv[1]) if n != "SEC_COUNT"]
    header_bytes = 8 * (3 + 2 * n_sec)
    offset = (header_bytes + 31) // 32 * 32  # sections start on 32-byte boundaries (256-bit loads of SEC_SPL_C)
    table, payload = [], []
    for name in names:
        arr = np.ascontiguousarray(sec[name], dtype=_DTYPES.get(name, np.float64))
        raw = arr.tobytes()
        table += [offset, arr.size]
        pad = (-len(raw)) % 32
        payload.append(raw + b"\0" * pad)
        offset += len(raw) + pad
    head = struct.pack(f"<{3 + 2 * n_sec}q", BLOB_MAGIC, BLOB_VERSION, n_sec, *table)
    head += b"\0" * ((-len(head)) % 32)
    blob = head + b"".join(payload)
    info["n_ulines"], info["blob_bytes"] = len(ul_rows), len(blob)
    return blob, info
