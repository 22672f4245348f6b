"""TEST INFRASTRUCTURE — synthetic ADF11 / ADF15 text files and what the REFERENCE's readers make of them.

Writes `tests/golden/adf/*.dat` (tables of `baysar_b200.atomic_data.SyntheticADAS` in the ADAS text layouts, via
`baysar_b200.adas_files.format_adf1*`) and `tests/golden/adf_parsed.npz`: the arrays returned by the unmodified
`/root/reference/OpenADAS/read_adf11.py: build_adf11_dataarray` and `read_adf15.py: build_adf15_dataarray` (run under
the xarray stand-in of `oracle/ref_shim.py`) on exactly those files.  `tests/test_adas_files.py` pins
`baysar_b200.adas_files.parse_adf1*` on them.  Run in the build container only: `python -m oracle.make_adf_golden`.
"""
import pathlib
import sys

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from baysar_b200 import adas_files  # noqa: E402
from baysar_b200.atomic_data import SyntheticADAS  # noqa: E402
from oracle import ref_shim  # noqa: E402


IMP_NE = np.logspace(11.5, 15.5, 9)
IMP_TE = np.logspace(-1.2, 2.3, 15)


def synthetic_impurity_adf11(elem, kind):
    """Closed-form, physically shaped stand-in rate coefficients of every charge state of ``elem`` (NOT ADAS data)."""
    from baysar_b200.atomic_data import number_of_ions

    z = number_of_ions(elem) - 1
    g_ne, g_te = np.meshgrid(IMP_NE, IMP_TE, indexing="ij")
    salt = {"N": 1.0, "C": 0.8}.get(elem, 0.9)
    blocks = []
    for k in range(z):
        e_ion = salt * 9.0 * (k + 1) ** 1.6  # eV, grows with charge
        if kind == "scd":
            blocks.append(2.0e-8 / (k + 1) * np.exp(-e_ion / g_te) * np.sqrt(g_te) / (1.0 + g_te / (12.0 * e_ion)))
        elif kind == "acd":
            blocks.append(2.5e-13 * (k + 1) ** 2 * g_te ** -0.7 * (1.0 + (g_ne / 1e14) ** 0.25))
        elif kind == "plt":
            blocks.append(1.1e-26 * (k + 1) * np.exp(-0.45 * e_ion / g_te) * (g_ne / 1e14) ** 0.08)
        else:  # prb
            blocks.append(2.0e-27 * (k + 1) ** 2 * g_te ** -0.4 * (g_ne / 1e14) ** 0.05)
    # (real ADF11 files floor their log10 rates too: the %10.5f columns hold at most two digits before the point)
    return adas_files.GridTable(1 + np.arange(z), IMP_NE, IMP_TE, np.array(blocks).clip(1e-40))


def write_impurity_files(out_dir):
    """Synthetic ADF11 files of N and C and the ADF15 files the multi-species workload's lines name (`pec` entries of
    the line catalogue, block numbers as catalogued): what `FileADAS(impurity_adf11=..., impurity_adf15=...)` reads."""
    from baysar_b200 import configs
    from baysar_b200.input_functions import make_input_dict

    provider = SyntheticADAS()
    for elem in ("N", "C"):
        for kind in ("scd", "acd", "plt", "prb"):
            text = adas_files.format_adf11(synthetic_impurity_adf11(elem, kind), element=elem, adf11type=kind)
            (out_dir / f"{kind}96_{elem.lower()}_synthetic.dat").write_text(text)
    d = make_input_dict(**configs.multi_species(1).input_kwargs)
    need = {}
    for species in d["species"]:
        if species[0] in "HDT" and species[1] == "_":  # hydrogen isotopes read the adf15 table as a whole
            continue
        for rec in d[species].values():
            if isinstance(rec, dict) and "pec" in rec:
                need[rec["pec"]] = max(need.get(rec["pec"], 0), rec["exc_block"], rec["rec_block"])
    g_ne, g_te = IMP_NE[::2], IMP_TE[::2]
    for pec, n_blocks in sorted(need.items()):
        data = np.array([provider.adf15_pecs(pec, b + 1, g_te, g_ne).T for b in range(n_blocks)])  # [block][ne][Te]
        table = adas_files.GridTable(1 + np.arange(n_blocks), g_ne, g_te, data)
        name = pec.split("/")[-1].replace("#", "_") + ".synthetic"
        (out_dir / name).write_text(adas_files.format_adf15(table, label=name))
        print("adf15", pec, "->", name, n_blocks, "blocks")


def main():
    provider = SyntheticADAS()
    mods = ref_shim.install(provider)
    del mods
    # the package re-exports functions named like its modules: take the modules from sys.modules
    ref11, ref15 = sys.modules["OpenADAS.read_adf11"], sys.modules["OpenADAS.read_adf15"]

    out_dir = ROOT / "tests" / "golden" / "adf"
    out_dir.mkdir(parents=True, exist_ok=True)
    arrays = {}
    t15 = provider.hydrogen_adf15()
    balmer = [6562.8, 4861.3, 4340.5, 4101.7, 3970.1, 3889.1, 3835.4, 3797.9, 3770.6, 3750.2, 3734.4, 3721.9, 3711.0,
              3703.9, 3697.2, 3691.6, 3686.8, 3682.8]
    wl = balmer + balmer[:12]
    types = ["EXCIT"] * 18 + ["RECOM"] * 12
    text = adas_files.format_adf15(t15, wavelengths=wl, types=types)
    (out_dir / "pec12_h0_synthetic.dat").write_text(text)
    ref = ref15.build_adf15_dataarray(text)
    arrays["adf15_block"], arrays["adf15_ne"], arrays["adf15_te"] = ref.coords["block"], ref.coords["ne"], ref.coords["Te"]
    arrays["adf15_data"] = ref.data
    for kind in ("scd", "acd", "plt", "prb"):
        text = adas_files.format_adf11(provider.hydrogen_adf11(kind), adf11type=kind)
        (out_dir / f"{kind}12_h_synthetic.dat").write_text(text)
        ref = ref11.build_adf11_dataarray(text)
        arrays[f"{kind}_block"], arrays[f"{kind}_ne"], arrays[f"{kind}_te"] = ref.coords["block"], ref.coords["ne"], ref.coords["Te"]
        arrays[f"{kind}_data"] = ref.data
    # a two-block adf11 (the reader's block stride) from a scaled copy
    t = provider.hydrogen_adf11("scd")
    two = adas_files.GridTable([1, 2], t.ne, t.te, np.concatenate([t.data, 0.37 * t.data]))
    text = adas_files.format_adf11(two, element="HELIUM", adf11type="scd")
    (out_dir / "scd_two_blocks_synthetic.dat").write_text(text)
    ref = ref11.build_adf11_dataarray(text)
    arrays["two_block"], arrays["two_ne"], arrays["two_te"], arrays["two_data"] = (ref.coords["block"], ref.coords["ne"],
                                                                                 ref.coords["Te"], ref.data)
    write_impurity_files(out_dir)
    np.savez_compressed(ROOT / "tests" / "golden" / "adf_parsed.npz", **arrays)
    print("wrote", sorted(p.name for p in out_dir.iterdir()), "and adf_parsed.npz")


if __name__ == "__main__":
    main()
