"""Static transition catalogue (`adas_line_data`).

Mirrors the *structure* of the reference catalogue (`/root/reference/baysar/line_data.py:40-771`):

    adas_line_data[element] = {"atomic_mass": .., "atomic_charge": .., "ionisation_balance_year": ..,
                               "ions": [...], <ion>: {"default_pecs": str, <cwl or (cwl, ...)>: {...}}}

with each transition record holding ``wavelenght`` (sic, the reference's spelling), ``jj_frac``, ``pec``,
``exc_block``, ``rec_block`` and, for hydrogen isotopes, ``n_upper``/``n_lower``
(reference ``add_ion_info``, line_data.py:14-36).  The facts themselves live in
``data/line_catalogue.json`` (exported by ``tools/export_line_catalogue.py``).

Unlike the reference, :func:`fresh_line_data` returns a *new* nested dict on every call, because the
reference's ``make_input_dict`` pops out-of-range transitions from the module-level catalogue in place
(input_functions.py:306-311), which silently shrinks the catalogue for later calls in the same process.
"""
import json
import pathlib

_CATALOGUE = pathlib.Path(__file__).resolve().parent / "data" / "line_catalogue.json"


def fresh_line_data():
    """Build a fresh copy of the catalogue in the reference's nested-dict layout."""
    raw = json.loads(_CATALOGUE.read_text())
    out = {}
    for e in raw["elements"]:
        edict = dict(e["meta"])
        for ion in e["ions"]:
            idict = {}
            if ion["default_pecs"] is not None:
                idict["default_pecs"] = ion["default_pecs"]
            for rec in ion["lines"]:
                key = tuple(rec["cwl"]) if rec["multiplet"] else rec["cwl"]
                entry = {"wavelenght": key}
                for f, v in rec.items():
                    if f not in ("cwl", "multiplet"):
                        entry[f] = v
                idict[key] = entry
            edict[ion["ion"]] = idict
        out[e["element"]] = edict
    return out


adas_line_data = fresh_line_data()
