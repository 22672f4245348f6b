"""Export the reference's static line catalogue to a neutral JSON table.

The catalogue (`/root/reference/baysar/line_data.py:40-771`, `adas_line_data`) is *data*
(wavelengths, multiplet fractions, ADAS block ids), not an algorithm.  It is needed verbatim
for a drop-in `make_input_dict(..., species=[...], ions=[...])`.  This script imports the
reference module read-only (it only depends on numpy) and writes the facts in a flat,
order-preserving JSON layout to `baysar_b200/data/line_catalogue.json`.

Run in the build container only (the reference does not exist on the GPU box):

    python tools/export_line_catalogue.py
"""
import importlib.util
import json
import pathlib

REF = pathlib.Path("/root/reference/baysar/line_data.py")
OUT = pathlib.Path(__file__).resolve().parent.parent / "baysar_b200" / "data" / "line_catalogue.json"


def main():
    spec = importlib.util.spec_from_file_location("_ref_line_data", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    cat = mod.adas_line_data

    out = {"format": 1, "source": "dgahle/baysar baysar/line_data.py (adas_line_data)", "elements": []}
    for elem, edata in cat.items():
        e = {"element": elem, "meta": {}, "ions": []}
        for key, val in edata.items():
            if isinstance(val, dict):
                ion = {"ion": key, "default_pecs": val.get("default_pecs"), "lines": []}
                for lkey, lval in val.items():
                    if lkey == "default_pecs":
                        continue
                    cwl = list(lkey) if isinstance(lkey, tuple) else float(lkey)
                    rec = {"cwl": cwl, "multiplet": isinstance(lkey, tuple)}
                    for f, v in lval.items():
                        if f == "wavelenght":
                            continue  # duplicate of the key
                        rec[f] = v
                    ion["lines"].append(rec)
                e["ions"].append(ion)
            else:
                e["meta"][key] = val
        out["elements"].append(e)
    def _plain(o):  # numpy scalars -> python scalars
        if hasattr(o, "item"):
            return o.item()
        raise TypeError(type(o))

    OUT.write_text(json.dumps(out, indent=1, default=_plain))
    n = sum(len(i["lines"]) for e in out["elements"] for i in e["ions"])
    print(f"wrote {OUT} ({len(out['elements'])} elements, {n} transitions)")


if __name__ == "__main__":
    main()
