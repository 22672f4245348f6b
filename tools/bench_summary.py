#!/usr/bin/env python
"""Print the headline numbers of a bench.py JSON line (and of its extra workloads)."""
import json
import sys


def show(l, indent=""):
    e2e = l.get("e2e") or {}
    par = l.get("parity") or {}
    roof = l.get("roofline") or {}
    print(f"{indent}{l['config']['workload'][:40]:40s} n={l['n_gpus']} {l['scaling']:6s} {l['value'] / 1e6:8.3f} M evals/s  "
          f"{l['ms_per_step']:.4f} ms/step  e2e {e2e.get('value', 0) / 1e6:.3f} M  launches/step {l['gpu_launches'] // l['steps']}  "
          f"ok {l.get('status_ok_fraction', float('nan')):.3f}")
    if "stage_ms" not in l:  # the ensemble line
        print(f"{indent}   {l['implementation']['ensemble']}; acceptance {l['acceptance_fraction']:.3f}, walkers {l['config']['walkers']}")
        return
    print(f"{indent}   stages " + " ".join(f"{k}={v:.4f}" for k, v in l["stage_ms"].items() if v > 0) +
          f" | all-gather: {l['implementation']['logp_all_gather']}")
    if par:
        print(f"{indent}   parity ok={par['ok']} n={par['n']} logp_max_rel={par['logp_max_rel']} spectra_max_rel={par['spectra_max_rel']} "
              f"status_mismatch={par['status_mismatch']}")
    if roof:
        print(f"{indent}   roofline {roof['kernel']} bound={roof['bound']} frac={roof['frac']} share={roof['kernel_share_of_step']:.2f} src={roof['counters_source']}")
        for s, k in roof["kernels"].items():
            if "pipes" in k:
                print(f"{indent}      {s:9s} {k['ms']:.4f} ms " + " ".join(f"{p}:{v['frac']:.3f}" for p, v in k["pipes"].items()))


for raw in open(sys.argv[1]):
    raw = raw.strip()
    if not raw.startswith("{"):
        continue
    line = json.loads(raw)
    show(line)
    if line.get("cpu_baseline"):
        c = line["cpu_baseline"]
        print(f"   cpu baseline {c['value']:.0f} evals/s on {c['cores']} cores ({c['kind']})  clocks {line.get('clocks')}")
    for x in line.get("extra_workloads", []):
        show(x, "  + ")
