#!/usr/bin/env python
"""Small invocation of every kernel of the path for compute-sanitizer (tools/sanitize.sh): the Balmer demo chord
(spectral stage -> folded tcgen05 contraction -> pixel stage), the multi-species chord (fused FP32 kernel), the
every-column contraction, the ensemble kernels and the contraction unit-test hook.  Prints one OK line per part."""
import sys

import numpy as np
import torch

sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])

from baysar_b200 import configs, runtime  # noqa: E402
from baysar_b200.atomic_data import SyntheticADAS  # noqa: E402
from baysar_b200.ensemble import StretchEnsemble  # noqa: E402
from baysar_b200.input_functions import make_input_dict  # noqa: E402
from baysar_b200.lineshapes import EsymmtricCauchyPlasma  # noqa: E402
from baysar_b200.posterior import BaysarPosterior  # noqa: E402


def build(wl, options=None):
    np.random.seed(0)
    prof = EsymmtricCauchyPlasma(x=wl.los) if wl.los is not None else None
    return BaysarPosterior(make_input_dict(**wl.input_kwargs), profile_function=prof, atomic_data=SyntheticADAS(),
                           skip_test=True, options=options)


def thetas(post, n, seed=1):
    tb = np.asarray(post.plasma.theta_bounds, dtype=float)
    return tb[:, 0] + np.random.default_rng(seed).random((n, len(tb))) * (tb[:, 1] - tb[:, 0])


def main():
    which = sys.argv[1:] or ["balmer", "toeplitz", "multi", "hot", "ensemble", "hook"]
    n = 160  # two theta blocks of the contraction, the second one partial
    if "balmer" in which:
        post = build(configs.balmer_series())
        lp = post(thetas(post, n))
        print("OK balmer folded contraction", np.isfinite(lp).sum(), "finite of", n)
    if "toeplitz" in which:
        post = build(configs.balmer_series(), {"folded_tile": -1})
        lp = post(thetas(post, n))
        post.spectra_batch(thetas(post, 4), 0, want_fine=True)
        print("OK balmer every-column contraction", np.isfinite(lp).sum(), "finite of", n)
    if "multi" in which:
        post = build(configs.multi_species(1))
        lp = post(thetas(post, 64))
        print("OK multi-species fused", np.isfinite(lp).sum(), "finite of 64")
    if "hot" in which:
        post = build(configs.multi_species(1, flags=dict(cold_neutrals=False, cold_ions=False)))
        lp = post(thetas(post, 32))
        print("OK multi-species hot", np.isfinite(lp).sum(), "finite of 32")
    if "ensemble" in which:
        wl = configs.multi_species(1)
        post = build(wl)
        np.random.seed(11)
        centre = wl.theta_from_tags(post.plasma.slices, post.random_start())
        tb = np.asarray(post.plasma.theta_bounds, dtype=float)
        start = centre + 1e-3 * (tb[:, 1] - tb[:, 0]) * np.random.default_rng(3).normal(size=(64, len(centre)))
        ens = StretchEnsemble(post, start, seed=1)
        ens.run(2)
        print("OK ensemble acceptance", ens.acceptance_fraction)
    if "hook" in which:
        from tests.pxc_cases import make_case

        s, taps, idx, frac = make_case(3002, 455, 20.0, 1024, n=128)
        res = runtime.debug_pix_contract(torch.as_tensor(s, dtype=torch.float32, device="cuda"), taps, idx, frac, 128)
        print("OK pix_contract hook", float(res["out"].abs().max()))
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
