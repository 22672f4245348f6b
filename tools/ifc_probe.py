"""GPU probe of the tcgen05 instrument-function contraction: accuracy statistics and timing at bench size."""
import sys
import numpy as np
import torch

sys.path.insert(0, ".")
from baysar_b200 import runtime  # noqa: E402


def main(n=4096, F=8002, K=8001, sigma=25.0):
    rng = np.random.default_rng(1)
    x = np.arange(F)
    s = 1e9 * (1 + rng.random((256, F)))
    for _ in range(6):
        c, wdt, amp = rng.uniform(0, F, 256), rng.uniform(0.3, 6, 256), 10 ** rng.uniform(12, 16, 256)
        s += amp[:, None] / (1 + np.abs((x[None] - c[:, None]) / wdt[:, None]) ** 2.5)
    s = np.tile(s, (n // 256, 1))
    taps = np.exp(-0.5 * ((np.arange(K) - (K - 1) / 2 + 0.3) / sigma) ** 2)
    taps[taps < 1e-300] = 0.0
    taps /= taps.sum()
    sd = torch.as_tensor(s, dtype=torch.float32, device="cuda")
    row_min = sd.min(dim=1).values.contiguous()
    out, part, ms = runtime.debug_if_contract(sd, taps, row_min)
    out, part, ms = runtime.debug_if_contract(sd, taps, row_min)
    o = (K - 1) // 2
    k = torch.as_tensor(taps[::-1].copy(), device="cuda")[None, None, :]
    ref = torch.nn.functional.conv1d(sd[:256].double()[:, None, :], k, padding=K - 1)[:, 0][:, o:o + F]
    ref = torch.maximum(ref, row_min[:256].double()[:, None])
    rel = (out[:256].double() - ref) / ref
    print(f"n={n} F={F} K={K}: kernel {ms:.3f} ms; signed rel err mean {rel.mean().item():.3e} std {rel.std().item():.3e} "
          f"min {rel.min().item():.3e} max {rel.max().item():.3e}")
    ratio = (out[:256].double().sum(1) / ref.sum(1))
    rel2 = out[:256].double() / ratio[:, None] / ref - 1
    print(f"after area renormalisation: mean {rel2.mean().item():.3e} std {rel2.std().item():.3e} max|.| {rel2.abs().max().item():.3e}")


if __name__ == "__main__":
    main(*[int(a) for a in sys.argv[1:3]])
