"""GPU probe: cycles per kind::tf32 tcgen05.mma for the contraction kernels' issue pattern (csrc/mma_probe.cuh)."""
import ctypes
import sys

sys.path.insert(0, ".")
from baysar_b200 import runtime  # noqa: E402


def probe(tn, form, n_acc, b_stages, pattern, iters=4000):
    lib = runtime.tools_library()
    out = ctypes.c_double()
    rc = lib.baysar_debug_mma_probe(0, tn, form, n_acc, b_stages, pattern, iters, ctypes.byref(out))
    if rc != 0:
        return float("nan")
    return out.value


if __name__ == "__main__":
    import torch
    torch.cuda.init()
    print("tile_n form n_acc_hi b_stages pattern cycles/MMA  (floor 128*N/256)")
    for tn in (32, 64, 96, 128, 256):
        for form in (0, 1):
            for n_acc, pattern in ((1, 0), (1, 1), (1, 2), (2, 0), (4, 0)):
                if (n_acc + 1) * tn > 512:
                    continue
                for bs in (1, 4):
                    c = probe(tn, form, n_acc, bs, pattern)
                    print(f"{tn:5d} {'ts' if form == 0 else 'ss'} {n_acc:3d} {bs:3d} {pattern:3d}  {c:8.1f}   ({tn / 2:.0f})", flush=True)
