"""TEST INFRASTRUCTURE — build and call the CPU-emulated kernels (tests/emu) on a compiled model blob."""
import ctypes
import pathlib
import subprocess
import tempfile

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent
_lib = None


def emu_library():
    global _lib
    if _lib is None:
        out = pathlib.Path(tempfile.gettempdir()) / "libbaysar_emu.so"
        srcs = list((ROOT / "baysar_b200" / "csrc").glob("*")) + list((ROOT / "tests" / "emu").glob("*"))
        if not out.exists() or out.stat().st_mtime < max(s.stat().st_mtime for s in srcs):
            cmd = ["g++", "-std=c++20", "-O1", "-fPIC", "-shared", "-pthread", "-I", str(ROOT / "tests" / "emu"), "-o",
                   str(out), str(ROOT / "tests" / "emu" / "emu_driver.cpp")]
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode:
                raise RuntimeError(res.stderr)
        _lib = ctypes.CDLL(str(out))
        _lib.emu_eval.restype = ctypes.c_int
    return _lib


def host_model(workload, atomic_data=None, gaussian_likelihood=False):
    """Host objects + blob without touching a GPU (BaysarPosterior itself needs one)."""
    from baysar_b200.atomic_data import SyntheticADAS
    from baysar_b200.input_functions import make_input_dict
    import baysar_b200.lineshapes as host_ns
    from baysar_b200.model import compile_model
    from baysar_b200.plasmas import PlasmaLine
    from baysar_b200.posterior import build_components, set_sensible_bounds

    d = make_input_dict(**workload.input_kwargs)
    plasma = PlasmaLine(d, profile_function=workload.make_profile(host_ns), atomic_data=atomic_data or workload.make_provider())
    for k, v in workload.plasma_attrs.items():
        setattr(plasma, k, v)
    import baysar_b200.priors as host_priors

    chords, priors = build_components(plasma, d, curvature=workload.curvature,
                                      passed_priors=workload.make_priors(host_priors, late=False))
    set_sensible_bounds(plasma, chords)
    blob, info = compile_model(plasma, chords, priors, gaussian_likelihood=gaussian_likelihood)
    return plasma, chords, priors, blob, info


def emu_eval(blob, theta, chord=0, n_chords=1, n_priors=0, n_ul=1, L=50, P=1, F=1):
    lib = emu_library()
    theta = np.ascontiguousarray(theta, dtype=np.float64)
    n = theta.shape[0]
    out = dict(logp=np.zeros(n), status=np.zeros(n, dtype=np.uint32), comps=np.zeros((n, n_chords + n_priors)),
               lines=np.zeros((n, max(n_ul, 1), 16)), profiles=np.zeros((n, 4, L)), spectra=np.zeros((n, P)),
               fine_pre=np.zeros((n, F), dtype=np.float32), fine_post=np.zeros((n, F), dtype=np.float32))
    buf = (ctypes.c_ubyte * len(blob)).from_buffer_copy(blob)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    rc = lib.emu_eval(buf, p(theta), ctypes.c_long(n), ctypes.c_int(chord), p(out["logp"]), p(out["status"]),
                      p(out["comps"]), p(out["lines"]), p(out["profiles"]), p(out["spectra"]), p(out["fine_pre"]),
                      p(out["fine_post"]))
    assert rc == 0
    return out


def emu_eval_split(blob, theta, chord=0, n_chords=1, n_priors=0, n_ul=1, P=1):
    """Same evaluation through the tensor-core path's spectral + pixel stages (contraction emulated by a loop)."""
    lib = emu_library()
    lib.emu_eval_split.restype = ctypes.c_int
    theta = np.ascontiguousarray(theta, dtype=np.float64)
    n = theta.shape[0]
    out = dict(logp=np.zeros(n), status=np.zeros(n, dtype=np.uint32), comps=np.zeros((n, n_chords + n_priors)),
               lines=np.zeros((n, max(n_ul, 1), 16)), spectra=np.zeros((n, P)))
    buf = (ctypes.c_ubyte * len(blob)).from_buffer_copy(blob)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    rc = lib.emu_eval_split(buf, p(theta), ctypes.c_long(n), ctypes.c_int(chord), p(out["logp"]), p(out["status"]),
                            p(out["comps"]), p(out["lines"]), p(out["spectra"]))
    assert rc == 0
    return out
