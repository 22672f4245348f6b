"""ctypes binding of the C ABI (include/baysar_b200.h) and the in-tree build of the CUDA library.

PyTorch supplies device memory and streams (plumbing); every computation on the posterior path happens inside
``libbaysar_b200.so``.  There is NO CPU fallback: if the library is missing or no CUDA device is present the
functions here raise.
"""
import ctypes
import os
import pathlib
import subprocess

PKG = pathlib.Path(__file__).resolve().parent
CSRC = PKG / "csrc"
INCLUDE = PKG.parent / "include"
LIB_PATH = PKG / "libbaysar_b200.so"              # the product: include/baysar_b200.h
TOOLS_LIB_PATH = PKG / "libbaysar_b200_tools.so"  # measurement / unit-test hooks: include/baysar_b200_tools.h
LINE_STRIDE = 16  # BAYSAR_LINE_STRIDE (include/baysar_b200.h)
IFC_TILE_N = 128  # ifc::TILE_N (csrc/if_contract.cuh): fine-grid outputs per contraction tile
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--compiler-options", "-fPIC", "-shared"]
OPTION_FIELDS = ("tensor_cores", "ifc_min_taps", "ifc_budget_mb", "folded_tile", "los_min_blocks", "los_team_points",
                 "chord_min_blocks", "chord_threads", "los_team_warps", "serial_launches",
                 "contraction_b_image", "host_chunk_thetas")  # baysar_options_t, in declaration order


class Options(ctypes.Structure):
    """baysar_options_t (include/baysar_b200.h): 0 = library default for every field."""
    _fields_ = [(name, ctypes.c_int) for name in OPTION_FIELDS] + [("reserved", ctypes.c_int * 4)]


def _sources():
    return (sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.cuh")) + sorted(CSRC.glob("*.h")) +
            sorted(INCLUDE.glob("*.h")))


def _build_one(target, source, force, verbose, extra=()):
    if target.exists() and not force:
        if target.stat().st_mtime >= max(s.stat().st_mtime for s in _sources()):
            return None
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = list(extra) + os.environ.get("BAYSAR_NVCC_EXTRA", "").split()  # build experiments only (-D switches)
    cmd = [nvcc, *NVCC_FLAGS, *extra, "-Xptxas", "-v", "-o", str(target), str(source)]
    return subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)


def build_library(force=False, verbose=False, tools=True):
    """Compile csrc/baysar_abi.cu (and csrc/baysar_tools.cu) for sm_100a into the package directory; nvcc cross-compiles
    without a GPU.  The two compilations run side by side.  A library is rebuilt when any source under csrc/ or any
    header under include/ is newer than it."""
    jobs = [(LIB_PATH, _build_one(LIB_PATH, CSRC / "baysar_abi.cu", force, verbose))]
    if tools:
        jobs.append((TOOLS_LIB_PATH, _build_one(TOOLS_LIB_PATH, CSRC / "baysar_tools.cu", force, verbose)))
    for target, proc in jobs:
        if proc is None:
            continue
        out, _ = proc.communicate()
        if proc.returncode != 0:
            raise RuntimeError(f"nvcc failed for {target.name}:\n" + out)
        if verbose:
            print(out)
    return LIB_PATH


_lib = None
_tools = None


def library():
    """Load the CUDA library; fail loudly when it has not been built (no silent fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(the posterior path has no CPU fallback)")
    lib = ctypes.CDLL(str(LIB_PATH))
    vp, i64, i32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int
    lib.baysar_model_create.argtypes = [vp, ctypes.c_size_t, i32, ctypes.POINTER(vp)]
    lib.baysar_model_create.restype = i32
    lib.baysar_model_create_ex.argtypes = [vp, ctypes.c_size_t, i32, ctypes.POINTER(Options), ctypes.POINTER(vp)]
    lib.baysar_model_create_ex.restype = i32
    lib.baysar_model_destroy.argtypes = [vp]
    lib.baysar_model_destroy.restype = None
    lib.baysar_model_query.argtypes = [vp, i32, i32]
    lib.baysar_model_query.restype = i64
    lib.baysar_eval_logp.argtypes = [vp, vp, i64, vp, vp, vp, vp]
    lib.baysar_eval_logp.restype = i32
    lib.baysar_eval_spectra.argtypes = [vp, vp, i64, i32, vp, vp, vp, vp, vp]
    lib.baysar_eval_spectra.restype = i32
    lib.baysar_eval_lines.argtypes = [vp, vp, i64, vp, vp, vp, vp]
    lib.baysar_eval_lines.restype = i32
    lib.baysar_eval_logp_host.argtypes = [vp, vp, i64, vp, vp]
    lib.baysar_eval_logp_host.restype = i32
    lib.baysar_model_set_profiling.argtypes = [vp, i32]
    lib.baysar_model_set_profiling.restype = i32
    lib.baysar_model_get_profile.argtypes = [vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(i64)]
    lib.baysar_model_get_profile.restype = i32
    u64 = ctypes.c_uint64
    lib.baysar_stretch_propose.argtypes = [vp, vp, i32, i64, i64, ctypes.c_double, u64, u64, i64, vp, vp, vp]
    lib.baysar_stretch_propose.restype = i32
    lib.baysar_stretch_accept.argtypes = [vp, vp, vp, vp, vp, vp, i32, i64, u64, u64, i64, vp, vp]
    lib.baysar_stretch_accept.restype = i32
    lib.baysar_last_error.restype = ctypes.c_char_p
    lib.baysar_version.restype = ctypes.c_char_p
    _lib = lib
    return lib


def tools_library():
    """Load the measurement / unit-test library (peak-rate probes, contraction hooks)."""
    global _tools
    if _tools is not None:
        return _tools
    if not TOOLS_LIB_PATH.exists():
        raise RuntimeError(f"{TOOLS_LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'`")
    lib = ctypes.CDLL(str(TOOLS_LIB_PATH))
    vp, i64, i32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int
    lib.baysar_microbench.argtypes = [i32, i32, ctypes.POINTER(ctypes.c_double)]
    lib.baysar_microbench.restype = i32
    lib.baysar_debug_if_contract.argtypes = [vp, i64, i32, vp, i32, vp, vp, vp, ctypes.POINTER(i64),
                                             ctypes.POINTER(i32), ctypes.POINTER(ctypes.c_float)]
    lib.baysar_debug_if_contract.restype = i32
    lib.baysar_debug_pix_contract.argtypes = [vp, i64, i32, vp, i32, vp, vp, i32, i32, vp, i64, vp, i64, vp, vp,
                                              ctypes.POINTER(ctypes.c_float)]
    lib.baysar_debug_pix_contract.restype = i32
    lib.baysar_debug_pix_plan_check.argtypes = [vp, i32, vp, i32, vp, vp, i32, i32, vp]
    lib.baysar_debug_pix_plan_check.restype = i32
    lib.baysar_debug_mma_probe.argtypes = [i32, i32, i32, i32, i32, i32, i32, ctypes.POINTER(ctypes.c_double)]
    lib.baysar_debug_mma_probe.restype = i32
    lib.baysar_tools_last_error.restype = ctypes.c_char_p
    _tools = lib
    return lib


EXPORTED_SYMBOLS = ["baysar_model_create", "baysar_model_create_ex", "baysar_model_destroy", "baysar_model_query",
                    "baysar_eval_logp", "baysar_eval_spectra", "baysar_eval_lines", "baysar_eval_logp_host",
                    "baysar_last_error", "baysar_version", "baysar_stretch_propose", "baysar_stretch_accept",
                    "baysar_model_set_profiling", "baysar_model_get_profile"]
TOOLS_EXPORTED_SYMBOLS = ["baysar_tools_last_error", "baysar_microbench", "baysar_debug_if_contract",
                          "baysar_debug_mma_probe", "baysar_debug_pix_contract", "baysar_debug_pix_plan_check"]


def _tools_check(rc):
    if rc != 0:
        raise RuntimeError("baysar_b200 tools: " + tools_library().baysar_tools_last_error().decode())


def microbench(device=0):
    """Measured FP32-FMA / MUFU / FP64-FMA peak rates of this GPU (flop/s, op/s, flop/s)."""
    lib = tools_library()
    out = {}
    for what, key in ((0, "fp32_fma_flops"), (1, "mufu_ops"), (2, "fp64_fma_flops"), (3, "fp32_fma_reg_flops"),
                      (4, "fp32_fma2_flops")):
        val = ctypes.c_double()
        _tools_check(lib.baysar_microbench(int(device), what, ctypes.byref(val)))
        out[key] = val.value
    return out


def debug_if_contract(spectra, taps, row_min=None):
    """Run the tcgen05 instrument-function contraction on a CUDA float32 tensor [n_rows (multiple of 128), F].
    -> (out [n_rows, F] float32, trapz partial sums [n_rows] float64, kernel milliseconds)."""
    import numpy as np
    import torch

    lib = tools_library()
    spectra = spectra.contiguous()
    n, F = spectra.shape
    taps = np.ascontiguousarray(taps, dtype=np.float64)
    n_tiles = -(-F // IFC_TILE_N)
    out = torch.empty((n, n_tiles * IFC_TILE_N), dtype=torch.float32, device=spectra.device)
    part = torch.empty((n, n_tiles), dtype=torch.float64, device=spectra.device)
    pitch, tiles, ms = ctypes.c_int64(), ctypes.c_int(), ctypes.c_float()
    rc = lib.baysar_debug_if_contract(spectra.data_ptr(), n, F, taps.ctypes.data, len(taps),
                                      row_min.data_ptr() if row_min is not None else None, out.data_ptr(),
                                      part.data_ptr(), ctypes.byref(pitch), ctypes.byref(tiles), ctypes.byref(ms))
    _tools_check(rc)
    torch.cuda.synchronize()
    return out[:, :F], part.sum(dim=1), ms.value


def debug_pix_contract(spectra, taps, pix_idx, pix_frac, tile_n=128):
    """Run the folded pixel contraction (instrument convolution + static wavelength map) on a CUDA float32 tensor
    [n_rows (multiple of 128), F].  -> dict(out [n_rows, out_pitch] float32, pix_col [P], col_left0, x_l, col_right0,
    x_r, n_col_tiles, k_blocks, ring, ms)."""
    import numpy as np
    import torch

    lib = tools_library()
    spectra = spectra.contiguous()
    n, F = spectra.shape
    taps = np.ascontiguousarray(taps, dtype=np.float64)
    pix_idx = np.ascontiguousarray(pix_idx, dtype=np.int32)
    pix_frac = np.ascontiguousarray(pix_frac, dtype=np.float64)
    P = len(pix_idx)
    info = (ctypes.c_int64 * 8)()
    pix_col = np.zeros(P, dtype=np.int32)
    ms = ctypes.c_float()
    args = [spectra.data_ptr(), n, F, taps.ctypes.data, len(taps), pix_idx.ctypes.data, pix_frac.ctypes.data, P, tile_n]
    rc = lib.baysar_debug_pix_contract(*args, None, 0, None, 0, ctypes.cast(info, ctypes.c_void_p), pix_col.ctypes.data,
                                       ctypes.byref(ms))
    _tools_check(rc)
    out = torch.empty((n, info[0]), dtype=torch.float32, device=spectra.device)
    tsum = torch.zeros((n, info[5], 3), dtype=torch.float64, device=spectra.device)  # pxc::TSUM_SLOTS
    rc = lib.baysar_debug_pix_contract(*args, out.data_ptr(), out.numel(), tsum.data_ptr(), tsum.numel(),
                                       ctypes.cast(info, ctypes.c_void_p), pix_col.ctypes.data, ctypes.byref(ms))
    _tools_check(rc)
    torch.cuda.synchronize()
    keys = ("out_pitch", "col_left0", "x_l", "col_right0", "x_r", "n_col_tiles", "k_blocks", "ring")
    res = {k: int(info[i]) for i, k in enumerate(keys)}
    res.update(out=out, tsum=tsum.sum(dim=2), tsum_groups=tsum, pix_col=pix_col, ms=ms.value, tile_n=tile_n)
    return res


def debug_pix_plan_check(spectrum, taps, pix_idx, pix_frac, tile_n=128):
    """Host-only check of the folded-contraction plan (no GPU): see baysar_debug_pix_plan_check in the header."""
    import numpy as np

    lib = tools_library()
    spectrum = np.ascontiguousarray(spectrum, dtype=np.float64)
    taps = np.ascontiguousarray(taps, dtype=np.float64)
    pix_idx = np.ascontiguousarray(pix_idx, dtype=np.int32)
    pix_frac = np.ascontiguousarray(pix_frac, dtype=np.float64)
    res = np.zeros(9)
    rc = lib.baysar_debug_pix_plan_check(spectrum.ctypes.data, len(spectrum), taps.ctypes.data, len(taps),
                                         pix_idx.ctypes.data, pix_frac.ctypes.data, len(pix_idx), tile_n, res.ctypes.data)
    _tools_check(rc)
    keys = ("err_pixels", "err_edges", "err_area", "k_blocks", "n_folded", "clip_violation", "wlen", "ring", "err_colsum")
    return dict(zip(keys, res.tolist()))


class DeviceModel:
    """One compiled model resident in one GPU's HBM."""

    def __init__(self, blob, device=0, options=None):
        """``options``: dict of baysar_options_t fields (see OPTION_FIELDS; 0 / missing = library default)."""
        import torch

        if not torch.cuda.is_available():
            raise RuntimeError("baysar_b200 needs a CUDA device: the posterior path has no CPU fallback")
        self.lib = library()
        self.torch = torch
        self.device = int(device)
        self.handle = ctypes.c_void_p()
        buf = (ctypes.c_ubyte * len(blob)).from_buffer_copy(blob)
        opts = Options()
        for key, val in (options or {}).items():
            if key not in OPTION_FIELDS:
                raise KeyError(f"unknown model option {key!r}; known: {OPTION_FIELDS}")
            setattr(opts, key, int(val))
        self.options = dict(options or {})
        rc = self.lib.baysar_model_create_ex(ctypes.cast(buf, ctypes.c_void_p), len(blob), self.device,
                                             ctypes.byref(opts), ctypes.byref(self.handle))
        self._check(rc)
        q = lambda what, arg=0: int(self.lib.baysar_model_query(self.handle, what, arg))  # noqa: E731
        self.n_params, self.n_chords, self.n_ulines, self.n_priors = q(0), q(1), q(2), q(3)
        self.n_los = q(8)
        self.pixels = [q(6, c) for c in range(self.n_chords)]
        self.fine = [q(7, c) for c in range(self.n_chords)]
        self.uses_tensor_cores = bool(q(11))
        # folded pixel contraction (csrc/pix_contract.cuh) per chord: (k-blocks per theta block, column tile width)
        self.folded_contraction = [(q(12, c), q(13, c)) for c in range(q(1))]

    @property
    def launches_last_eval(self):
        return int(self.lib.baysar_model_query(self.handle, 9, 0))

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError("baysar_b200: " + self.lib.baysar_last_error().decode())

    def close(self):
        if getattr(self, "handle", None) and self.handle.value:
            self.lib.baysar_model_destroy(self.handle)
            self.handle = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return ctypes.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def _theta(self, theta):
        t = self.torch
        if theta.dtype != t.float64 or not theta.is_cuda or theta.dim() != 2 or theta.shape[1] != self.n_params:
            raise ValueError(f"theta must be a CUDA float64 tensor of shape (N, {self.n_params})")
        return theta.contiguous()

    def eval_logp(self, theta, want_components=False):
        t = self.torch
        theta = self._theta(theta)
        n = theta.shape[0]
        logp = t.empty(n, dtype=t.float64, device=theta.device)
        status = t.empty(n, dtype=t.int32, device=theta.device)
        comps = t.empty((n, self.n_chords + self.n_priors), dtype=t.float64, device=theta.device) if want_components else None
        self._check(self.lib.baysar_eval_logp(self.handle, theta.data_ptr(), n, logp.data_ptr(), status.data_ptr(),
                                              comps.data_ptr() if comps is not None else None, self._stream()))
        return logp, status, comps

    def eval_spectra(self, theta, chord, want_fine=False):
        t = self.torch
        theta = self._theta(theta)
        n = theta.shape[0]
        spectra = t.empty((n, self.pixels[chord]), dtype=t.float64, device=theta.device)
        status = t.empty(n, dtype=t.int32, device=theta.device)
        pre = t.empty((n, self.fine[chord]), dtype=t.float32, device=theta.device) if want_fine else None
        post = t.empty((n, self.fine[chord]), dtype=t.float32, device=theta.device) if want_fine else None
        self._check(self.lib.baysar_eval_spectra(self.handle, theta.data_ptr(), n, chord, spectra.data_ptr(),
                                                 pre.data_ptr() if want_fine else None,
                                                 post.data_ptr() if want_fine else None, status.data_ptr(),
                                                 self._stream()))
        return spectra, status, pre, post

    def eval_lines(self, theta, want_profiles=False):
        t = self.torch
        theta = self._theta(theta)
        n = theta.shape[0]
        lines = t.empty((n, max(self.n_ulines, 1), LINE_STRIDE), dtype=t.float64, device=theta.device)
        status = t.empty(n, dtype=t.int32, device=theta.device)
        prof = t.empty((n, 4, self.n_los), dtype=t.float64, device=theta.device) if want_profiles else None
        self._check(self.lib.baysar_eval_lines(self.handle, theta.data_ptr(), n, lines.data_ptr(),
                                               prof.data_ptr() if want_profiles else None, status.data_ptr(),
                                               self._stream()))
        return lines, status, prof

    def set_profiling(self, on):
        self._check(self.lib.baysar_model_set_profiling(self.handle, int(bool(on))))

    def get_profile(self):
        """-> (per-stage total ms {los, chord, contract, pixel, finalize}, number of eval_logp calls) since profiling
        was enabled.  `chord` is the fused chord kernel or, on the tensor-core path, its spectral stage."""
        ms = (ctypes.c_double * 5)()
        calls = ctypes.c_int64()
        self._check(self.lib.baysar_model_get_profile(self.handle, ms, ctypes.byref(calls)))
        return dict(zip(("los", "chord", "contract", "pixel", "finalize"), list(ms))), int(calls.value)

    def eval_logp_host(self, theta_np):
        """HOST buffers in, HOST buffers out: the copies happen inside the C call (the `e2e` path of bench.py)."""
        import numpy as np

        if not (theta_np.dtype == np.float64 and theta_np.flags.c_contiguous):
            theta_np = np.ascontiguousarray(theta_np, dtype=np.float64)
        n = theta_np.shape[0]
        logp = np.empty(n, dtype=np.float64)
        status = np.empty(n, dtype=np.uint32)
        # addresses through __array_interface__: `ndarray.ctypes` builds a helper object per access (microseconds that the
        # 4096-theta step, 0.3 ms end to end, would see)
        addr = lambda a: a.__array_interface__["data"][0]  # noqa: E731
        self._check(self.lib.baysar_eval_logp_host(self.handle, addr(theta_np), n, addr(logp), addr(status)))
        return logp, status
