"""The C-ABI libraries build for sm_100a, load, and export every function their headers declare (include/baysar_b200.h:
the product boundary; include/baysar_b200_tools.h: measurement and unit-test hooks)."""
import ctypes
import pathlib
import re

import pytest

from baysar_b200 import model, runtime

ROOT = pathlib.Path(__file__).resolve().parent.parent


def declared_functions(header="baysar_b200.h"):
    text = (ROOT / "include" / header).read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(baysar_\w+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    lib_path = runtime.build_library()
    lib = ctypes.CDLL(str(lib_path))  # loading does not need a GPU
    names = declared_functions()
    assert sorted(runtime.EXPORTED_SYMBOLS) == names
    for name in names:
        assert hasattr(lib, name), name
    lib.baysar_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.baysar_version()


def test_tools_library_exports_every_declared_symbol_and_nothing_leaks_into_the_product():
    runtime.build_library()
    tools = ctypes.CDLL(str(runtime.TOOLS_LIB_PATH))
    names = declared_functions("baysar_b200_tools.h")
    assert sorted(runtime.TOOLS_EXPORTED_SYMBOLS) == names
    for name in names:
        assert hasattr(tools, name), name
    product = ctypes.CDLL(str(runtime.LIB_PATH))
    for name in names:  # debug hooks, probes and micro-benchmarks are not part of the drop-in boundary
        assert not hasattr(product, name), name


def test_model_create_rejects_malformed_blobs():
    """baysar_model_create validates the descriptor before anything indexes it (no GPU needed: validation comes first)."""
    lib = runtime.library()
    out = ctypes.c_void_p()
    garbage = (ctypes.c_ubyte * 4096)()
    assert lib.baysar_model_create(ctypes.cast(garbage, ctypes.c_void_p), 4096, 0, ctypes.byref(out)) != 0
    assert b"magic" in lib.baysar_last_error()
    import struct

    n_sec = model.SEC["SEC_COUNT"]
    # right magic / version / section count, but every section claims more elements than the blob holds
    table = []
    for _ in range(n_sec):
        table += [16 * ((8 * (3 + 2 * n_sec) + 15) // 16), 1 << 20]
    head = struct.pack(f"<{3 + 2 * n_sec}q", model.BLOB_MAGIC, model.BLOB_VERSION, n_sec, *table)
    blob = head + b"\0" * 1024
    buf = (ctypes.c_ubyte * len(blob)).from_buffer_copy(blob)
    assert lib.baysar_model_create(ctypes.cast(buf, ctypes.c_void_p), len(blob), 0, ctypes.byref(out)) != 0
    assert b"out of range" in lib.baysar_last_error()


def test_layout_header_parses():
    assert model.SEC["SEC_COUNT"] == len(model.SEC) - 1
    assert model.ENUMS["BaysarI"]["I_COUNT"] == len(model.ENUMS["BaysarI"]) - 1
    assert model.BLOB_VERSION >= 1


def test_product_fails_loudly_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        runtime.DeviceModel(b"\0" * 64, 0)


def test_product_never_imports_the_oracle():
    for py in (ROOT / "baysar_b200").glob("*.py"):
        src = py.read_text()
        assert "import oracle" not in src and "from oracle" not in src, py


def test_options_struct_mirrors_the_header():
    """`runtime.Options` (ctypes) has the fields of `baysar_options_t` in declaration order and the same size."""
    import ctypes
    import re

    from baysar_b200 import runtime

    text = (runtime.INCLUDE / "baysar_b200.h").read_text()
    body = text[text.index("typedef struct baysar_options {"):text.index("} baysar_options_t;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    decls = re.findall(r"\bint\s+(\w+)(?:\[(\d+)\])?\s*;", body)
    names = [n for n, dim in decls if not dim]
    reserved = [int(dim) for n, dim in decls if dim]
    assert tuple(names) == runtime.OPTION_FIELDS
    assert ctypes.sizeof(runtime.Options) == 4 * (len(names) + sum(reserved)) == 64
