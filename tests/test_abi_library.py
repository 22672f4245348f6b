"""The C-ABI library builds for sm_100a, loads, and exports every function include/baysar_b200.h declares."""
import ctypes
import pathlib
import re

import pytest

from baysar_b200 import model, runtime

ROOT = pathlib.Path(__file__).resolve().parent.parent


def declared_functions():
    text = (ROOT / "include" / "baysar_b200.h").read_text()
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
