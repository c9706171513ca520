"""CPU-only: the C-ABI library builds, loads, and exports every symbol include/agrs.h declares.
No compute is called (there is no GPU here); context creation must fail loudly, not fall back."""
import ctypes
import os

import pytest

import autograd_rs_b200  # noqa: F401
from autograd_rs_b200 import _capi


def test_every_declared_symbol_is_exported():
    names = _capi.declared_symbols()
    assert len(names) >= 40
    L = ctypes.CDLL(_capi.LIB_PATH)
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_binding_covers_header():
    L = _capi.lib()
    for n in _capi.declared_symbols():
        assert getattr(L, n).restype is not None


def test_engine_exports_every_declared_symbol():
    from autograd_rs_b200 import frontend
    names = frontend.declared_symbols()
    assert len(names) >= 50
    L = frontend.lib()
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    for n in names:
        assert getattr(L, n).argtypes is not None or n == "agrs_eng_last_error", f"{n} is not bound in frontend.py"


def test_engine_fails_loudly_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from autograd_rs_b200 import frontend
    with pytest.raises(frontend.Panic) as e:
        frontend.Device(0)
    assert e.value.code == 5 and "no CPU fallback" in str(e.value)


def test_rust_sys_crate_matches_headers():
    """rust/agrs-sys/src/lib.rs (generated, uncompiled here) declares exactly the functions the two headers declare"""
    import importlib.util
    import re
    root = os.path.dirname(_capi.HERE)
    spec = importlib.util.spec_from_file_location("gen_rust_sys", os.path.join(root, "tools", "gen_rust_sys.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    on_disk = open(os.path.join(root, "rust", "agrs-sys", "src", "lib.rs")).read()
    assert on_disk == gen.render(), "run python tools/gen_rust_sys.py"
    from autograd_rs_b200 import frontend
    declared = set(_capi.declared_symbols()) | set(frontend.declared_symbols())
    assert set(re.findall(r"pub fn (agrs_\w+)\(", on_disk)) == declared


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(_capi.AgrsError) as e:
        _capi.Ctx(0)
    assert e.value.code == 5  # AGRS_ERR_NODEVICE


def test_product_path_never_imports_oracle():
    root = os.path.dirname(_capi.HERE)
    pkg = _capi.HERE
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cpp", ".hpp", ".h", ".cu")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), os.path.join(dirpath, f)
    assert os.path.isdir(os.path.join(root, "oracle"))
