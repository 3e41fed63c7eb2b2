"""The drop-in boundary without a GPU: the C-ABI library builds for sm_100a, loads, exports
exactly what include/kllm_b200.h declares, does not depend on the oracle, and the Python
loader fails loudly when the library is absent."""
import re
import subprocess

import pytest

from kuiperllama_b200 import HEADER_PATH, LIB_PATH, KllmError, _SIGNATURES, load_library


def header_functions():
    text = HEADER_PATH.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"\b(kllm_[a-z0-9_]+)\s*\(", text)
    # drop the struct-member callback and type names
    return sorted(set(n for n in names if n not in ("kllm_decoder",)))


def test_every_declared_symbol_is_exported(kllm_lib):
    declared = header_functions()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(kllm_lib, name), f"{name} declared in kllm_b200.h but not exported"
        assert name in _SIGNATURES, f"{name} has no ctypes prototype"
    assert sorted(_SIGNATURES) == declared


def test_library_is_sm100a_and_oracle_free(kllm_lib):
    sass = subprocess.run(["cuobjdump", "-lelf", str(LIB_PATH)], capture_output=True, text=True).stdout
    assert "sm_100a" in sass, sass
    syms = subprocess.run(["nm", "-D", str(LIB_PATH)], capture_output=True, text=True).stdout
    assert "ko_" not in syms and "kref_" not in syms, "product library must not contain oracle code"
    deps = subprocess.run(["ldd", str(LIB_PATH)], capture_output=True, text=True).stdout
    assert "oracle" not in deps and "kuiper_ref" not in deps


def test_product_sources_never_touch_the_oracle():
    from pathlib import Path
    pkg = Path(LIB_PATH).parent.parent
    for p in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")) + \
            list(pkg.rglob("*.h")) + list(pkg.rglob("*.cpp")) + list(pkg.rglob("CMakeLists.txt")):
        text = p.read_text(errors="replace")
        assert "kuiper_oracle" not in text and "oracle.binding" not in text and \
            "liboracle" not in text, f"{p} references the oracle"


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(KllmError):
        load_library(tmp_path / "libkllm_b200.so")


def test_version_and_error_strings(kllm_lib):
    assert b"sm_100a" in kllm_lib.kllm_version()
    assert kllm_lib.kllm_error_string(-1) == b"invalid argument"


def test_argument_validation_without_device(kllm_lib):
    # pure host-side checks: must return KLLM_E_INVALID before touching CUDA
    assert kllm_lib.kllm_gemv_f32(None, None, None, 4, 4, None) == -1
    assert kllm_lib.kllm_rmsnorm_f32(None, None, None, 0, 1e-5, None) == -1
    assert kllm_lib.kllm_decoder_create(None, None, None) == -1
