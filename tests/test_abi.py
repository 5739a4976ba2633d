"""The C-ABI library loads on a CPU-only box and exports every symbol include/ecckd_b200.h declares."""
import ctypes
import os
import re
import subprocess

import pytest

from conftest import ROOT


def header_functions():
    src = open(os.path.join(ROOT, "include", "ecckd_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ecckd_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_path():
    fns = header_functions()
    for must in ("ecckd_model_create", "ecckd_model_load", "ecckd_gas_optics_lw_f64", "ecckd_gas_optics_sw_f64",
                 "ecckd_gas_optics_lw_f32", "ecckd_gas_optics_sw_f32", "ecckd_gas_optics_lw_host_f64",
                 "ecckd_gas_optics_sw_host_f64", "ecckd_interpolation_indices_f64"):
        assert must in fns


def test_library_exports_every_declared_symbol():
    from ecckd_b200 import _abi
    lib = ctypes.CDLL(_abi.LIB_PATH)
    missing = [f for f in header_functions() if not hasattr(lib, f)]
    assert not missing, f"declared in the header but not exported: {missing}"


def test_python_binding_matches_header():
    from ecckd_b200 import _abi
    assert sorted(_abi.EXPORTS) == header_functions()
    _abi.load_library()   # attaches every prototype; AttributeError if one is missing


def test_no_torch_or_cxx_types_in_abi():
    src = open(os.path.join(ROOT, "include", "ecckd_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)   # comments may mention what is NOT there
    assert "torch" not in src and "std::" not in src and "at::" not in src
    # compiles as plain C
    r = subprocess.run(["gcc", "-std=c99", "-fsyntax-only", "-x", "c", os.path.join(ROOT, "include", "ecckd_b200.h")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_version_string():
    from ecckd_b200 import _abi
    v = _abi.load_library().ecckd_version()
    assert b"sm_100a" in v


def test_library_is_built_for_sm_100a_only():
    from ecckd_b200 import _abi
    r = subprocess.run(["cuobjdump", "-lelf", _abi.LIB_PATH], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_\d+a?", r.stdout))
    assert archs == {"sm_100a"}, archs


def test_product_does_not_reference_the_oracle():
    """The oracle is test infrastructure: nothing under rte-ecckd_b200/ may import or link it."""
    bad = []
    for dp, _, fns in os.walk(os.path.join(ROOT, "rte-ecckd_b200")):
        if os.sep + "lib" in dp:
            continue
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".f90", "Makefile")):
                txt = open(os.path.join(dp, fn), errors="ignore").read()
                if re.search(r"\boracle\b", txt):
                    bad.append(os.path.join(dp, fn))
    assert not bad, bad


def test_compute_entry_points_fail_loudly_without_a_gpu():
    """No CPU fallback: on a box without CUDA, model creation returns ECCKD_ERR_CUDA."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    import ecckd_b200
    from conftest import FILES
    with pytest.raises(ecckd_b200.EcckdError) as ei:
        ecckd_b200.GasOpticsEcckd.load(FILES["lw_fsck"])
    assert ei.value.code == ecckd_b200.ERR_CUDA
