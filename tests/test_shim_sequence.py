"""tests/c/shim_sequence.c replays the Fortran shim's call sequence against the C ABI in plain C (gcc, the public
header, no Python in the loop).  Here it is compiled, run, and what it wrote is compared with the oracle."""
import os
import subprocess

import numpy as np
import pytest

from conftest import FILES, ROOT, rel_err
from ecckd_b200 import _abi

SRC = os.path.join(ROOT, "tests", "c", "shim_sequence.c")


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("shim") / "shim_sequence")
    libdir = os.path.dirname(_abi.LIB_PATH)
    subprocess.run(["gcc", "-std=c11", "-O1", "-Wall", "-Werror", "-D_POSIX_C_SOURCE=200809L", "-I", os.path.join(ROOT, "include"), SRC,
                    "-o", out, "-L", libdir, "-lecckd_b200", f"-Wl,-rpath,{libdir}", "-lm"], check=True)
    return out


def test_header_is_plain_c_and_the_shim_sequence_links(exe):
    assert os.access(exe, os.X_OK)


def test_without_a_device_model_create_fails_loudly(exe):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = subprocess.run([exe, "--no-device", FILES["lw_fsck"]], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "refused loudly" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(37, 60), (4100, 12)], ids=["odd-ncol", "several-blocks-of-nothing-special"])
@pytest.mark.parametrize("key", list(FILES))
def test_shim_sequence_matches_the_oracle(key, shape, exe, tmp_path, ck_models, c_oracles):
    ncol, nlay = shape
    out = str(tmp_path / "out.bin")
    r = subprocess.run([exe, FILES[key], out, str(ncol), str(nlay)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr + r.stdout
    raw = open(out, "rb").read()
    n_, l_, ngpt, sw = np.frombuffer(raw, np.int32, 4)
    assert (n_, l_) == (ncol, nlay)
    off = 16

    def take(*shape_):
        nonlocal off
        n = int(np.prod(shape_))
        a = np.frombuffer(raw, np.float64, n, off).reshape(shape_, order="F")
        off += 8 * n
        return np.asfortranarray(a)
    plev, tlay, tlev, tsfc = take(ncol, nlay + 1), take(ncol, nlay), take(ncol, nlay + 1), take(ncol)
    h2o, o3 = take(ncol, nlay), take(ncol, nlay)
    co2, ch4, n2o, o2, cfc11, cfc12, no2 = take(7)
    gases = [("co2", float(co2)), ("ch4", float(ch4)), ("n2o", float(n2o)), ("o2", float(o2)), ("cfc11", float(cfc11)),
             ("cfc12", float(cfc12)), ("h2o", h2o), ("o3", o3), ("no2", float(no2))]
    m, co = ck_models[key], c_oracles[key]
    if sw:
        got = {k: take(ncol, nlay, ngpt) for k in ("tau", "ssa", "g")}
        got["toa_src"] = take(ncol, ngpt)
        ref, err = co.gas_optics_ext(plev, tlay, gases)
    else:
        got = {k: take(ncol, nlay, ngpt) for k in ("tau", "lay_source", "lev_source_inc", "lev_source_dec")}
        got["sfc_source"] = take(ncol, ngpt)
        ref, err = co.gas_optics_int(plev, tlay, tsfc, gases, tlev)
    assert err == "" and off == len(raw)
    for k, v in ref.items():
        assert rel_err(got[k], v) <= 1e-12, f"{key} {k}"
