"""Python host mirror of the reference's Fortran interface for this path.

`GasOpticsEcckd` mirrors `ty_gas_optics_ecckd`
(/root/reference/src/gas_optics_ecckd.f90:23-48): `load` is load_and_init
(example/rfmip-rad-irf/mo_load_coefficients.F90:19), `gas_optics` is the
inherited generic that resolves to gas_optics_int (:381) when `tsfc` is given
and to gas_optics_ext (:431) otherwise, and the getters are :477-553.
`GasConcs` is the slice of rte-rrtmgp's ty_gas_concs the path uses
(get_num_gases / get_gas_names / get_vmr, call sites :340-351).

Every compute call goes through the C ABI (include/ecckd_b200.h) into the CUDA
library; nothing here computes.  Arrays are indexed like the Fortran --
plev[col, lev], tau[col, lay, g] -- and must be column-major in memory
(numpy order='F', or torch tensors made by `fortran_empty`); numpy arrays take
the host entry points, torch CUDA tensors the device entry points.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Iterable, List, Optional, Sequence, Tuple, Union

import numpy as np

from . import _abi

try:  # torch is plumbing (device memory, streams); the host path works without it
    import torch
except Exception:  # pragma: no cover
    torch = None


class EcckdError(RuntimeError):
    """A non-zero return of the C ABI.  `.code` is the ECCKD_ERR_* value; for
    ECCKD_ERR_REFERENCE the message is the reference's error_msg verbatim and
    `.partial` holds the outputs the reference would already have written."""

    def __init__(self, code: int, msg: str, partial=None):
        super().__init__(msg)
        self.code, self.partial = code, partial


# --------------------------------------------------------------------------- helpers
def _is_torch(a) -> bool:
    return torch is not None and isinstance(a, torch.Tensor)


def _fstrides(shape) -> Tuple[int, ...]:
    s, out = 1, []
    for n in shape:
        out.append(s)
        s *= n
    return tuple(out)


def fortran_empty(shape: Sequence[int], dtype, device=None):
    """Uninitialised column-major array of `shape`: numpy (device None) or torch."""
    shape = tuple(int(s) for s in shape)
    if device is None:
        return np.empty(shape, dtype=dtype, order="F")
    tdt = {np.dtype("float64"): torch.float64, np.dtype("float32"): torch.float32,
           np.dtype("int32"): torch.int32}.get(np.dtype(dtype)) if not isinstance(dtype, torch.dtype) else dtype
    t = torch.empty(shape[::-1], dtype=tdt, device=device)
    return t.permute(*range(len(shape) - 1, -1, -1))


def to_fortran(a, dtype=None, device=None):
    """Column-major copy of an array-like with the given dtype (numpy or torch)."""
    if device is None and not _is_torch(a):
        return np.asfortranarray(np.asarray(a, dtype=dtype))
    t = a if _is_torch(a) else torch.as_tensor(np.ascontiguousarray(a))
    out = fortran_empty(tuple(t.shape), dtype if dtype is not None else t.dtype, device if device is not None else t.device)
    out.copy_(t)
    return out


_EMPTY = C.create_string_buffer(16)   # stands in for the (never dereferenced) address of a zero-size array


def _ptr(a, shape, np_dtype, what: str) -> int:
    """Address of a column-major array after checking shape / dtype / layout."""
    if a is None:
        return 0
    if int(np.prod(tuple(a.shape))) == 0 and tuple(a.shape) == tuple(int(s) for s in shape):
        return C.addressof(_EMPTY)        # present but empty: must not look like an absent optional argument
    shape = tuple(int(s) for s in shape)
    if tuple(a.shape) != shape:
        raise ValueError(f"{what}: shape {tuple(a.shape)} != {shape}")
    n = int(np.prod(shape)) if shape else 1
    if _is_torch(a):
        want = torch.float64 if np_dtype == np.float64 else (torch.float32 if np_dtype == np.float32 else torch.int32)
        if a.dtype != want:
            raise ValueError(f"{what}: dtype {a.dtype} != {want}")
        if n > 1 and tuple(a.stride()) != _fstrides(shape) and not (len(shape) == 1 and a.is_contiguous()):
            # strides of size-1 dims are irrelevant
            ok = all(st == fs or sz == 1 for st, fs, sz in zip(a.stride(), _fstrides(shape), shape))
            if not ok:
                raise ValueError(f"{what}: not column-major (strides {tuple(a.stride())})")
        return a.data_ptr()
    if a.dtype != np_dtype:
        raise ValueError(f"{what}: dtype {a.dtype} != {np.dtype(np_dtype)}")
    if n > 1 and not a.flags.f_contiguous:
        raise ValueError(f"{what}: not column-major (use order='F')")
    return a.ctypes.data


# --------------------------------------------------------------------------- gas concentrations
class GasConcs:
    """The part of rte-rrtmgp's ty_gas_concs this path touches.  Names are
    lower-cased as ty_gas_concs%init does; values may be scalar, (nlay) or (ncol, nlay)."""

    def __init__(self, gas_names: Iterable[str] = ()):
        self._names: List[str] = []
        self._values: Dict[str, object] = {}
        self.init(gas_names)

    def init(self, gas_names: Iterable[str]) -> None:
        self._names = [g.strip().lower() for g in gas_names]
        self._values = {}

    def set_vmr(self, gas: str, value) -> None:
        gas = gas.strip().lower()
        if gas not in self._names:
            self._names.append(gas)
        self._values[gas] = value

    def get_num_gases(self) -> int:
        return len(self._names)

    def get_gas_names(self) -> List[str]:
        return list(self._names)

    def items(self) -> List[Tuple[str, object]]:
        missing = [g for g in self._names if g not in self._values]
        if missing:
            raise EcckdError(_abi.ERR_REFERENCE, f"ty_gas_concs%get_vmr; gas {missing[0]} not found")
        return [(g, self._values[g]) for g in self._names]


class PreparedGases:
    """names[] + ecckd_vmr_view[] for one (ncol, nlay, dtype, device); keeps the
    backing arrays alive.  Build once, reuse across calls."""

    def __init__(self, gas_desc: Union[GasConcs, Sequence[Tuple[str, object]]], ncol: int, nlay: int, np_dtype, device=None):
        items = gas_desc.items() if isinstance(gas_desc, GasConcs) else list(gas_desc)
        self.n = len(items)
        self.ncol, self.nlay, self.np_dtype, self.device = ncol, nlay, np.dtype(np_dtype), device
        self.names = (C.c_char_p * max(self.n, 1))()
        self.views = (_abi.VmrView * max(self.n, 1))()
        self.keep = []
        for j, (name, value) in enumerate(items):
            self.names[j] = name.strip().lower().encode()
            if _is_torch(value):
                v = value
                shape = tuple(v.shape)
            else:
                v = np.asarray(value, dtype=np_dtype)
                shape = v.shape
            if len(shape) == 0 or int(np.prod(shape)) == 1 and len(shape) <= 1:
                arr, cs, ls = self._place(np.asarray(v.cpu() if _is_torch(v) else v, dtype=np_dtype).reshape(1)), 0, 0
            elif len(shape) == 1:
                if shape[0] != nlay:
                    raise EcckdError(_abi.ERR_REFERENCE, f"ty_gas_concs%get_vmr; gas {name} array is wrong size (nlay)")
                arr, cs, ls = self._place(v), 0, 1
            else:
                if shape != (ncol, nlay):
                    raise EcckdError(_abi.ERR_REFERENCE, f"ty_gas_concs%get_vmr; gas {name} array is wrong size (ncol)")
                arr, cs, ls = self._place(v, fortran=True), 1, ncol
            self.keep.append(arr)
            addr = arr.data_ptr() if _is_torch(arr) else arr.ctypes.data
            self.views[j] = _abi.VmrView(addr, cs, ls)

    def _place(self, v, fortran: bool = False):
        if self.device is None:
            a = np.asarray(v.cpu() if _is_torch(v) else v, dtype=self.np_dtype)
            return np.asfortranarray(a) if fortran else np.ascontiguousarray(a)
        tdt = torch.float64 if self.np_dtype == np.float64 else torch.float32
        if fortran:
            if _is_torch(v) and v.dtype == tdt and v.device == torch.device(self.device) and tuple(v.stride()) == _fstrides(v.shape):
                return v
            return to_fortran(v, tdt, self.device)
        t = v if _is_torch(v) else torch.as_tensor(np.ascontiguousarray(v))
        return t.to(device=self.device, dtype=tdt).contiguous()


# --------------------------------------------------------------------------- the model
class GasOpticsEcckd:
    """ty_gas_optics_ecckd behind the C ABI."""

    def __init__(self):
        self._lib = _abi.load_library()
        self._h = C.c_void_p()
        self._keep = None

    # -- construction -------------------------------------------------------
    @classmethod
    def load(cls, filename: str, devices: Optional[Sequence[int]] = None) -> "GasOpticsEcckd":
        """load_and_init (mo_load_coefficients.F90:19-146) through ecckd_model_load."""
        self = cls()
        devs = list(devices) if devices is not None else [0]
        arr = (C.c_int * len(devs))(*devs)
        err = C.create_string_buffer(_abi.ERRLEN)
        rc = self._lib.ecckd_model_load(filename.encode(), len(devs), arr, C.byref(self._h), err)
        if rc != _abi.OK:
            raise EcckdError(rc, err.value.decode())
        return self

    @classmethod
    def from_tables(cls, tables: "_abi.HostTables", keep, devices: Optional[Sequence[int]] = None) -> "GasOpticsEcckd":
        """ecckd_model_create from host tables (what the Fortran shim does after its own loader)."""
        self = cls()
        devs = list(devices) if devices is not None else [0]
        arr = (C.c_int * len(devs))(*devs)
        err = C.create_string_buffer(_abi.ERRLEN)
        rc = self._lib.ecckd_model_create(C.byref(tables), len(devs), arr, C.byref(self._h), err)
        if rc != _abi.OK:
            raise EcckdError(rc, err.value.decode())
        self._keep = keep
        return self

    @classmethod
    def from_components(cls, m, devices: Optional[Sequence[int]] = None) -> "GasOpticsEcckd":
        """ecckd_model_create from the public components of a ty_gas_optics_ecckd-like object -- what the Fortran shim
        does with the object its loader filled (gas_optics_ecckd.f90:23-36).  `m` needs the attributes shortwave, gas
        (names), absorption[i].{coefficient, composite_only, concentration_dependence_code, mole_fraction,
        reference_mole_fraction}, log_pressure, temperature and planck_function / temperature_planck or
        solar_irradiance / rayleigh_molar_scattering_coeff; arrays are C-ordered with the Fortran shape reversed
        (coefficient[v, t, p, g] == coefficient(g, p, t, v)).  Band limits are not passed (nband = 0), like the shim."""
        f8 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        keep = []

        def ptr(a):
            a = f8(a)
            keep.append(a)
            return a.ctypes.data
        ngas = len(m.gas)
        gts = (_abi.GasTable * max(ngas, 1))()
        ngpt = None
        for i, (name, tab) in enumerate(zip(m.gas, m.absorption)):
            k = f8(tab.coefficient)
            if k.ndim == 3:
                k = k[None]
            ngpt = k.shape[-1]
            gts[i].name = name.encode().ljust(_abi.NAME_LEN)[:_abi.NAME_LEN]          # blank padded, like character(len=32)
            gts[i].composite_only = int(bool(tab.composite_only))
            gts[i].concentration_dependence_code = int(tab.concentration_dependence_code)
            mf = getattr(tab, "mole_fraction", None)
            gts[i].n_mole_fraction = 0 if mf is None else int(np.asarray(mf).shape[0])
            gts[i].mole_fraction = None if mf is None else ptr(mf)
            gts[i].reference_mole_fraction = float(tab.reference_mole_fraction)
            gts[i].coefficient = ptr(k)
        t = _abi.HostTables()
        t.shortwave = int(bool(m.shortwave))
        temperature = f8(m.temperature)                                                # [t, p]
        t.ngpt = int(ngpt if ngpt is not None else (m.solar_irradiance if m.shortwave else m.planck_function).shape[-1])
        t.npress, t.ntemp = int(temperature.shape[1]), int(temperature.shape[0])
        t.num_gases = ngas
        t.log_pressure, t.temperature = ptr(m.log_pressure), ptr(temperature)
        if m.shortwave:
            t.nplanck = 0
            t.solar_irradiance, t.rayleigh_molar_scattering_coeff = ptr(m.solar_irradiance), ptr(m.rayleigh_molar_scattering_coeff)
        else:
            t.nplanck = int(np.asarray(m.temperature_planck).shape[0])
            t.temperature_planck, t.planck_function = ptr(m.temperature_planck), ptr(m.planck_function)
        t.gas = gts
        t.nband = 0
        keep.append(gts)
        return cls.from_tables(t, keep, devices)

    def close(self) -> None:
        if self._h:
            self._lib.ecckd_model_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- getters (:477-553 and the inherited get_ngpt/get_nband) --------------
    def get_ngas(self) -> int:
        return self._lib.ecckd_get_ngas(self._h)

    def get_gases(self) -> List[str]:
        out = []
        for i in range(self.get_ngas()):
            b = C.create_string_buffer(_abi.NAME_LEN)
            self._lib.ecckd_get_gas_name(self._h, i, b)
            out.append(b.value.decode())
        return out

    def get_ngpt(self) -> int:
        return self._lib.ecckd_get_ngpt(self._h)

    def get_nband(self) -> int:
        return self._lib.ecckd_get_nband(self._h)

    def get_band2gpt(self) -> np.ndarray:
        b = np.zeros((2, self.get_nband()), dtype=np.int32, order="F")
        self._lib.ecckd_get_band2gpt(self._h, b.ctypes.data)
        return b

    def get_band_lims_wavenumber(self) -> np.ndarray:
        b = np.zeros((2, self.get_nband()), dtype=np.float64, order="F")
        self._lib.ecckd_get_band_lims_wvn(self._h, b.ctypes.data)
        return b

    def get_table_warnings(self) -> str:
        """Grids that are not uniform although the algorithm assumes so ("" when all are)."""
        return self._lib.ecckd_get_table_warnings(self._h).decode()

    def get_ndev(self) -> int:
        return self._lib.ecckd_get_ndev(self._h)

    def source_is_internal(self) -> bool:
        return bool(self._lib.ecckd_source_is_internal(self._h))

    def source_is_external(self) -> bool:
        return bool(self._lib.ecckd_source_is_external(self._h))

    def get_press_min(self) -> float:
        return self._lib.ecckd_get_press_min(self._h)

    def get_press_max(self) -> float:
        return self._lib.ecckd_get_press_max(self._h)

    def get_temp_min(self) -> float:
        return self._lib.ecckd_get_temp_min(self._h)

    def get_temp_max(self) -> float:
        return self._lib.ecckd_get_temp_max(self._h)

    def get_total_solar_irradiance(self) -> float:
        return self._lib.ecckd_get_total_solar_irradiance(self._h)

    # -- compute ------------------------------------------------------------
    def gas_optics(self, play, plev, tlay, tsfc=None, gas_desc=None, *, tlev=None, col_dry=None,
                   out: Optional[dict] = None, two_stream: bool = True, kernel: int = _abi.KERNEL_AUTO,
                   stream: Optional[int] = None, dev_index: int = 0, block_cols: int = 0) -> dict:
        """The generic `gas_optics`: with `tsfc` -> gas_optics_int (:381-426, LW),
        without -> gas_optics_ext (:431-473, SW).  `play` and `col_dry` are accepted
        and ignored, exactly as the reference ignores them.  Returns the output
        arrays (allocated column-major unless passed in `out`)."""
        del play, col_dry
        lw = tsfc is not None
        on_dev = _is_torch(plev) and plev.is_cuda
        npdt = np.dtype(np.float32) if (plev.dtype in (np.float32,) or (_is_torch(plev) and plev.dtype == torch.float32)) else np.dtype(np.float64)
        sfx = "f32" if npdt == np.float32 else "f64"
        ncol, nlay = int(tlay.shape[0]), int(tlay.shape[1])
        ngpt = self.get_ngpt()
        device = plev.device if on_dev else None
        gases = gas_desc if isinstance(gas_desc, PreparedGases) else PreparedGases(gas_desc if gas_desc is not None else [], ncol, nlay, npdt, device)
        if (gases.ncol, gases.nlay, gases.np_dtype) != (ncol, nlay, npdt):
            raise ValueError("PreparedGases was built for a different shape/dtype")
        out = dict(out) if out else {}
        names3 = ("tau", "lay_source", "lev_source_inc", "lev_source_dec") if lw else (("tau", "ssa", "g") if two_stream else ("tau",))
        name2 = "sfc_source" if lw else "toa_src"
        for k in names3:
            if k not in out:
                out[k] = fortran_empty((ncol, nlay, ngpt), npdt, device)
        if name2 not in out:
            out[name2] = fortran_empty((ncol, ngpt), npdt, device)
        p3 = {k: _ptr(out.get(k), (ncol, nlay, ngpt), npdt, k) for k in ("tau", "lay_source", "lev_source_inc", "lev_source_dec", "ssa", "g")}
        p2 = _ptr(out[name2], (ncol, ngpt), npdt, name2)
        a_plev = _ptr(plev, (ncol, nlay + 1), npdt, "plev")
        a_tlay = _ptr(tlay, (ncol, nlay), npdt, "tlay")
        a_tlev = _ptr(tlev, (ncol, nlay + 1), npdt, "tlev")
        a_tsfc = _ptr(tsfc, (ncol,), npdt, "tsfc") if lw else 0
        err = C.create_string_buffer(_abi.ERRLEN)
        if on_dev:
            if stream is None:
                stream = torch.cuda.current_stream(device).cuda_stream
            if lw:
                rc = getattr(self._lib, f"ecckd_gas_optics_lw_{sfx}")(
                    self._h, dev_index, ncol, nlay, a_plev, a_tlay, a_tlev, a_tsfc, gases.n, gases.names, gases.views,
                    p3["tau"], p3["lay_source"], p3["lev_source_inc"], p3["lev_source_dec"], p2, kernel, stream, err)
            else:
                rc = getattr(self._lib, f"ecckd_gas_optics_sw_{sfx}")(
                    self._h, dev_index, ncol, nlay, a_plev, a_tlay, gases.n, gases.names, gases.views,
                    p3["tau"], p3["ssa"], p3["g"], p2, kernel, stream, err)
        else:
            if lw:
                rc = getattr(self._lib, f"ecckd_gas_optics_lw_host_{sfx}")(
                    self._h, ncol, nlay, a_plev, a_tlay, a_tlev, a_tsfc, gases.n, gases.names, gases.views,
                    p3["tau"], p3["lay_source"], p3["lev_source_inc"], p3["lev_source_dec"], p2, block_cols, err)
            else:
                rc = getattr(self._lib, f"ecckd_gas_optics_sw_host_{sfx}")(
                    self._h, ncol, nlay, a_plev, a_tlay, gases.n, gases.names, gases.views,
                    p3["tau"], p3["ssa"], p3["g"], p2, block_cols, err)
        if rc != _abi.OK:
            raise EcckdError(rc, err.value.decode(), partial=out)
        return out

    # -- next row N1: flux solvers on the device (what the reference drivers get from rte-rrtmgp) ---------------------
    def rte_lw(self, out: dict, sfc_emis: float = 0.98, diffusivity: float = 1.66, dev_index: int = 0, stream=None):
        """Broadband LW fluxes (up, down), each (ncol, nlay+1), from the fp64 device tensors gas_optics returned."""
        import torch
        tau = out["tau"]
        ncol, nlay, ngpt = tau.shape
        up, dn = fortran_empty((ncol, nlay + 1), torch.float64, tau.device), fortran_empty((ncol, nlay + 1), torch.float64, tau.device)
        if stream is None:
            stream = torch.cuda.current_stream(tau.device).cuda_stream
        err = C.create_string_buffer(_abi.ERRLEN)
        P = lambda t: C.c_void_p(t.data_ptr())
        rc = self._lib.ecckd_rte_lw_f64(self._h, dev_index, ncol, nlay, ngpt, P(tau), P(out["lay_source"]), P(out["lev_source_inc"]),
                                        P(out["lev_source_dec"]), P(out["sfc_source"]), sfc_emis, diffusivity, P(up), P(dn), stream, err)
        if rc != _abi.OK:
            raise EcckdError(rc, err.value.decode())
        return up, dn

    def rte_sw(self, out: dict, mu0: float = 0.6, sfc_albedo: float = 0.2, dev_index: int = 0, stream=None):
        """Broadband SW fluxes (up, down total, down direct), each (ncol, nlay+1), from gas_optics' fp64 device tensors."""
        import torch
        tau = out["tau"]
        ncol, nlay, ngpt = tau.shape
        mk = lambda: fortran_empty((ncol, nlay + 1), torch.float64, tau.device)
        up, dn, dr = mk(), mk(), mk()
        if stream is None:
            stream = torch.cuda.current_stream(tau.device).cuda_stream
        err = C.create_string_buffer(_abi.ERRLEN)
        P = lambda t: C.c_void_p(t.data_ptr())
        rc = self._lib.ecckd_rte_sw_f64(self._h, dev_index, ncol, nlay, ngpt, P(tau), P(out["ssa"]), P(out["g"]), P(out["toa_src"]),
                                        mu0, sfc_albedo, P(up), P(dn), P(dr), stream, err)
        if rc != _abi.OK:
            raise EcckdError(rc, err.value.decode())
        return up, dn, dr

    def interpolation_indices(self, plev, tlay, lut_gas: Optional[str] = None, lut_vmr=None, dev_index: int = 0) -> dict:
        """Debug: 1-based ip0/it0/iv0 and upper weights from the kernels' own index function (device tensors)."""
        ncol, nlay = int(tlay.shape[0]), int(tlay.shape[1])
        dev = plev.device
        ints = {k: fortran_empty((ncol, nlay), np.int32, dev) for k in ("ip0", "it0", "iv0")}
        dbl = {k: fortran_empty((ncol, nlay), np.float64, dev) for k in ("pw1", "tw1", "vw1")}
        view = None
        keep = None
        if lut_gas is not None:
            keep = PreparedGases([(lut_gas, lut_vmr)], ncol, nlay, np.float64, dev)
            view = C.byref(keep.views[0])
        err = C.create_string_buffer(_abi.ERRLEN)
        rc = self._lib.ecckd_interpolation_indices_f64(
            self._h, dev_index, ncol, nlay, _ptr(plev, (ncol, nlay + 1), np.float64, "plev"),
            _ptr(tlay, (ncol, nlay), np.float64, "tlay"), lut_gas.encode() if lut_gas else None, view,
            ints["ip0"].data_ptr(), dbl["pw1"].data_ptr(), ints["it0"].data_ptr(), dbl["tw1"].data_ptr(),
            ints["iv0"].data_ptr(), dbl["vw1"].data_ptr(), torch.cuda.current_stream(dev).cuda_stream, err)
        if rc != _abi.OK:
            raise EcckdError(rc, err.value.decode())
        return {**ints, **dbl}

    def planck_indices(self, temperature, dev_index: int = 0):
        t = temperature.contiguous().view(-1)
        it0 = torch.empty(t.numel(), dtype=torch.int32, device=t.device)
        w1 = torch.empty(t.numel(), dtype=torch.float64, device=t.device)
        err = C.create_string_buffer(_abi.ERRLEN)
        rc = self._lib.ecckd_planck_indices_f64(self._h, dev_index, t.numel(), t.data_ptr(), it0.data_ptr(), w1.data_ptr(),
                                                torch.cuda.current_stream(t.device).cuda_stream, err)
        if rc != _abi.OK:
            raise EcckdError(rc, err.value.decode())
        return it0, w1
