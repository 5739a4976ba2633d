"""Minimal netCDF-3 "classic" (CDF-1 / CDF-2) reader in numpy.

TEST INFRASTRUCTURE ONLY (see oracle/README.md): used by the oracle and the
tests to read the ecCKD ck-definition files without libnetcdf, which this image
does not have.  The product has its own, independent C++ reader
(rte-ecckd_b200/csrc/netcdf3.cpp); the two are cross-checked in
tests/test_loader.py.

Format follows the published "NetCDF Classic Format Specification"
(header = magic, numrecs, dim_list, gatt_list, var_list; all big-endian,
names/values padded to 4 bytes).  Only fixed-size (non-record) variables are
supported, which is all the ck-definition files contain.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field

import numpy as np

_NC_DIMENSION, _NC_VARIABLE, _NC_ATTRIBUTE = 0x0A, 0x0B, 0x0C
_TYPES = {1: ">i1", 2: "S1", 3: ">i2", 4: ">i4", 5: ">f4", 6: ">f8"}


@dataclass
class Variable:
    name: str
    dims: tuple
    shape: tuple
    dtype: str
    attrs: dict
    begin: int
    data: np.ndarray = field(repr=False, default=None)


class Dataset:
    """Parsed classic netCDF file: .dims, .attrs, .variables[name].data"""

    def __init__(self, path: str):
        with open(path, "rb") as f:
            self._buf = f.read()
        self._pos = 0
        magic = self._buf[:4]
        if magic[:3] != b"CDF" or magic[3] not in (1, 2):
            raise ValueError(f"{path}: not a netCDF classic file (magic {magic!r})")
        self.version = magic[3]
        self._pos = 4
        self.numrecs = self._i4()
        self.dims = {}
        self._dim_names = []
        self._read_dims()
        self.attrs = self._read_attrs()
        self.variables = {}
        self._read_vars()
        del self._buf

    # -- primitive readers -------------------------------------------------
    def _i4(self) -> int:
        (v,) = struct.unpack_from(">i", self._buf, self._pos)
        self._pos += 4
        return v

    def _i8(self) -> int:
        (v,) = struct.unpack_from(">q", self._buf, self._pos)
        self._pos += 8
        return v

    def _name(self) -> str:
        n = self._i4()
        s = self._buf[self._pos:self._pos + n].decode("ascii")
        self._pos += (n + 3) & ~3
        return s

    # -- header lists ------------------------------------------------------
    def _read_dims(self):
        tag, n = self._i4(), self._i4()
        if tag == 0 and n == 0:
            return
        if tag != _NC_DIMENSION:
            raise ValueError("bad dim_list tag")
        for _ in range(n):
            name = self._name()
            length = self._i4()
            self.dims[name] = length
            self._dim_names.append(name)

    def _read_attrs(self) -> dict:
        tag, n = self._i4(), self._i4()
        out = {}
        if tag == 0 and n == 0:
            return out
        if tag != _NC_ATTRIBUTE:
            raise ValueError("bad att_list tag")
        for _ in range(n):
            name = self._name()
            nc_type = self._i4()
            nelems = self._i4()
            dt = np.dtype(_TYPES[nc_type])
            nbytes = dt.itemsize * nelems
            raw = self._buf[self._pos:self._pos + nbytes]
            self._pos += (nbytes + 3) & ~3
            if nc_type == 2:
                out[name] = raw.decode("ascii", errors="replace")
            else:
                out[name] = np.frombuffer(raw, dtype=dt).astype(dt.newbyteorder("="))
        return out

    def _read_vars(self):
        tag, n = self._i4(), self._i4()
        if tag == 0 and n == 0:
            return
        if tag != _NC_VARIABLE:
            raise ValueError("bad var_list tag")
        for _ in range(n):
            name = self._name()
            ndims = self._i4()
            dimids = [self._i4() for _ in range(ndims)]
            attrs = self._read_attrs()
            nc_type = self._i4()
            _vsize = self._i4()
            begin = self._i4() if self.version == 1 else self._i8()
            dims = tuple(self._dim_names[d] for d in dimids)
            shape = tuple(self.dims[d] for d in dims)
            if any(s == 0 for s in shape):
                raise NotImplementedError(f"record variable {name} not supported")
            dt = np.dtype(_TYPES[nc_type])
            count = int(np.prod(shape, dtype=np.int64)) if shape else 1
            data = np.frombuffer(self._buf, dtype=dt, count=count, offset=begin)
            data = data.reshape(shape).astype(dt.newbyteorder("="))
            self.variables[name] = Variable(name, dims, shape, dt.str, attrs, begin, data)

    def __contains__(self, name):
        return name in self.variables

    def __getitem__(self, name) -> np.ndarray:
        return self.variables[name].data
