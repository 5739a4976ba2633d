"""Synthetic atmospheres of the benchmark shapes.

Counter-based: every random number is a hash of (seed, stream, global column
index[, layer]), so a column's atmosphere does not depend on which shard or
block it is generated in -- shards of a sharded run reproduce the single-GPU
inputs.  Runs on any torch device (CPU in the tests, CUDA in bench.py).

Arrays come back column-major, indexed [col, lev] like the Fortran.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Tuple

import torch

from .gas_optics import fortran_empty

SEED = 20260101

# scalar well-mixed gases: RFMIP present-day values (utils.f90:47-52 lists the names);
# ch4/n2o scaled above the tables' reference mole fractions so that the
# relative_linear gases do real work instead of being clamped to zero
SCALARS = {"co2": 397.5e-6, "ch4": 1831e-9 * 1.2, "n2o": 327e-9 * 1.1, "o2": 0.209,
           "cfc11": 233e-12, "cfc12": 520e-12}
# the RFMIP driver's gas order (co2, ch4, n2o, o2, cfc11, cfc12, h2o, o3; no2 is unknown to ecckd)
GAS_ORDER = ("co2", "ch4", "n2o", "o2", "cfc11", "cfc12", "h2o", "o3", "no2")


def _mix(x: torch.Tensor) -> torch.Tensor:
    """splitmix64 finaliser on int64 (logical shifts emulated with masks)."""
    x = (x ^ ((x >> 30) & 0x3FFFFFFFF)) * -4658895280553007687      # 0xBF58476D1CE4E5B9
    x = (x ^ ((x >> 27) & 0x1FFFFFFFFF)) * -7723592293110705685     # 0x94D049BB133111EB
    return x ^ ((x >> 31) & 0x1FFFFFFFF)


def _uniform(seed: int, stream: int, idx: torch.Tensor) -> torch.Tensor:
    """U(0,1) in float64 from a 53-bit hash of (seed, stream, idx); never exactly 0."""
    h = _mix(idx * -7046029254386353131 + (seed * 1000003 + stream) * 0x632BE59BD9B4E019 % (1 << 62))
    h = _mix(h + stream)
    return (((h >> 11) & ((1 << 53) - 1)).to(torch.float64) + 0.5) * (2.0 ** -53)


def _normal(seed: int, stream: int, idx: torch.Tensor) -> torch.Tensor:
    u1, u2 = _uniform(seed, stream, idx), _uniform(seed, stream + 7919, idx)
    return torch.sqrt(-2.0 * torch.log(u1)) * torch.cos(2.0 * math.pi * u2)


def _t_std(p: torch.Tensor) -> torch.Tensor:
    """Piecewise standard-atmosphere-like T(p) [K]."""
    trop = 288.15 * (p / 101325.0).clamp(min=1e-9) ** 0.190263
    strat = 216.65 + 12.0 * torch.log(5474.9 / p)
    meso = 264.65 - 10.0 * torch.log(100.0 / p)
    t = torch.where(p > 22632.0, trop, torch.full_like(p, 216.65))
    t = torch.where(p < 5474.9, strat, t)
    t = torch.where(p < 100.0, meso.clamp(min=170.0), t)
    return t


@dataclass
class Atmosphere:
    ncol: int
    nlay: int
    plev: torch.Tensor   # (ncol, nlay+1) Pa, TOA first
    tlay: torch.Tensor   # (ncol, nlay)
    tlev: torch.Tensor   # (ncol, nlay+1)
    tsfc: torch.Tensor   # (ncol,)
    h2o: torch.Tensor    # (ncol, nlay)
    o3: torch.Tensor     # (ncol, nlay)
    scalars: Dict[str, float]

    def gas_list(self) -> List[Tuple[str, object]]:
        """(name, value) in the RFMIP driver's order; no2 is there to be skipped."""
        vals = dict(self.scalars, h2o=self.h2o, o3=self.o3, no2=0.0)
        return [(g, vals[g]) for g in GAS_ORDER]

    def astype(self, dtype) -> "Atmosphere":
        f = lambda t: _copy_f(t, dtype)
        return Atmosphere(self.ncol, self.nlay, f(self.plev), f(self.tlay), f(self.tlev), self.tsfc.to(dtype),
                          f(self.h2o), f(self.o3), dict(self.scalars))


def _copy_f(t: torch.Tensor, dtype) -> torch.Tensor:
    out = fortran_empty(tuple(t.shape), dtype, t.device)
    out.copy_(t)
    return out


def make_atmosphere(ncol: int, nlay: int = 137, col0: int = 0, device="cpu", seed: int = SEED,
                    ptop: float = 1.0, scalars: Dict[str, float] = None, noise: float = 1.0) -> Atmosphere:
    """Columns col0 .. col0+ncol-1 of the synthetic GCM block.
    `noise` scales the cell-by-cell white noise of the specification (N(0, 5 K) on tlay, LogN(0, 0.5) on h2o); 0 gives
    profiles that vary only from column to column -- used by bench.py to show how much of the kernel time is the noise."""
    dev = torch.device(device)
    col = torch.arange(col0, col0 + ncol, dtype=torch.int64, device=dev)
    lev = torch.arange(0, nlay + 1, dtype=torch.int64, device=dev)
    psfc = 5.0e4 + (1.05e5 - 5.0e4) * _uniform(seed, 1, col)
    s = (lev.to(torch.float64) / nlay) ** 2.5
    # column-major result: build as (nlev, ncol) contiguous and view transposed
    plev_t = ptop + (psfc[None, :] - ptop) * s[:, None]
    play_t = 0.5 * (plev_t[1:, :] + plev_t[:-1, :])
    cell = col[None, :] * 4096 + lev[:nlay, None]
    tlay_t = (_t_std(play_t) + (5.0 * noise) * _normal(seed, 2, cell)).clamp(160.0, 320.0)
    tlev_t = torch.empty_like(plev_t)
    tlev_t[1:-1, :] = 0.5 * (tlay_t[1:, :] + tlay_t[:-1, :])
    tlev_t[0, :] = tlay_t[0, :]
    tlev_t[-1, :] = tlay_t[-1, :]
    tsfc = tlev_t[-1, :] + 2.0 * _normal(seed, 3, col)
    q0 = 1.0e-3 + (2.5e-2 - 1.0e-3) * _uniform(seed, 4, col)
    h2o_t = (q0[None, :] * (play_t / psfc[None, :]) ** 3 * torch.exp((0.5 * noise) * _normal(seed, 5, cell))).clamp(1.0e-7, 4.0e-2)
    o3_t = 8.0e-6 * torch.exp(-0.5 * torch.log(play_t / 1000.0) ** 2) + 2.0e-8
    T = lambda x: x.contiguous().t()
    return Atmosphere(ncol, nlay, T(plev_t), T(tlay_t), T(tlev_t), tsfc.contiguous(), T(h2o_t), T(o3_t),
                      dict(SCALARS if scalars is None else scalars))


def make_rfmip_like(nexp: int = 18, ncol_per_exp: int = 100, nlay: int = 60, device="cpu", seed: int = SEED,
                    press_min: float = 0.6940532) -> List[Atmosphere]:
    """RFMIP-shaped stand-in (the RFMIP file itself is not available offline): `nexp`
    experiments of 100 columns x 60 layers, top level forced to just above the table's
    minimum pressure as ecckd_rfmip_lw.F90:90-94 does, per-experiment scalar gases that
    straddle the ch4/n2o reference mole fractions (both clamp branches)."""
    out = []
    for e in range(nexp):
        f = 0.25 + 0.25 * e                      # 0.25x .. 4.5x present-day
        sc = {"co2": 397.5e-6 * f, "ch4": 1831e-9 * (0.4 + 0.1 * e), "n2o": 327e-9 * (0.8 + 0.04 * e),
              "o2": 0.209, "cfc11": 233e-12 * (e % 3), "cfc12": 520e-12 * ((e + 1) % 3)}
        atm = make_atmosphere(ncol_per_exp, nlay, col0=e * ncol_per_exp, device=device, seed=seed + 17, scalars=sc)
        atm.plev[:, 0] = press_min + 2.220446049250313e-16 * press_min * 4   # press_min + epsilon
        out.append(atm)
    return out
