"""ecckd_b200 -- Python host side of the B200-native ECCKD gas-optics path.

Thin by design: the product is the CUDA library behind include/ecckd_b200.h;
this package mirrors the reference's Fortran interface for tests, bench.py and
Python hosts.  Importing it does not need a GPU; computing does.
"""
from ._abi import (ERR_ARG, ERR_CUDA, ERR_IO, ERR_REFERENCE, KERNEL_AUTO, KERNEL_COLWALK, KERNEL_GENERIC, KERNEL_QUAD, KERNEL_WALK, LIB_PATH, OK,
                   load_library)
from .gas_optics import EcckdError, GasConcs, GasOpticsEcckd, PreparedGases, fortran_empty, to_fortran

__all__ = ["GasOpticsEcckd", "GasConcs", "PreparedGases", "EcckdError", "fortran_empty", "to_fortran", "load_library",
           "KERNEL_AUTO", "KERNEL_WALK", "KERNEL_GENERIC", "KERNEL_QUAD", "KERNEL_COLWALK", "OK", "ERR_ARG", "ERR_CUDA", "ERR_REFERENCE", "ERR_IO", "LIB_PATH"]
