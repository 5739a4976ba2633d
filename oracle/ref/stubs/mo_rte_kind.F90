! Stub of rte-rrtmgp's mo_rte_kind (TEST INFRASTRUCTURE ONLY -- see oracle/ref/README.md).
! The reference module uses only `wp` from it (src/gas_optics_ecckd.f90:6); double precision as in
! rte-rrtmgp's default build.
module mo_rte_kind
  use, intrinsic :: iso_c_binding, only: c_double, c_bool
  implicit none
  public
  integer, parameter :: dp = c_double
  integer, parameter :: wp = dp
  integer, parameter :: wl = kind(.true.)
end module mo_rte_kind
