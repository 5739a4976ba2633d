! Stub of rte-rrtmgp's mo_gas_optics: the abstract ty_gas_optics the reference type extends
! (src/gas_optics_ecckd.f90:4,23) with the deferred bindings it implements (:38-47) and the generic
! gas_optics the drivers call (example/rfmip-rad-irf/ecckd_rfmip_lw.F90:120, ecckd_rfmip_sw.F90:118).
! TEST INFRASTRUCTURE ONLY -- written from the published interface.
module mo_gas_optics
  use mo_rte_kind, only: wp
  use mo_source_functions, only: ty_source_func_lw
  use mo_gas_concentrations, only: ty_gas_concs
  use mo_optical_props, only: ty_optical_props, ty_optical_props_arry
  implicit none
  private

  type, abstract, extends(ty_optical_props), public :: ty_gas_optics
  contains
    generic, public :: gas_optics => gas_optics_int, gas_optics_ext
    procedure(gas_optics_int_abstract), deferred :: gas_optics_int
    procedure(gas_optics_ext_abstract), deferred :: gas_optics_ext
    procedure(logical_abstract), deferred, public :: source_is_internal
    procedure(logical_abstract), deferred, public :: source_is_external
    procedure(real_abstract), deferred, public :: get_press_min
    procedure(real_abstract), deferred, public :: get_press_max
    procedure(real_abstract), deferred, public :: get_temp_min
    procedure(real_abstract), deferred, public :: get_temp_max
  end type ty_gas_optics

  abstract interface
    function gas_optics_int_abstract(this, play, plev, tlay, tsfc, gas_desc, &
                                     optical_props, sources, col_dry, tlev) result(error_msg)
      import ty_gas_optics, wp, ty_gas_concs, ty_optical_props_arry, ty_source_func_lw
      class(ty_gas_optics), intent(in) :: this
      real(wp), dimension(:,:), intent(in) :: play
      real(wp), dimension(:,:), intent(in) :: plev
      real(wp), dimension(:,:), intent(in) :: tlay
      real(wp), dimension(:), intent(in) :: tsfc
      type(ty_gas_concs), intent(in) :: gas_desc
      class(ty_optical_props_arry), intent(inout) :: optical_props
      class(ty_source_func_lw), intent(inout) :: sources
      character(len=128) :: error_msg
      real(wp), dimension(:,:), intent(in), target, optional :: col_dry
      real(wp), dimension(:,:), intent(in), target, optional :: tlev
    end function gas_optics_int_abstract

    function gas_optics_ext_abstract(this, play, plev, tlay, gas_desc, optical_props, toa_src, col_dry) &
      result(error_msg)
      import ty_gas_optics, wp, ty_gas_concs, ty_optical_props_arry
      class(ty_gas_optics), intent(in) :: this
      real(wp), dimension(:,:), intent(in) :: play
      real(wp), dimension(:,:), intent(in) :: plev
      real(wp), dimension(:,:), intent(in) :: tlay
      type(ty_gas_concs), intent(in) :: gas_desc
      class(ty_optical_props_arry), intent(inout) :: optical_props
      real(wp), dimension(:,:), intent(out) :: toa_src
      real(wp), dimension(:,:), intent(in), target, optional :: col_dry
      character(len=128) :: error_msg
    end function gas_optics_ext_abstract

    function logical_abstract(this)
      import ty_gas_optics
      class(ty_gas_optics), intent(in) :: this
      logical :: logical_abstract
    end function logical_abstract

    function real_abstract(this)
      import ty_gas_optics, wp
      class(ty_gas_optics), intent(in) :: this
      real(wp) :: real_abstract
    end function real_abstract
  end interface

end module mo_gas_optics
