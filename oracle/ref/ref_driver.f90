! ref_driver -- runs the UNMODIFIED reference module (/root/reference/src/gas_optics_ecckd.f90, compiled from
! where it lies) on a flat binary dump of a ck-definition model and of an atmosphere, and writes what it returns.
! TEST INFRASTRUCTURE ONLY (oracle/ref/README.md): this is how parity gets pinned to the reference itself the moment a
! Fortran compiler is available.  The type's public components are filled exactly as the reference loader fills
! them (example/rfmip-rad-irf/mo_load_coefficients.F90:46-144); the tables come from oracle/ckd_model.py.
!
! usage: ref_driver tables.bin inputs.bin outputs.bin     (stream files, native endianness; layouts in run_ref.py)
program ref_driver
  use mo_rte_kind, only: wp
  use mo_gas_concentrations, only: ty_gas_concs
  use mo_optical_props, only: ty_optical_props_1scl, ty_optical_props_2str
  use mo_source_functions, only: ty_source_func_lw
  use gas_optics_ecckd, only: ty_gas_optics_ecckd
  implicit none

  type(ty_gas_optics_ecckd) :: k
  type(ty_gas_concs) :: gases
  type(ty_optical_props_1scl) :: props_1scl
  type(ty_optical_props_2str) :: props_2str
  type(ty_source_func_lw) :: sources
  character(len=1024) :: f_tables, f_inputs, f_outputs
  character(len=128) :: err
  character(len=32) :: name
  character(len=32), dimension(:), allocatable :: names
  integer :: u, magic, shortwave, ngpt, npress, ntemp, nplanck, ngas, i, nv, code, comp, kind_
  integer :: ncol, nlay, nlist, one_scalar, with_tlev
  real(wp) :: scalar
  real(wp), dimension(:,:), allocatable :: plev, play, tlay, tlev, toa_src, field
  real(wp), dimension(:), allocatable :: tsfc, profile
  real(wp), dimension(:,:), allocatable :: band_lims

  if (command_argument_count() /= 3) stop "usage: ref_driver tables.bin inputs.bin outputs.bin"
  call get_command_argument(1, f_tables)
  call get_command_argument(2, f_inputs)
  call get_command_argument(3, f_outputs)

  ! ---- the model --------------------------------------------------------------------------------
  open(newunit=u, file=trim(f_tables), access="stream", form="unformatted", status="old", action="read")
  read(u) magic, shortwave, ngpt, npress, ntemp, nplanck, ngas
  if (magic /= 20260101) stop "ref_driver: bad tables file"
  k%shortwave = shortwave /= 0
  k%num_gases = ngas
  k%num_composite_gases = 0
  do i = 1, ngas
    read(u) name, comp, code, nv
    k%gas(i) = name
    k%absorption(i)%composite_only = comp /= 0
    if (comp /= 0) k%num_composite_gases = k%num_composite_gases + 1
    k%absorption(i)%concentration_dependence_code = code
    read(u) k%absorption(i)%reference_mole_fraction
    if (nv > 0) then
      allocate(k%absorption(i)%mole_fraction(nv))
      read(u) k%absorption(i)%mole_fraction
    end if
    allocate(k%absorption(i)%coefficient(ngpt, npress, ntemp, max(nv, 1)))
    read(u) k%absorption(i)%coefficient
  end do
  allocate(k%log_pressure(npress), k%temperature(npress, ntemp), k%gpoint_fraction(1, ngpt))
  read(u) k%log_pressure
  read(u) k%temperature
  k%gpoint_fraction(:,:) = 1._wp       ! only its size(,2) = ngpt is used (:110,265,307,344,468)
  if (k%shortwave) then
    allocate(k%solar_irradiance(ngpt), k%rayleigh_molar_scattering_coeff(ngpt))
    read(u) k%solar_irradiance
    read(u) k%rayleigh_molar_scattering_coeff
    read(u) k%total_solar_irradiance
  else
    allocate(k%temperature_planck(nplanck), k%planck_function(ngpt, nplanck))
    read(u) k%temperature_planck
    read(u) k%planck_function
  end if
  close(u)
  allocate(band_lims(2, 1))
  band_lims(:, 1) = [0._wp, 1._wp]
  err = k%init(band_lims, reshape([1, ngpt], [2, 1]), "ecckd")
  if (len_trim(err) > 0) stop "ref_driver: init failed"

  ! ---- the atmosphere ---------------------------------------------------------------------------
  open(newunit=u, file=trim(f_inputs), access="stream", form="unformatted", status="old", action="read")
  read(u) magic, ncol, nlay, with_tlev, nlist
  if (magic /= 20260102) stop "ref_driver: bad inputs file"
  allocate(plev(ncol, nlay + 1), play(ncol, nlay), tlay(ncol, nlay), tlev(ncol, nlay + 1), tsfc(ncol))
  read(u) plev
  read(u) tlay
  read(u) tlev
  read(u) tsfc
  play(:,:) = 0.5_wp * (plev(:, 1:nlay) + plev(:, 2:nlay + 1))   ! never read by the reference (Q5)
  allocate(names(nlist))
  do i = 1, nlist                       ! names first: ty_gas_concs%init takes the whole list, caller's order
    read(u) names(i)
  end do
  err = gases%init(names)
  if (len_trim(err) > 0) stop "ref_driver: gas_concs init failed"
  do i = 1, nlist
    read(u) kind_
    if (kind_ == 0) then
      read(u) scalar
      err = gases%set_vmr(trim(names(i)), scalar)
    else if (kind_ == 1) then
      allocate(profile(nlay))
      read(u) profile
      err = gases%set_vmr(trim(names(i)), profile)
      deallocate(profile)
    else
      allocate(field(ncol, nlay))
      read(u) field
      err = gases%set_vmr(trim(names(i)), field)
      deallocate(field)
    end if
    if (len_trim(err) > 0) stop "ref_driver: set_vmr failed"
  end do
  close(u)

  ! ---- the call (the drivers' own: ecckd_rfmip_lw.F90:120, ecckd_rfmip_sw.F90:118) ---------------
  open(newunit=u, file=trim(f_outputs), access="stream", form="unformatted", status="replace", action="write")
  if (k%shortwave) then
    err = props_2str%alloc_2str(ncol, nlay, ngpt)
    allocate(toa_src(ncol, ngpt))
    toa_src(:,:) = 0._wp
    err = k%gas_optics(play, plev, tlay, gases, props_2str, toa_src)
    write(u) 20260103, len_trim(err)
    write(u) err
    write(u) props_2str%tau
    write(u) props_2str%ssa
    write(u) props_2str%g
    write(u) toa_src
  else
    err = props_1scl%alloc_1scl(ncol, nlay, ngpt)
    err = sources%alloc(ncol, nlay, ngpt)
    if (with_tlev /= 0) then
      err = k%gas_optics(play, plev, tlay, tsfc, gases, props_1scl, sources, tlev=tlev)
    else
      err = k%gas_optics(play, plev, tlay, tsfc, gases, props_1scl, sources)
    end if
    write(u) 20260103, len_trim(err)
    write(u) err
    write(u) props_1scl%tau
    write(u) sources%lay_source
    write(u) sources%lev_source_inc
    write(u) sources%lev_source_dec
    write(u) sources%sfc_source
  end if
  close(u)
  one_scalar = 0
end program ref_driver
