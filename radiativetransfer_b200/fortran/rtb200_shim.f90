!--- rtb200_shim: ISO_C_BINDING bridge between the unchanged FTTE driver (equiSources.f90) and librtb200.so.
!--- The octree (zoneType, definitionsModule.f90:163-182) has pointer components and is not interoperable, so the
!--- shim flattens the leaves in writeCell order (equiSources.f90:4044-4079), calls the C-ABI of include/rtb200.h and
!--- scatters the results back into the same cells.  A non-zero status reproduces the reference's `write; stop`.
!--- NOTE: no Fortran compiler exists in the build image of this repository; this file is syntax-reviewed only.
module rtb200_shim

  use iso_c_binding
  use definitions

  implicit none
  type(c_ptr), save :: rtbContext = c_null_ptr
  integer(c_int64_t), save :: rtbLeaves = 0
  integer(c_int8_t), dimension(:), allocatable, target, save :: flatLevel
  real(c_double), dimension(:), allocatable, target, save :: flatHI, flatHeI, flatHeII, flatRho, flatAbun2, &
       flatJ1, flatJ2, flatJ3
  integer(c_int64_t), save :: icursor
  ! point sources: six rate fields, leaf number of the first leaf of every base cell
  real(c_double), dimension(:), allocatable, target, save :: flatK24, flatK25, flatK26, flatC24, flatC25, flatC26
  integer(c_int64_t), dimension(:,:,:), allocatable, save :: baseFirstLeaf

  interface
     integer(c_int) function rtb200_create(device, ctx) bind(C, name='rtb200_create')
       import :: c_int, c_ptr
       integer(c_int), value :: device
       type(c_ptr) :: ctx
     end function rtb200_create
     integer(c_int) function rtb200_destroy(ctx) bind(C, name='rtb200_destroy')
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx
     end function rtb200_destroy
     integer(c_int) function rtb200_grid_set(ctx, nx, nleaf, level, HI, HeI, HeII, rho, abun2, boxSize) &
          bind(C, name='rtb200_grid_set')
       import :: c_int, c_int64_t, c_ptr, c_double
       type(c_ptr), value :: ctx
       integer(c_int), value :: nx
       integer(c_int64_t), value :: nleaf
       type(c_ptr), value :: level, HI, HeI, HeII, rho, abun2
       real(c_double), value :: boxSize
     end function rtb200_grid_set
     integer(c_int) function rtb200_grid_update_species(ctx, HI, HeI, HeII) bind(C, name='rtb200_grid_update_species')
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx, HI, HeI, HeII
     end function rtb200_grid_update_species
     integer(c_int) function rtb200_diffuse(ctx, nAngularLevel, uvb, beta, rays, nrays, J1, J2, J3, nseg) &
          bind(C, name='rtb200_diffuse')
       import :: c_int, c_int32_t, c_ptr
       type(c_ptr), value :: ctx
       integer(c_int), value :: nAngularLevel
       type(c_ptr), value :: uvb, beta, rays
       integer(c_int32_t), value :: nrays
       type(c_ptr), value :: J1, J2, J3, nseg
     end function rtb200_diffuse
     integer(c_int) function rtb200_chemistry_tables(ctx, nratec, logtem0, logtem9, dlogtem, k1a, k2a, k3a, k4a, k5a, k6a) &
          bind(C, name='rtb200_chemistry_tables')
       import :: c_int, c_ptr, c_double
       type(c_ptr), value :: ctx
       integer(c_int), value :: nratec
       real(c_double), value :: logtem0, logtem9, dlogtem
       type(c_ptr), value :: k1a, k2a, k3a, k4a, k5a, k6a
     end function rtb200_chemistry_tables
     integer(c_int) function rtb200_chemistry_temperature(ctx, tgas) bind(C, name='rtb200_chemistry_temperature')
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx, tgas
     end function rtb200_chemistry_temperature
     integer(c_int) function rtb200_chemistry_device(ctx, rates_device, J_device, ksi, uniform, maxChange, stream) &
          bind(C, name='rtb200_chemistry_device')
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx, rates_device, J_device, ksi, uniform, maxChange, stream
     end function rtb200_chemistry_device
     integer(c_int) function rtb200_compute_mass(ctx, neutralMass, totalMass, stream) bind(C, name='rtb200_compute_mass')
       import :: c_int, c_ptr, c_double
       type(c_ptr), value :: ctx
       real(c_double), intent(out) :: neutralMass, totalMass   ! neutralHydrogenMass, totalHydrogenMass [msun]
       type(c_ptr), value :: stream
     end function rtb200_compute_mass
     integer(c_int) function rtb200_grid_get_species(ctx, HI, HeI, HeII) bind(C, name='rtb200_grid_get_species')
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx, HI, HeI, HeII
     end function rtb200_grid_get_species
     integer(c_int) function rtb200_point(ctx, nWave, wavelength, lum, metallicity, coefSpectrum, aDust, &
          dustApproximation, maxPixelLevel, nsrc, srcLeaf, srcWeight, k24, k25, k26, c24, c25, c26, &
          ndotRemaining, ndotBoundary, ndotDust, ndotSpectrum, nseg) bind(C, name='rtb200_point')
       import :: c_int, c_int32_t, c_ptr, c_double
       type(c_ptr), value :: ctx
       integer(c_int), value :: nWave
       type(c_ptr), value :: wavelength, lum, metallicity
       real(c_double), value :: coefSpectrum
       type(c_ptr), value :: aDust
       integer(c_int), value :: dustApproximation, maxPixelLevel
       integer(c_int32_t), value :: nsrc
       type(c_ptr), value :: srcLeaf, srcWeight, k24, k25, k26, c24, c25, c26
       type(c_ptr), value :: ndotRemaining, ndotBoundary, ndotDust, ndotSpectrum, nseg
     end function rtb200_point
  end interface

contains

  subroutine rtbCheck(status, where)
    integer(c_int), intent(in) :: status
    character(len=*), intent(in) :: where
    if (status.ne.0) then
       write(*,*) 'rtb200 error in ', where, ' status =', status
       stop
    endif
  end subroutine rtbCheck

  ! depth-first leaf order of writeCell (equiSources.f90:4044-4079): children i, j, k
  recursive subroutine countLeaves(currentCell, n)
    type(zoneType) :: currentCell
    integer(c_int64_t), intent(inout) :: n
    integer :: i, j, k
    if (currentCell%refined) then
       do i = 1, 2
          do j = 1, 2
             do k = 1, 2
                call countLeaves(currentCell%cell(i,j,k), n)
             enddo
          enddo
       enddo
    else
       n = n + 1
    endif
  end subroutine countLeaves

  recursive subroutine flattenCell(currentCell, level)
    type(zoneType) :: currentCell
    integer, intent(in) :: level
    integer :: i, j, k
    if (currentCell%refined) then
       do i = 1, 2
          do j = 1, 2
             do k = 1, 2
                call flattenCell(currentCell%cell(i,j,k), level+1)
             enddo
          enddo
       enddo
    else
       icursor = icursor + 1
       flatLevel(icursor) = int(level, c_int8_t)
       flatHI(icursor) = currentCell%HI
       flatHeI(icursor) = currentCell%HeI
       flatHeII(icursor) = currentCell%HeII
       flatRho(icursor) = currentCell%rho
       flatAbun2(icursor) = currentCell%abun2
    endif
  end subroutine flattenCell

  recursive subroutine scatterJ(currentCell)
    type(zoneType) :: currentCell
    integer :: i, j, k
    if (currentCell%refined) then
       do i = 1, 2
          do j = 1, 2
             do k = 1, 2
                call scatterJ(currentCell%cell(i,j,k))
             enddo
          enddo
       enddo
    else
       icursor = icursor + 1
       currentCell%Jmean1 = flatJ1(icursor)
       currentCell%Jmean2 = flatJ2(icursor)
       currentCell%Jmean3 = flatJ3(icursor)
    endif
  end subroutine scatterJ

  ! once, after the octree is built (equiSources.f90:628) -- and again whenever its topology changes
  subroutine rtbSetGrid(nx, ny, nz)
    integer, intent(in) :: nx, ny, nz
    integer :: i, j, k
    integer(c_int) :: device
    if (nx.ne.ny .or. nx.ne.nz) then
       write(*,*) 'rtb200: cubic base grid required'
       stop
    endif
    device = 0
    if (.not.c_associated(rtbContext)) call rtbCheck(rtb200_create(device, rtbContext), 'rtb200_create')
    rtbLeaves = 0
    do i = 1, nx
       do j = 1, ny
          do k = 1, nz
             call countLeaves(baseGrid%cell(i,j,k), rtbLeaves)
          enddo
       enddo
    enddo
    if (allocated(flatLevel)) deallocate(flatLevel, flatHI, flatHeI, flatHeII, flatRho, flatAbun2, flatJ1, flatJ2, flatJ3)
    allocate(flatLevel(rtbLeaves), flatHI(rtbLeaves), flatHeI(rtbLeaves), flatHeII(rtbLeaves), flatRho(rtbLeaves), &
         flatAbun2(rtbLeaves), flatJ1(rtbLeaves), flatJ2(rtbLeaves), flatJ3(rtbLeaves))
    icursor = 0
    do i = 1, nx
       do j = 1, ny
          do k = 1, nz
             call flattenCell(baseGrid%cell(i,j,k), 0)
          enddo
       enddo
    enddo
    call rtbCheck(rtb200_grid_set(rtbContext, int(nx, c_int), rtbLeaves, c_loc(flatLevel), c_loc(flatHI), &
         c_loc(flatHeI), c_loc(flatHeII), c_loc(flatRho), c_loc(flatAbun2), real(physicalBoxSize, c_double)), &
         'rtb200_grid_set')
  end subroutine rtbSetGrid

  ! replaces equiSources.f90:1372-1808 (runUVBTransfer block) inside the outer loop
  subroutine rtbDiffuse(nx, ny, nz)
    integer, intent(in) :: nx, ny, nz
    integer :: i, j, k
    real(c_double), dimension(3), target :: uvb
    real(c_double), dimension(9), target :: beta
    uvb = (/ uvb1, uvb2, uvb3 /)
    ! [group][beta24, beta26, beta25] as computeOpacities uses them (equiSources.f90:4974-4977)
    beta = (/ group1%beta24, group1%beta26, group1%beta25, &
              group2%beta24, group2%beta26, group2%beta25, &
              group3%beta24, group3%beta26, group3%beta25 /)
    ! chemistry changed HI, HeI, HeII since the last call (equiSources.f90:3671-3673)
    icursor = 0
    do i = 1, nx
       do j = 1, ny
          do k = 1, nz
             call flattenCell(baseGrid%cell(i,j,k), 0)
          enddo
       enddo
    enddo
    call rtbCheck(rtb200_grid_update_species(rtbContext, c_loc(flatHI), c_loc(flatHeI), c_loc(flatHeII)), &
         'rtb200_grid_update_species')
    call rtbCheck(rtb200_diffuse(rtbContext, int(nAngularLevel, c_int), c_loc(uvb), c_loc(beta), c_null_ptr, &
         0_c_int32_t, c_loc(flatJ1), c_loc(flatJ2), c_loc(flatJ3), c_null_ptr), 'rtb200_diffuse')
    icursor = 0
    do i = 1, nx
       do j = 1, ny
          do k = 1, nz
             call scatterJ(baseGrid%cell(i,j,k))
          enddo
       enddo
    enddo
  end subroutine rtbDiffuse

  ! ---- point sources ---------------------------------------------------------------------------------------

  recursive subroutine scatterRates(currentCell)
    type(zoneType) :: currentCell
    integer :: i, j, k
    if (currentCell%refined) then
       do i = 1, 2
          do j = 1, 2
             do k = 1, 2
                call scatterRates(currentCell%cell(i,j,k))
             enddo
          enddo
       enddo
    else
       icursor = icursor + 1
       currentCell%krate24 = flatK24(icursor)
       currentCell%krate25 = flatK25(icursor)
       currentCell%krate26 = flatK26(icursor)
       currentCell%crate24 = flatC24(icursor)
       currentCell%crate25 = flatC25(icursor)
       currentCell%crate26 = flatC26(icursor)
    endif
  end subroutine scatterRates

  ! leaf number (0-based, writeCell order) of a star's host cell from its call sequence star%position
  ! (equiSources.f90:753-756): leaves of the preceding base cells + leaves of the preceding siblings on every level
  function rtbLeafOfStar(currentStar) result(leaf)
    type(starType), intent(in) :: currentStar
    integer(c_int64_t) :: leaf, n
    type(zoneType), pointer :: cell
    integer :: l, i, j, k, ii, jj, kk
    i = currentStar%position(1); j = currentStar%position(2); k = currentStar%position(3)
    leaf = baseFirstLeaf(i,j,k)
    cell => baseGrid%cell(i,j,k)
    do l = 1, currentStar%level
       i = currentStar%position(3*l+1); j = currentStar%position(3*l+2); k = currentStar%position(3*l+3)
       do ii = 1, 2
          do jj = 1, 2
             do kk = 1, 2
                if ((ii-1)*4+(jj-1)*2+kk-1 .lt. (i-1)*4+(j-1)*2+k-1) then
                   n = 0
                   call countLeaves(cell%cell(ii,jj,kk), n)
                   leaf = leaf + n
                endif
             enddo
          enddo
       enddo
       cell => cell%cell(i,j,k)
    enddo
  end function rtbLeafOfStar

  ! replaces equiSources.f90:1256-1370 (runStellarTransfer block): all sources with weight > 0 in one call.
  ! iSpectrum / coefSpectrum are the driver's locals computed at :1236-1242.
  subroutine rtbPoint(nx, ny, nz, iSpectrum, coefSpectrum, maxPixelLevel, nStarsSpecificAge)
    integer, intent(in) :: nx, ny, nz, iSpectrum, maxPixelLevel, nStarsSpecificAge
    real(kind=RealKind), intent(in) :: coefSpectrum
    integer :: i, j, k, iStar, nsrc, iradius, im
    integer(c_int64_t) :: n
    integer(c_int32_t), dimension(:), allocatable, target :: srcLeaf, srcWeight
    real(c_double), dimension(:,:), allocatable, target :: remaining, boundary, spectrum   ! (7,nsrc), (7,nsrc), (300,nsrc)
    real(c_double), dimension(:), allocatable, target :: dustEscape
    real(c_double), dimension(nWavelengths,2,nMetallicity), target :: lumPack   ! = C [5][2][nWave]
    real(c_double), dimension(5,7), target :: dustPack                         ! = C [7][5]
    real(c_double), dimension(nWavelengths), target :: wl
    real(c_double), dimension(nMetallicity), target :: met
    real(kind=RealKind) :: ndot1, fraction(7)

    if (.not.allocated(baseFirstLeaf)) then
       allocate(baseFirstLeaf(nx,ny,nz))
       allocate(flatK24(rtbLeaves), flatK25(rtbLeaves), flatK26(rtbLeaves), flatC24(rtbLeaves), flatC25(rtbLeaves), &
            flatC26(rtbLeaves))
       n = 0
       do i = 1, nx
          do j = 1, ny
             do k = 1, nz
                baseFirstLeaf(i,j,k) = n
                call countLeaves(baseGrid%cell(i,j,k), n)
             enddo
          enddo
       enddo
    endif
    nsrc = 0
    do iStar = 1, nStars
       if (star(iStar)%weight.gt.0) nsrc = nsrc + 1
    enddo
    allocate(srcLeaf(nsrc), srcWeight(nsrc), remaining(7,nsrc), boundary(7,nsrc), spectrum(300,nsrc), dustEscape(nsrc))
    nsrc = 0
    do iStar = 1, nStars
       if (star(iStar)%weight.gt.0) then
          nsrc = nsrc + 1
          srcLeaf(nsrc) = int(rtbLeafOfStar(star(iStar)), c_int32_t)
          srcWeight(nsrc) = star(iStar)%weight
       endif
    enddo
    do im = 1, nMetallicity
       lumPack(:,1,im) = specificLuminosity(im,iSpectrum,:)
       lumPack(:,2,im) = specificLuminosity(im,iSpectrum+1,:)
    enddo
    dustPack = transpose(a_smc)
    wl = wavelength
    met = metallicity
    ! absorber densities may have changed since rtbSetGrid (previous chemistry step)
    icursor = 0
    do i = 1, nx
       do j = 1, ny
          do k = 1, nz
             call flattenCell(baseGrid%cell(i,j,k), 0)
          enddo
       enddo
    enddo
    call rtbCheck(rtb200_grid_update_species(rtbContext, c_loc(flatHI), c_loc(flatHeI), c_loc(flatHeII)), &
         'rtb200_grid_update_species')
    flatK24 = 0.; flatK25 = 0.; flatK26 = 0.; flatC24 = 0.; flatC25 = 0.; flatC26 = 0.   ! setZeroRates (:1246)
    call rtbCheck(rtb200_point(rtbContext, int(nWavelengths, c_int), c_loc(wl), c_loc(lumPack), c_loc(met), &
         real(coefSpectrum, c_double), c_loc(dustPack), int(dustApproximation, c_int), int(maxPixelLevel, c_int), &
         int(nsrc, c_int32_t), c_loc(srcLeaf), c_loc(srcWeight), c_loc(flatK24), c_loc(flatK25), c_loc(flatK26), &
         c_loc(flatC24), c_loc(flatC25), c_loc(flatC26), c_loc(remaining), c_loc(boundary), c_loc(dustEscape), &
         c_loc(spectrum), c_null_ptr), 'rtb200_point')
    icursor = 0
    do i = 1, nx
       do j = 1, ny
          do k = 1, nz
             call scatterRates(baseGrid%cell(i,j,k))
          enddo
       enddo
    enddo
    ! per-source escape fractions and the escaping spectrum (equiSources.f90:1342-1366)
    cosmicSpectrum = 0.
    nsrc = 0
    do iStar = 1, nStars
       if (star(iStar)%weight.gt.0) then
          nsrc = nsrc + 1
          ndot1 = float(star(iStar)%weight)
          do iradius = 1, 7
             if (boundary(iradius,nsrc).lt.1.) then
                fraction(iradius) = remaining(iradius,nsrc)/(ndot1-boundary(iradius,nsrc))
             else
                fraction(iradius) = 0.
             endif
          enddo
          cosmicSpectrum = cosmicSpectrum + float(star(iStar)%weight) * spectrum(:,nsrc)/(ndot1-boundary(7,nsrc))
          write(*,1015) iStar, star(iStar)%level, fraction, star(iStar)%weight
       endif
    enddo
1015 format('src: ', i5, i3, 7f9.5, i8)
    cosmicSpectrum = cosmicSpectrum / float(nStarsSpecificAge)
    deallocate(srcLeaf, srcWeight, remaining, boundary, spectrum, dustEscape)
  end subroutine rtbPoint

end module rtb200_shim
