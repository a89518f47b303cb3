#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY.  Writes a copy of the reference driver (equiSources.f90) with four insertions that dump,
in double precision, what the transport hot path reads and what it leaves behind.  Nothing of the reference's own code
is altered; the copy goes to oracle/_ref/ (git-ignored), never into the repository.

    python patch_driver.py /path/to/reference/equiSources.f90 oracle/_ref/src/equiSources_harness.f90

Insertions (anchored on the reference's own lines, checked before patching):
  1. before `maxDepth = 0.` (after the setZeroRates loop, equiSources.f90:1246-1254): `call rtbHarnessDump(0)` -- inputs.
  2. after the `src:` line of every source (`:1353-1357`): `call rtbHarnessSource(iStar)` -- escape diagnostics.
  3. before `open(unit=51, file='time' ...` (`:1833`): `call rtbHarnessDump(1)` then `stop` -- Jmean1..3, the six rate
     fields and the species after solveRateEquations of the FIRST outer iteration (the driver's loop never ends).
  4. before `end program pointTransfer`: the internal procedures below (host association gives them the driver's state).
Dump layout: see the comments in HARNESS_PROCEDURES and oracle/ref_harness/compare.py, which reads it.
"""
import sys

HARNESS_PROCEDURES = '''
  ! ---- rtb200 reference harness (inserted by oracle/ref_harness/patch_driver.py; not part of the reference) ----------
  recursive subroutine rtbHarnessCount(currentCell, n)
    implicit none
    type(zoneType), target :: currentCell
    integer, intent(inout) :: n
    integer :: i, j, k
    if (currentCell%refined) then
       do i = 1, 2
          do j = 1, 2
             do k = 1, 2
                call rtbHarnessCount(currentCell%cell(i,j,k), n)
             enddo
          enddo
       enddo
    else
       n = n + 1
    endif
  end subroutine rtbHarnessCount

  ! what = 0: level (int32); 1..15: one double field per call, leaves in writeCell order (children i, j, k)
  recursive subroutine rtbHarnessField(currentCell, level, what, unitNo)
    implicit none
    type(zoneType), target :: currentCell
    integer, intent(in) :: level, what, unitNo
    integer :: i, j, k
    if (currentCell%refined) then
       do i = 1, 2
          do j = 1, 2
             do k = 1, 2
                call rtbHarnessField(currentCell%cell(i,j,k), level+1, what, unitNo)
             enddo
          enddo
       enddo
    else
       select case (what)
       case (0);  write(unitNo) int(level, 4)
       case (1);  write(unitNo) currentCell%HI
       case (2);  write(unitNo) currentCell%HeI
       case (3);  write(unitNo) currentCell%HeII
       case (4);  write(unitNo) currentCell%rho
       case (5);  write(unitNo) currentCell%abun2
       case (6);  write(unitNo) currentCell%tgas
       case (7);  write(unitNo) currentCell%Jmean1
       case (8);  write(unitNo) currentCell%Jmean2
       case (9);  write(unitNo) currentCell%Jmean3
       case (10); write(unitNo) currentCell%krate24
       case (11); write(unitNo) currentCell%krate25
       case (12); write(unitNo) currentCell%krate26
       case (13); write(unitNo) currentCell%crate24
       case (14); write(unitNo) currentCell%crate25
       case (15); write(unitNo) currentCell%crate26
       end select
    endif
  end subroutine rtbHarnessField

  subroutine rtbHarnessAllLeaves(what, unitNo)
    implicit none
    integer, intent(in) :: what, unitNo
    integer :: i, j, k
    do i = 1, nx
       do j = 1, ny
          do k = 1, nz
             call rtbHarnessField(baseGrid%cell(i,j,k), 0, what, unitNo)
          enddo
       enddo
    enddo
  end subroutine rtbHarnessAllLeaves

  ! leaf number (0-based, writeCell order) of a cell: counts the leaves visited before it
  recursive subroutine rtbHarnessFind(currentCell, target_, n, found)
    implicit none
    type(zoneType), target :: currentCell
    type(zoneType), pointer :: target_
    integer, intent(inout) :: n
    logical, intent(inout) :: found
    integer :: i, j, k
    if (found) return
    if (currentCell%refined) then
       do i = 1, 2
          do j = 1, 2
             do k = 1, 2
                call rtbHarnessFind(currentCell%cell(i,j,k), target_, n, found)
             enddo
          enddo
       enddo
    else
       if (associated(target_, currentCell)) then
          found = .true.
       else
          n = n + 1
       endif
    endif
  end subroutine rtbHarnessFind

  ! stage 0: everything the transport reads.  stage 1: everything it (and the chemistry after it) leaves behind.
  subroutine rtbHarnessDump(stage)
    implicit none
    integer, intent(in) :: stage
    integer :: i, j, k, n, im, what
    if (stage.eq.0) then
       open(unit=91, file='ftte_reference_dump.bin', form='unformatted', access='stream', status='replace')
       open(unit=92, file='ftte_reference_sources.bin', form='unformatted', access='stream', status='replace')
       n = 0
       do i = 1, nx
          do j = 1, ny
             do k = 1, nz
                call rtbHarnessCount(baseGrid%cell(i,j,k), n)
             enddo
          enddo
       enddo
       ! header: 10 x int32
       write(91) int(20260001, 4), int(nx, 4), int(n, 4), int(nStars, 4), int(nWavelengths, 4), int(nMetallicity, 4), &
            int(iSpectrum, 4), int(dustApproximation, 4), int(maxPixelLevel, 4), int(nAngularLevel, 4)
       ! scalars: 9 doubles, then beta24/25/26 and ksi24/25/26 of the three groups (18 doubles)
       write(91) physicalBoxSize, coefSpectrum, currentRedshift, uvb1, uvb2, uvb3, alpha(1), alpha(2), alpha(3)
       write(91) group1%beta24, group1%beta25, group1%beta26, group2%beta24, group2%beta25, group2%beta26, &
            group3%beta24, group3%beta25, group3%beta26
       write(91) group1%ksi24, group1%ksi25, group1%ksi26, group2%ksi24, group2%ksi25, group2%ksi26, &
            group3%ksi24, group3%ksi25, group3%ksi26
       ! population synthesis tables as the source loop uses them: [metallicity][2 time slices][wavelength]
       write(91) wavelength
       write(91) metallicity
       do im = 1, nMetallicity
          write(91) specificLuminosity(im, iSpectrum, :)
          write(91) specificLuminosity(im, iSpectrum+1, :)
       enddo
       do i = 1, 7
          write(91) a_smc(i, :)
       enddo
       ! stars: weight of every star (the host leaf is written with the diagnostics, once it is known)
       do i = 1, nStars
          write(91) int(star(i)%weight, 4)
       enddo
       do what = 0, 6
          call rtbHarnessAllLeaves(what, 91)
       enddo
    else
       do what = 7, 15
          call rtbHarnessAllLeaves(what, 91)
       enddo
       do what = 1, 3
          call rtbHarnessAllLeaves(what, 91)      ! HI, HeI, HeII after solveRateEquations
       enddo
       close(91)
       close(92)
    endif
  end subroutine rtbHarnessDump

  ! after the rays of source iStar: host leaf, highestPixelLevel, escape diagnostics
  subroutine rtbHarnessSource(iStarArg)
    implicit none
    integer, intent(in) :: iStarArg
    integer :: i, j, k, n
    logical :: found
    n = 0
    found = .false.
    do i = 1, nx
       do j = 1, ny
          do k = 1, nz
             call rtbHarnessFind(baseGrid%cell(i,j,k), star(iStarArg)%hostCell, n, found)
          enddo
       enddo
    enddo
    if (.not.found) n = -1
    write(92) int(iStarArg, 4), int(n, 4), int(star(iStarArg)%weight, 4), int(highestPixelLevel, 4)
    write(92) ndotRemaining, ndotBoundary, ndotDust
    write(92) ndotSpectrum
  end subroutine rtbHarnessSource
  ! ---- end of the rtb200 reference harness -----------------------------------------------------------------------------

'''


def patch(src_lines):
    out = []
    state = dict(dump0=False, source=False, dump1=False, procs=False)
    in_src_write = False
    for i, line in enumerate(src_lines):
        stripped = line.strip()
        if not state["dump0"] and stripped == "maxDepth = 0.":
            out.append("     call rtbHarnessDump(0)   ! rtb200 harness\n")
            state["dump0"] = True
        if state["dump0"] and not state["dump1"] and stripped.startswith("open(unit=51, file='time'"):
            out.append("     call rtbHarnessDump(1)   ! rtb200 harness\n")
            out.append("     stop                     ! rtb200 harness: one outer iteration\n")
            state["dump1"] = True
        if not state["procs"] and stripped == "end program pointTransfer":
            out.append(HARNESS_PROCEDURES)
            state["procs"] = True
        out.append(line)
        if not state["source"]:
            if stripped.startswith("write(*,1015) iStar"):
                in_src_write = True
            if in_src_write and not stripped.endswith("&"):
                out.append("              call rtbHarnessSource(iStar)   ! rtb200 harness\n")
                state["source"] = True
                in_src_write = False
    missing = [k for k, v in state.items() if not v]
    if missing:
        raise SystemExit(f"patch_driver: anchors not found in the reference source: {missing}")
    return out


if __name__ == "__main__":
    if len(sys.argv) != 3:
        raise SystemExit(__doc__)
    with open(sys.argv[1]) as f:
        lines = f.readlines()
    with open(sys.argv[2], "w") as f:
        f.writelines(patch(lines))
    print("wrote", sys.argv[2])
