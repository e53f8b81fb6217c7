!> Drop-in for atm/dcpam_sfc_implicit_coupling_mod.f90: same module name, same three public
!! procedures with the same dummy lists (ref :25-27, :72-79, :420).  The swept matrices that the
!! reference keeps in module `save` arrays (ref :16-18) live on the device inside `handle`.
module dcpam_sfc_implicit_coupling_mod
  use gridset, only: imax, jmax, kmax
  use composition, only: ncmax, IndexH2Ovap
  use dc_types, only: DP
  use dccm_b200_c
  implicit none
  private
  public :: dcpam_sfc_implicit_coupling_Init, SfcImplicitCoupling_VDiffBackward, SfcImplicitCoupling_VDiffForward
  type(c_ptr), save :: handle = c_null_ptr
contains
  subroutine dcpam_sfc_implicit_coupling_Init()
    use constants, only: Grav, CpDry, GasRDry
    use timeset, only: DelTime
    call dccm_check( dccm_vdiff_create(imax, jmax, kmax, ncmax, IndexH2OVap, Grav, CpDry, GasRDry, DelTime, handle), &
         & "dcpam_sfc_implicit_coupling_Init")
  end subroutine

  subroutine SfcImplicitCoupling_VDiffForward( &
    & xyr_MomFluxX, xyr_MomFluxY, xyr_HeatFlux, xyrf_QMixFlux, xyr_Press, xyz_Exner, xyr_Exner, &
    & xyr_VirTemp, xyz_Height, xyr_VelDiffCoef, xyr_TempDiffCoef, xyr_QMixDiffCoef, &
    & xyz_DUDt, xyz_DVDt, xyz_DTempDt, xyzf_DQMixDt, xya_ImplCplCoef1, xya_ImplCplCoef2 )
    real(DP), intent(in) :: xyr_MomFluxX(0:imax-1,1:jmax,0:kmax), xyr_MomFluxY(0:imax-1,1:jmax,0:kmax)
    real(DP), intent(in) :: xyr_HeatFlux(0:imax-1,1:jmax,0:kmax), xyrf_QMixFlux(0:imax-1,1:jmax,0:kmax,1:ncmax)
    real(DP), intent(in) :: xyr_Press(0:imax-1,1:jmax,0:kmax), xyz_Exner(0:imax-1,1:jmax,1:kmax), xyr_Exner(0:imax-1,1:jmax,0:kmax)
    real(DP), intent(in) :: xyr_VirTemp(0:imax-1,1:jmax,0:kmax), xyz_Height(0:imax-1,1:jmax,1:kmax)
    real(DP), intent(in) :: xyr_VelDiffCoef(0:imax-1,1:jmax,0:kmax), xyr_TempDiffCoef(0:imax-1,1:jmax,0:kmax)
    real(DP), intent(in) :: xyr_QMixDiffCoef(0:imax-1,1:jmax,0:kmax)
    real(DP), intent(out) :: xyz_DUDt(0:imax-1,1:jmax,1:kmax), xyz_DVDt(0:imax-1,1:jmax,1:kmax)
    real(DP), intent(out) :: xyz_DTempDt(0:imax-1,1:jmax,1:kmax), xyzf_DQMixDt(0:imax-1,1:jmax,1:kmax,1:ncmax)
    real(DP), intent(out) :: xya_ImplCplCoef1(0:imax-1,1:jmax,4), xya_ImplCplCoef2(0:imax-1,1:jmax,4)
    call dccm_check( dccm_vdiff_forward_host(handle, xyr_MomFluxX, xyr_MomFluxY, xyr_HeatFlux, xyrf_QMixFlux, &
         & xyr_Press, xyz_Exner, xyr_Exner, xyr_VirTemp, xyz_Height, xyr_VelDiffCoef, xyr_TempDiffCoef, &
         & xyr_QMixDiffCoef, xyz_DUDt, xyz_DVDt, xyz_DTempDt, xyzf_DQMixDt, xya_ImplCplCoef1, xya_ImplCplCoef2), &
         & "SfcImplicitCoupling_VDiffForward")
  end subroutine

  subroutine SfcImplicitCoupling_VDiffBackward( xyz_DUDt, xyz_DVDt, xyz_DTempDt, xyzf_DQMixDt )
    real(DP), intent(inout) :: xyz_DUDt(0:imax-1,1:jmax,1:kmax), xyz_DVDt(0:imax-1,1:jmax,1:kmax)
    real(DP), intent(inout) :: xyz_DTempDt(0:imax-1,1:jmax,1:kmax), xyzf_DQMixDt(0:imax-1,1:jmax,1:kmax,1:ncmax)
    call dccm_check( dccm_vdiff_backward_host(handle, xyz_DUDt, xyz_DVDt, xyz_DTempDt, xyzf_DQMixDt), &
         & "SfcImplicitCoupling_VDiffBackward")
  end subroutine
end module dcpam_sfc_implicit_coupling_mod
