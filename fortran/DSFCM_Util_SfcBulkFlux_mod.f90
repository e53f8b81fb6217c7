!> Drop-in for sfc/DSFCM_Util_SfcBulkFlux_mod.f90: same module name, same public procedures,
!! same dummy list of DSFCM_Util_SfcBulkFlux_Get (ref :108-151).  IA, JA still come from
!! DSFCM_Admin_Grid_mod, the state arrays from DSFCM_Admin_Variable_mod (both unchanged).
module DSFCM_Util_SfcBulkFlux_mod
  use dc_types, only: DP
  use DSFCM_Admin_Grid_mod, only: IA, JA
  use DSFCM_Admin_Variable_mod, only: SFC_PROP_MAX
  use dccm_b200_c
  implicit none
  private
  public :: DSFCM_Util_SfcBulkFlux_Init, DSFCM_Util_SfcBulkFlux_Final, DSFCM_Util_SfcBulkFlux_Get
contains
  subroutine DSFCM_Util_SfcBulkFlux_Init()
    ! limits of the reference's read_config (ref :574-591) are compiled into the kernel
  end subroutine
  subroutine DSFCM_Util_SfcBulkFlux_Final()
  end subroutine

  subroutine DSFCM_Util_SfcBulkFlux_Get( &
       & xya_WindStressX, xya_WindStressY, xya_SenHFlx, xya_QVapMFlx, xya_LatHFlx,  &
       & xya_SfcVelTransCoef, xya_SfcTempTransCoef, xya_SfcQVapTransCoef, &
       & xya_DelVarImplCPL, xya_SUwRFlx, xya_LUwRFlx, &
       & xya_SfcHFlx_ns, xya_SfcHFlx_sr, xya_DSfcHFlxDTs, &
       & xy_WindU, xy_WindV, xy_SfcAirTemp, xy_QVap1, xy_SDwRFlx, xy_LDwRFlx, &
       & xya_ImplCplCoef1, xya_ImplCplCoef2, xya_SfcTemp, xya_SfcAlbedo, xy_SIceCon, &
       & a_Sig1Info, xy_SfcHeight, xy_SfcPress )
    real(DP), intent(out) :: xya_WindStressX(IA,JA,SFC_PROP_MAX), xya_WindStressY(IA,JA,SFC_PROP_MAX)
    real(DP), intent(out) :: xya_SenHFlx(IA,JA,SFC_PROP_MAX), xya_QVapMFlx(IA,JA,SFC_PROP_MAX), xya_LatHFlx(IA,JA,SFC_PROP_MAX)
    real(DP), intent(out) :: xya_SfcVelTransCoef(IA,JA,SFC_PROP_MAX), xya_SfcTempTransCoef(IA,JA,SFC_PROP_MAX)
    real(DP), intent(out) :: xya_SfcQVapTransCoef(IA,JA,SFC_PROP_MAX)
    real(DP), intent(out) :: xya_SfcHFlx_ns(IA,JA,SFC_PROP_MAX), xya_SfcHFlx_sr(IA,JA,SFC_PROP_MAX), xya_DSfcHFlxDTs(IA,JA,SFC_PROP_MAX)
    real(DP), intent(out) :: xya_DelVarImplCPL(IA,JA,4)
    real(DP), intent(out) :: xya_SUwRFlx(IA,JA,SFC_PROP_MAX), xya_LUwRFlx(IA,JA,SFC_PROP_MAX)
    real(DP), intent(in) :: xy_WindU(IA,JA), xy_WindV(IA,JA), xy_SfcAirTemp(IA,JA), xy_QVap1(IA,JA)
    real(DP), intent(in) :: xy_SDwRFlx(IA,JA), xy_LDwRFlx(IA,JA)
    real(DP), intent(in) :: xya_ImplCplCoef1(IA,JA,4), xya_ImplCplCoef2(IA,JA,4)
    real(DP), intent(inout) :: xya_SfcTemp(IA,JA,SFC_PROP_MAX), xya_SfcAlbedo(IA,JA,SFC_PROP_MAX)
    real(DP), intent(in) :: xy_SIceCon(IA,JA), a_Sig1Info(2), xy_SfcHeight(IA,JA), xy_SfcPress(IA,JA)

    call dccm_check( dccm_bulkflux_get_host(IA, JA, &
         & xya_WindStressX, xya_WindStressY, xya_SenHFlx, xya_QVapMFlx, xya_LatHFlx, &
         & xya_SfcVelTransCoef, xya_SfcTempTransCoef, xya_SfcQVapTransCoef, xya_DelVarImplCPL, &
         & xya_SUwRFlx, xya_LUwRFlx, xya_SfcHFlx_ns, xya_SfcHFlx_sr, xya_DSfcHFlxDTs, &
         & xy_WindU, xy_WindV, xy_SfcAirTemp, xy_QVap1, xy_SDwRFlx, xy_LDwRFlx, &
         & xya_ImplCplCoef1, xya_ImplCplCoef2, xya_SfcTemp, xya_SfcAlbedo, xy_SIceCon, &
         & a_Sig1Info, xy_SfcHeight, xy_SfcPress), "DSFCM_Util_SfcBulkFlux_Get")
  end subroutine DSFCM_Util_SfcBulkFlux_Get
end module DSFCM_Util_SfcBulkFlux_mod
