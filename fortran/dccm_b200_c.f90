!> ISO_C_BINDING interfaces of libdccm_b200.so (include/dccm_b200.h).
!! Source-only deliverable: there is no Fortran compiler in the build image; the C side of
!! every entry point below is exercised by tests/ through the same ABI.
module dccm_b200_c
  use iso_c_binding
  implicit none
  public

  interface
     function dccm_init(device) bind(C, name="dccm_init") result(rc)
       import; integer(c_int), value :: device; integer(c_int) :: rc
     end function
     !> page-lock a module array once at init (c_loc(array), bytes): lets the *_host calls overlap their copies
     function dccm_host_register(ptr, bytes) bind(C, name="dccm_host_register") result(rc)
       import; type(c_ptr), value :: ptr; integer(c_int64_t), value :: bytes; integer(c_int) :: rc
     end function
     function dccm_last_error() bind(C, name="dccm_last_error") result(msg)
       import; type(c_ptr) :: msg
     end function
     function dccm_remap_create(nops, send_index, recv_index, coef, n_send, n_recv, handle) &
          & bind(C, name="dccm_remap_create") result(rc)
       import
       integer(c_int64_t), value :: nops
       integer(c_int32_t), intent(in) :: send_index(*), recv_index(*)
       real(c_double), intent(in) :: coef(*)
       integer(c_int), value :: n_send, n_recv
       type(c_ptr), intent(out) :: handle
       integer(c_int) :: rc
     end function
     function dccm_interp_register(recv_model, send_model, mapping_tag, handle) &
          & bind(C, name="dccm_interp_register") result(rc)
       import; integer(c_int), value :: recv_model, send_model, mapping_tag
       type(c_ptr), value :: handle; integer(c_int) :: rc
     end function
     function dccm_interpolate_data(recv_model, send_model, mapping_tag, sn1, sn2, send_data, &
          & rn1, rn2, recv_data, num_of_data) bind(C, name="dccm_interpolate_data") result(rc)
       import
       integer(c_int), value :: recv_model, send_model, mapping_tag, sn1, sn2, rn1, rn2, num_of_data
       real(c_double), intent(in) :: send_data(sn1, *)
       real(c_double), intent(inout) :: recv_data(rn1, *)
       integer(c_int) :: rc
     end function
     function dccm_bulkflux_get_host(IA, JA, &
          & WindStressX, WindStressY, SenHFlx, QVapMFlx, LatHFlx, &
          & SfcVelTransCoef, SfcTempTransCoef, SfcQVapTransCoef, DelVarImplCPL, &
          & SUwRFlx, LUwRFlx, SfcHFlx_ns, SfcHFlx_sr, DSfcHFlxDTs, &
          & WindU, WindV, SfcAirTemp, QVap1, SDwRFlx, LDwRFlx, ImplCplCoef1, ImplCplCoef2, &
          & SfcTemp, SfcAlbedo, SIceCon, Sig1Info, SfcHeight, SfcPress) &
          & bind(C, name="dccm_bulkflux_get_host") result(rc)
       import
       integer(c_int), value :: IA, JA
       real(c_double) :: WindStressX(*), WindStressY(*), SenHFlx(*), QVapMFlx(*), LatHFlx(*)
       real(c_double) :: SfcVelTransCoef(*), SfcTempTransCoef(*), SfcQVapTransCoef(*), DelVarImplCPL(*)
       real(c_double) :: SUwRFlx(*), LUwRFlx(*), SfcHFlx_ns(*), SfcHFlx_sr(*), DSfcHFlxDTs(*)
       real(c_double), intent(in) :: WindU(*), WindV(*), SfcAirTemp(*), QVap1(*), SDwRFlx(*), LDwRFlx(*)
       real(c_double), intent(in) :: ImplCplCoef1(*), ImplCplCoef2(*)
       real(c_double) :: SfcTemp(*), SfcAlbedo(*)
       real(c_double), intent(in) :: SIceCon(*), Sig1Info(2), SfcHeight(*), SfcPress(*)
       integer(c_int) :: rc
     end function
     function dccm_vdiff_create(imax, jmax, kmax, ncmax, index_h2ovap, Grav, CpDry, GasRDry, DelTime, handle) &
          & bind(C, name="dccm_vdiff_create") result(rc)
       import
       integer(c_int), value :: imax, jmax, kmax, ncmax, index_h2ovap
       real(c_double), value :: Grav, CpDry, GasRDry, DelTime
       type(c_ptr), intent(out) :: handle
       integer(c_int) :: rc
     end function
     function dccm_vdiff_forward_host(handle, MomFluxX, MomFluxY, HeatFlux, QMixFlux, Press, zExner, rExner, &
          & VirTemp, Height, VelDiffCoef, TempDiffCoef, QMixDiffCoef, DUDt, DVDt, DTempDt, DQMixDt, &
          & ImplCplCoef1, ImplCplCoef2) bind(C, name="dccm_vdiff_forward_host") result(rc)
       import
       type(c_ptr), value :: handle
       real(c_double), intent(in) :: MomFluxX(*), MomFluxY(*), HeatFlux(*), QMixFlux(*), Press(*), zExner(*), rExner(*)
       real(c_double), intent(in) :: VirTemp(*), Height(*), VelDiffCoef(*), TempDiffCoef(*), QMixDiffCoef(*)
       real(c_double) :: DUDt(*), DVDt(*), DTempDt(*), DQMixDt(*), ImplCplCoef1(*), ImplCplCoef2(*)
       integer(c_int) :: rc
     end function
     function dccm_vdiff_backward_host(handle, DUDt, DVDt, DTempDt, DQMixDt) &
          & bind(C, name="dccm_vdiff_backward_host") result(rc)
       import
       type(c_ptr), value :: handle
       real(c_double) :: DUDt(*), DVDt(*), DTempDt(*), DQMixDt(*)
       integer(c_int) :: rc
     end function
     !> operator straight from the axis arrays gmapgen hands to gen_gridmapfile_lonlat2lonlat
     !! (ref common/grid_mapping_util_jones99.f90:35-54): no table file, separable form for different longitudes
     function dccm_remap_create_jones99(nxs, x_LonS, nys, y_LatS, nxd, x_LonD, nyd, y_LatD, &
          & y_LatIntWtS, y_LatIntWtD, accuracy_order, lon_mode, handle) &
          & bind(C, name="dccm_remap_create_jones99") result(rc)
       import
       integer(c_int), value :: nxs, nys, nxd, nyd, accuracy_order, lon_mode
       real(c_double), intent(in) :: x_LonS(nxs), y_LatS(nys), x_LonD(nxd), y_LatD(nyd), y_LatIntWtS(nys), y_LatIntWtD(nyd)
       type(c_ptr), intent(out) :: handle
       integer(c_int) :: rc
     end function
     !> same for the bilinear generator (ref common/grid_mapping_util.f90:32-48)
     function dccm_remap_create_bilinear(nxs, x_LonS, nys, y_LatS, nxr, x_LonR, nyr, y_LatR, lon_mode, handle) &
          & bind(C, name="dccm_remap_create_bilinear") result(rc)
       import
       integer(c_int), value :: nxs, nys, nxr, nyr, lon_mode
       real(c_double), intent(in) :: x_LonS(nxs), y_LatS(nys), x_LonR(nxr), y_LatR(nyr)
       type(c_ptr), intent(out) :: handle
       integer(c_int) :: rc
     end function
     !> the operator of ONE rank's latitude band (destination rows jD0 .. jD1-1, 0-based; source cells numbered inside the
     !! rank's source buffer of src_rows rows starting at row src_row0): what the receiving component's MPI rank needs
     !! (ref common/interpolation_data_latlon_mod.f90:140-151), straight from the grid axes
     function dccm_remap_create_jones99_band(nxs, x_LonS, nys, y_LatS, nxd, x_LonD, nyd, y_LatD, &
          & y_LatIntWtS, y_LatIntWtD, accuracy_order, lon_mode, jD0, jD1, src_row0, src_rows, handle) &
          & bind(C, name="dccm_remap_create_jones99_band") result(rc)
       import
       integer(c_int), value :: nxs, nys, nxd, nyd, accuracy_order, lon_mode, jD0, jD1, src_row0, src_rows
       real(c_double), intent(in) :: x_LonS(nxs), y_LatS(nys), x_LonD(nxd), y_LatD(nyd), y_LatIntWtS(nys), y_LatIntWtD(nyd)
       type(c_ptr), intent(out) :: handle
       integer(c_int) :: rc
     end function
     function dccm_remap_create_bilinear_band(nxs, x_LonS, nys, y_LatS, nxr, x_LonR, nyr, y_LatR, lon_mode, &
          & jD0, jD1, src_row0, src_rows, handle) bind(C, name="dccm_remap_create_bilinear_band") result(rc)
       import
       integer(c_int), value :: nxs, nys, nxr, nyr, lon_mode, jD0, jD1, src_row0, src_rows
       real(c_double), intent(in) :: x_LonS(nxs), y_LatS(nys), x_LonR(nxr), y_LatR(nyr)
       type(c_ptr), intent(out) :: handle
       integer(c_int) :: rc
     end function
     !> table generators, files and the index arithmetic of set_mappingTable_interpCoef
     !! (ref common/grid_mapping_util_jones99.f90:35-54, :446-506; common/grid_mapping_util.f90:32-48, :181-241)
     function dccm_table_gen_jones99(nxs, x_LonS, nys, y_LatS, nxd, x_LonD, nyd, y_LatD, &
          & y_LatIntWtS, y_LatIntWtD, accuracy_order, lon_mode, table) &
          & bind(C, name="dccm_table_gen_jones99") result(rc)
       import
       integer(c_int), value :: nxs, nys, nxd, nyd, accuracy_order, lon_mode
       real(c_double), intent(in) :: x_LonS(nxs), y_LatS(nys), x_LonD(nxd), y_LatD(nyd), y_LatIntWtS(nys), y_LatIntWtD(nyd)
       type(c_ptr), intent(out) :: table
       integer(c_int) :: rc
     end function
     function dccm_table_gen_bilinear(nxs, x_LonS, nys, y_LatS, nxr, x_LonR, nyr, y_LatR, lon_mode, table) &
          & bind(C, name="dccm_table_gen_bilinear") result(rc)
       import
       integer(c_int), value :: nxs, nys, nxr, nyr, lon_mode
       real(c_double), intent(in) :: x_LonS(nxs), y_LatS(nys), x_LonR(nxr), y_LatR(nyr)
       type(c_ptr), intent(out) :: table
       integer(c_int) :: rc
     end function
     function dccm_table_write_text(table, filename) bind(C, name="dccm_table_write_text") result(rc)
       import
       type(c_ptr), value :: table
       character(kind=c_char), intent(in) :: filename(*)
       integer(c_int) :: rc
     end function
     function dccm_table_read_text(filename, table) bind(C, name="dccm_table_read_text") result(rc)
       import
       character(kind=c_char), intent(in) :: filename(*)
       type(c_ptr), intent(out) :: table
       integer(c_int) :: rc
     end function
     function dccm_table_size(table) bind(C, name="dccm_table_size") result(n)
       import
       type(c_ptr), value :: table
       integer(c_int64_t) :: n
     end function
     function dccm_table_index(table, gnxs, gnxr, send_index, recv_index, coef_s) &
          & bind(C, name="dccm_table_index") result(rc)
       import
       type(c_ptr), value :: table
       integer(c_int), value :: gnxs, gnxr
       integer(c_int32_t) :: send_index(*), recv_index(*)
       real(c_double) :: coef_s(*)
       integer(c_int) :: rc
     end function
     subroutine dccm_table_free(table) bind(C, name="dccm_table_free")
       import
       type(c_ptr), value :: table
     end subroutine
  end interface

contains

  !> abort through the caller's error channel with the library's message (the reference aborts
  !! through jcup_error / MessageNotify('E', ...))
  subroutine dccm_check(rc, where)
    integer(c_int), intent(in) :: rc
    character(*), intent(in) :: where
    character(kind=c_char), pointer :: msg(:)
    integer :: n
    if (rc == 0) return
    call c_f_pointer(dccm_last_error(), msg, (/ 1024 /))
    n = 1
    do while (n < 1024 .and. msg(n) /= c_null_char); n = n + 1; end do
    write(0,*) trim(where), ": libdccm_b200 error ", rc, ": ", msg(1:n-1)
    stop 1
  end subroutine dccm_check

end module dccm_b200_c
