!> Library-backed replacement of module grid_mapping_util_jones99 (reference: common/grid_mapping_util_jones99.f90).
!! Same public names and dummy lists (gen_gridmapfile_lonlat2lonlat :35-54, set_mappingTable_interpCoef :446-462), so
!! tool/gmapgen and the component glue compile against it unchanged; the work is done by libdccm_b200's host-side
!! generator / reader (index-for-index the reference algorithm, without gtool5's FileOpen and without the per-cell
!! print to standard output).  Optional: the reference module keeps working as it is -- this one removes the O(n) text
!! bookkeeping in Fortran and gives access to lon_mode = 1 (generalised longitude overlap) through DCCM_LON_MODE.
module grid_mapping_util_jones99
  use iso_c_binding
  use dccm_b200_c
  implicit none
  private
  public :: gen_gridmapfile_lonlat2lonlat, set_mappingTable_interpCoef
  integer, parameter :: DP = c_double
  integer, public, save :: DCCM_LON_MODE = 0        ! 0 = the reference's behaviour (equal longitudes or nx == 1)

contains

  subroutine gen_gridmapfile_lonlat2lonlat( filename, x_LonS, y_LatS, x_LonD, y_LatD, &
       & x_LonIntWtS, y_LatIntWtS, x_LonIntWtD, y_LatIntWtD, accuracy_order )
    character(*), intent(in) :: filename
    real(DP), intent(in) :: x_LonS(:), y_LatS(:), x_LonD(:), y_LatD(:)
    real(DP), intent(in) :: x_LonIntWtS(:), y_LatIntWtS(:), x_LonIntWtD(:), y_LatIntWtD(:)   ! longitude weights unused (ref :325)
    integer, intent(in) :: accuracy_order
    type(c_ptr) :: table

    call dccm_check( dccm_table_gen_jones99( &
         & int(size(x_LonS), c_int), x_LonS, int(size(y_LatS), c_int), y_LatS,  &
         & int(size(x_LonD), c_int), x_LonD, int(size(y_LatD), c_int), y_LatD,  &
         & y_LatIntWtS, y_LatIntWtD, int(accuracy_order, c_int), int(DCCM_LON_MODE, c_int), table ), &
         & "gen_gridmapfile_lonlat2lonlat" )
    call dccm_check( dccm_table_write_text( table, trim(filename)//c_null_char ), "gen_gridmapfile_lonlat2lonlat: write" )
    call dccm_table_free( table )
  end subroutine gen_gridmapfile_lonlat2lonlat

  subroutine set_mappingTable_interpCoef( gridmapfile, GNXS, GNXR, send_index, recv_index, coef_s )
    character(*), intent(in) :: gridmapfile
    integer, intent(in) :: GNXS, GNXR
    integer, intent(inout), allocatable :: send_index(:), recv_index(:)
    real(DP), intent(inout), allocatable :: coef_s(:)
    type(c_ptr) :: table
    integer :: n

    call dccm_check( dccm_table_read_text( trim(gridmapfile)//c_null_char, table ), "set_mappingTable_interpCoef: read" )
    n = int( dccm_table_size( table ) )
    if (allocated(coef_s)) deallocate(coef_s)                  ! as the reference does (ref :487)
    if (allocated(send_index)) deallocate(send_index)
    if (allocated(recv_index)) deallocate(recv_index)
    allocate( send_index(n), recv_index(n), coef_s(n) )
    call dccm_check( dccm_table_index( table, int(GNXS, c_int), int(GNXR, c_int), send_index, recv_index, coef_s ), &
         & "set_mappingTable_interpCoef" )
    call dccm_table_free( table )
  end subroutine set_mappingTable_interpCoef

end module grid_mapping_util_jones99
