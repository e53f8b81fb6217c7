!> GPU-backed replacement of the lone external procedure Jcup links against and calls back to remap
!! (reference: common/interpolate_data.f90, 12 dummies + hidden string lengths; the object is linked
!! on its own, Mkinclude:167).  The dummy list is the coupler's contract and therefore identical;
!! the body only translates component names to Jcup component numbers and hands the two host arrays
!! to libdccm_b200, which runs the row-sorted CSR / zonal-stencil gather on the GPU and returns with
!! recv_data filled (all rn2 columns zeroed first, columns 1..num_of_data accumulated).
subroutine interpolate_data( recv_model, send_model, mapping_tag,  &
     &                      sn1, sn2, send_data,                  &
     &                      rn1, rn2, recv_data,                  &
     &                      num_of_data, tn, exchange_tag )

  use iso_c_binding,  only: c_int, c_double
  use jcup_interface, only: jcup_get_comp_num_from_name
  use dccm_b200_c,    only: dccm_interpolate_data, dccm_check

  implicit none

  ! -- who receives, who sends, which table ------------------------------------------------
  character(len=*), intent(in)    :: recv_model
  character(len=*), intent(in)    :: send_model
  integer,          intent(in)    :: mapping_tag
  ! -- extents first, then the arrays they shape ---------------------------------------------
  integer,          intent(in)    :: sn1, sn2, rn1, rn2
  integer,          intent(in)    :: num_of_data, tn
  real(c_double),   intent(in)    :: send_data(sn1, sn2)
  real(c_double),   intent(inout) :: recv_data(rn1, rn2)
  integer,          intent(in)    :: exchange_tag(tn)      ! unused, as in the reference

  integer(c_int) :: id_recv, id_send, status

  id_recv = jcup_get_comp_num_from_name(recv_model)
  id_send = jcup_get_comp_num_from_name(send_model)

  status = dccm_interpolate_data( id_recv, id_send, int(mapping_tag, c_int),       &
       &                          int(sn1, c_int), int(sn2, c_int), send_data,     &
       &                          int(rn1, c_int), int(rn2, c_int), recv_data,     &
       &                          int(num_of_data, c_int) )
  call dccm_check(status, "interpolate_data")

end subroutine interpolate_data
