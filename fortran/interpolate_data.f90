!> Drop-in for common/interpolate_data.f90 (ref :1-17): the bare external symbol
!! `interpolate_data_` Jcup links against and calls back.  Same dummy list; the work is done by
!! libdccm_b200 (CSR gather-SpMV on the GPU, results back in recv_data on return).
subroutine interpolate_data(recv_model, send_model, mapping_tag, sn1, sn2, send_data, &
                            rn1, rn2, recv_data, num_of_data, tn, exchange_tag)
  use jcup_interface, only: jcup_get_comp_num_from_name
  use dccm_b200_c
  implicit none
  character(len=*), intent(IN) :: recv_model, send_model
  integer, intent(IN) :: mapping_tag
  integer, intent(IN) :: sn1, sn2
  real(kind=8), intent(IN) :: send_data(sn1,sn2)
  integer, intent(IN) :: rn1, rn2
  real(kind=8), intent(INOUT) :: recv_data(rn1,rn2)
  integer, intent(IN) :: num_of_data
  integer, intent(IN) :: tn
  integer, intent(IN) :: exchange_tag(tn)

  call dccm_check( dccm_interpolate_data( &
       & jcup_get_comp_num_from_name(recv_model), jcup_get_comp_num_from_name(send_model), mapping_tag, &
       & sn1, sn2, send_data, rn1, rn2, recv_data, num_of_data), "interpolate_data")
end subroutine interpolate_data
