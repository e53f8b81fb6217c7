!> Patch for common/interpolation_data_latlon_mod.f90: the two places where the reference has
!! both the local operation indices and the local coefficients in hand.  Everything else in the
!! module (Jcup calls, operation_index_type) stays as it is.
!!
!! Add at the end of set_interpolate_coef (ref :219-270), once coi%coefS is filled
!! (after jcup_set_local_coef / jcup_recv_coef):
!!
!!     call dccm_register_operation(recv_comp_name, send_comp_name, mapping_tag)
!!
!! and replace the body of interpolate_data_latlon (ref :293-302) by the call in
!! fortran/shim_interpolate_data.f90 (or keep it as a fallback-free thin wrapper).
subroutine dccm_register_operation(recv_comp_name, send_comp_name, mapping_tag)
  use interpolation_data_latlon_mod      ! needs `operation_index` made public (or move this inside the module)
  use jcup_interface, only: jcup_get_comp_num_from_name
  use dccm_b200_c
  implicit none
  character(*), intent(in) :: recv_comp_name, send_comp_name
  integer, intent(in) :: mapping_tag
  type(c_ptr) :: handle
  integer :: rid, sid, n_send, n_recv
  rid = jcup_get_comp_num_from_name(recv_comp_name)
  sid = jcup_get_comp_num_from_name(send_comp_name)
  associate (c => operation_index(rid, sid, mapping_tag))
    ! local 1-based indices + coefS in operation (= table) order.  n_send / n_recv only have to cover the indices
    ! used: interpolate_data accepts any sn1 >= n_send, rn1 >= n_recv and zero-fills every row of recv_data beyond
    ! them, as the reference does (ref :293).  A rank without operations (maxval of an empty array is -huge) gets a
    ! valid operator that only zero-fills.
    n_send = 1; n_recv = 1
    if (size(c%send_data_index) > 0) then
       n_send = max(1, maxval(c%send_data_index)); n_recv = max(1, maxval(c%recv_data_index))
    end if
    call dccm_check( dccm_remap_create(int(size(c%send_data_index), c_int64_t), c%send_data_index, &
         & c%recv_data_index, c%coefS, n_send, n_recv, handle), "dccm_remap_create")
    call dccm_check( dccm_interp_register(rid, sid, mapping_tag, handle), "dccm_interp_register")
  end associate
end subroutine dccm_register_operation
