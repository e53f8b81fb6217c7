"""dennou-ccm_b200 -- B200-native (sm_100a) atmosphere-ocean surface-exchange step of Dennou-CCM.

Only the hot path lives here: the remap apply, the bulk surface flux and
the column implicit coupling solve as hand-written CUDA behind a C ABI (csrc/, include/), and
thin host mirrors of the reference's Fortran module interfaces.  Import with
importlib.import_module("dennou-ccm_b200") (the directory name is not a Python identifier).
"""
from . import _lib
from ._lib import DccmError, build, build_driver, build_gmapgen, lib
from . import tables, grid_mapping_util, grid_mapping_util_jones99, cal_mappingtable
from . import interpolation_data_latlon_mod, dsfcm, dcpam_sfc_implicit_coupling_mod, dccm_ocn_mod, dccm_atm_mod, dcpam_main_mod
from .interpolation_data_latlon_mod import RemapOperator, interpolate_data
from .dcpam_sfc_implicit_coupling_mod import SfcImplicitCoupling
from .dsfcm import DSFCM_Util_SfcBulkFlux_Get

__all__ = ["DccmError", "build", "lib", "tables", "grid_mapping_util", "grid_mapping_util_jones99",
           "interpolation_data_latlon_mod", "dsfcm", "dcpam_sfc_implicit_coupling_mod",
           "RemapOperator", "interpolate_data", "SfcImplicitCoupling", "DSFCM_Util_SfcBulkFlux_Get"]
