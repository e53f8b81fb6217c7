"""Mirror of module grid_mapping_util_jones99 (ref common/grid_mapping_util_jones99.f90)."""
import numpy as np

from .tables import Grid, MappingTable, gen_table_jones99


def gen_gridmapfile_lonlat2lonlat(filename, x_LonS, y_LatS, x_LonD, y_LatD,
                                  x_LonIntWtS, y_LatIntWtS, x_LonIntWtD, y_LatIntWtD,
                                  accuracy_order, lon_mode=0):
    """ref :35-54 -- same argument list; writes the text table file."""
    f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    src = Grid(len(x_LonS), len(y_LatS), f(x_LonS), f(y_LatS), f(x_LonIntWtS), f(y_LatIntWtS))
    dst = Grid(len(x_LonD), len(y_LatD), f(x_LonD), f(y_LatD), f(x_LonIntWtD), f(y_LatIntWtD))
    gen_table_jones99(src, dst, accuracy_order, lon_mode).write(filename)


def set_mappingTable_interpCoef(gridmapfile, GNXS, GNXR):
    """ref :446-506 -- returns (send_index, recv_index, coef_s), 1-based indices."""
    return MappingTable.read(gridmapfile).index(GNXS, GNXR)
