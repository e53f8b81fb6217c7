"""Mirror of module grid_mapping_util (bilinear; ref common/grid_mapping_util.f90)."""
import numpy as np

from .tables import Grid, MappingTable, gen_table_bilinear


def gen_gridmapfile_lonlat2lonlat(filename, x_LonS, y_LatS, x_LonR, y_LatR, lon_mode=0):
    """ref :32-48 -- same argument list; writes the text table file."""
    f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    z = lambda a: np.zeros(len(a))
    src = Grid(len(x_LonS), len(y_LatS), f(x_LonS), f(y_LatS), z(x_LonS), z(y_LatS))
    dst = Grid(len(x_LonR), len(y_LatR), f(x_LonR), f(y_LatR), z(x_LonR), z(y_LatR))
    gen_table_bilinear(src, dst, lon_mode).write(filename)


def set_mappingTable_interpCoef(gridmapfile, GNXS, GNXR):
    """ref :181-241 (identical to the jones99 module's reader)."""
    return MappingTable.read(gridmapfile).index(GNXS, GNXR)
