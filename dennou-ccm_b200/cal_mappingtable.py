"""Mirror of common/cal_mappingtable.f90 (module mapping_table + the program around it).

The reference program is not built (it refers to undefined COMP1_NX ... names); its
`make_mapping_table` is the stand-alone bilinear generator for regular grids and is what north_star names.
"""
from .tables import make_mapping_table as _make


def make_mapping_table(file, nx_r, ny_r, nx_s, ny_s):
    """ref :10-49 -- same argument list, `file` being the table file name instead of a Fortran unit number;
    writes one list-directed line `i j is js coef` per entry with coef > 0."""
    _make(nx_r, ny_r, nx_s, ny_s).write(file)


def cal_mappingtable(atm, ocn, nx_a, ny_a, nx_o, ny_o, directory="."):
    """ref :81-109: the two files the program writes, ATM -> OCN under the component names and the reverse
    direction as mapping_table_1_to_2.txt (file names as in the reference, :100, :104)."""
    import os
    f1 = os.path.join(directory, f"mapping_table_{atm}_to_{ocn}.txt")
    f2 = os.path.join(directory, "mapping_table_1_to_2.txt")
    make_mapping_table(f1, nx_a, ny_a, nx_o, ny_o)        # :102 (receiver sizes first, as called there)
    make_mapping_table(f2, nx_o, ny_o, nx_a, ny_a)        # :106
    return f1, f2
