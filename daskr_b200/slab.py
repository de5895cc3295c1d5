"""y-slab decomposition of the mesh across ranks, host-side arithmetic.

Same formulas as dkb_setup_rbdpre in csrc/capi.cu: rows are split as evenly as possible, rank p
owns rows [jy0, jy0 + myloc) and exchanges one mesh row (ns*mx doubles) with each neighbour
before every residual evaluation; dot products and norms are all-reduced scalars.
"""


def partition(my, nranks, rank):
    """(jy0, myloc) of `rank`: first owned global row (0-based) and number of owned rows."""
    base, rem = divmod(my, nranks)
    myloc = base + (1 if rank < rem else 0)
    jy0 = rank * base + min(rank, rem)
    return jy0, myloc


def halo_bytes(ns, mx):
    """bytes sent to each neighbour per exchange"""
    return 8 * ns * mx


# Gauss-Seidel factor on several ranks (csrc/gs.cu: web_gauss_seidel / slab_overlap_exchange): rows of the neighbouring
# slabs that a rank sweeps redundantly.  Same constants and the same min() as the CUDA path.
GS_OV_BELOW = 32
GS_OV_ABOVE = 8


def gs_overlap(my, nranks, rank):
    """(rows_below, rows_above) that `rank` receives and sweeps in addition to its own rows."""
    my_min = my // nranks                       # every slab has at least this many rows
    ob, oa = min(GS_OV_BELOW, my_min), min(GS_OV_ABOVE, my_min)
    return (ob if rank > 0 else 0), (oa if rank < nranks - 1 else 0)
