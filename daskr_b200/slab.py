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
