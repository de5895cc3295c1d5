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


def set_grouping(m, ng):
    """set_grouping, src/daskr_rbgpre.f90:160-193: (jig, jr) = group of every mesh index (1-based values, 0-based list)
    and representative index of every group (1-based).  Same arithmetic as dkb_setup_rbgpre in csrc/capi.cu."""
    mper = m // ng
    ngm1 = ng - 1
    jig = [1 + j // mper for j in range(ngm1 * mper)] + [ng] * (m - ngm1 * mper)
    jr = [int(0.5 + (ig - 0.5) * mper) for ig in range(1, ng)] + [int(0.5 * (1 + ngm1 * mper + m))]
    return jig, jr


def rbg_representatives(mx, my, nxg, nyg, nranks, rank):
    """Local point index (row-major in the rank's slab) of each group's representative point, -1 if another rank owns
    it: the d_rep table of dkb_setup_rbgpre.  Every rank computes the blocks of the groups it owns; the unfactored
    blocks are summed over ranks (zeros elsewhere) before cj*I_d is added and they are factored."""
    _, jxr = set_grouping(mx, nxg)
    _, jyr = set_grouping(my, nyg)
    jy0, myloc = partition(my, nranks, rank)
    rep = []
    for igy in range(nyg):
        jyl = jyr[igy] - 1 - jy0
        for igx in range(nxg):
            rep.append(jyl * mx + (jxr[igx] - 1) if 0 <= jyl < myloc else -1)
    return rep
