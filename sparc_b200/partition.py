"""Band-parallel column split across GPUs: SPARC's ``npband`` axis.

Mirrors src/parallelization.c:403-428 -- ``NB = ceil(Nstates / npband)``; band-communicator
``r`` owns columns ``[r*NB, min((r+1)*NB, Nstates))`` (trailing ranks may own none).  The filter
has no communication along this axis, so each GPU simply filters its slice.
"""
from __future__ import annotations


def band_partition(nstates: int, npband: int, rank: int):
    """Return (first_column, ncol) of ``rank``."""
    nb = -(-int(nstates) // int(npband))
    start = min(rank * nb, nstates)
    end = min((rank + 1) * nb, nstates)
    return start, max(0, end - start)
