"""Host-side HEALPix helpers (numpy) for the shell set-up the reference does on the CPU:
``hp_shell_alloc`` (common.c:505-552) needs ``pix2vec_nest`` for the pixels a rank owns.
Restates the published NEST indexing (Gorski et al. 2005): pixel -> (x, y, face) -> (z, phi).
"""
from __future__ import annotations

import numpy as np

_JRLL = np.array([2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4])
_JPLL = np.array([1, 3, 5, 7, 0, 2, 4, 6, 1, 3, 5, 7])


def _compress_bits(v):
    x = v & 0x55555555
    x = (x | (x >> 1)) & 0x33333333
    x = (x | (x >> 2)) & 0x0f0f0f0f
    x = (x | (x >> 4)) & 0x00ff00ff
    x = (x | (x >> 8)) & 0x0000ffff
    return x


def pix2vec_nest(nside: int, ipix) -> np.ndarray:
    """Unit vectors [n,3] of NEST pixels (chealpix pix2vec_nest)."""
    ipix = np.asarray(ipix, dtype=np.int64)
    npface = nside * nside
    face = ipix // npface
    p = ipix & (npface - 1)
    ix = _compress_bits(p)
    iy = _compress_bits(p >> 1)
    nl4 = 4 * nside
    fact2 = 4.0 / (12 * nside * nside)
    jr = _JRLL[face] * nside - ix - iy - 1
    north, south = jr < nside, jr > 3 * nside
    nr = np.where(north, jr, np.where(south, nl4 - jr, nside))
    fact1 = (nside << 1) * fact2
    z = np.where(north, 1 - nr * nr * fact2, np.where(south, nr * nr * fact2 - 1, (2 * nside - jr) * fact1))
    kshift = np.where(north | south, 0, (jr - nside) & 1)
    jp = (_JPLL[face] * nr + ix - iy + 1 + kshift) // 2
    jp = np.where(jp > nl4, jp - nl4, jp)
    jp = np.where(jp < 1, jp + nl4, jp)
    phi = (jp - (kshift + 1) * 0.5) * ((np.pi / 2) / nr)
    st = np.sqrt((1.0 - z) * (1.0 + z))
    return np.stack([st * np.cos(phi), st * np.sin(phi), z], axis=1)


def hp_shell_pixels(nside: int, nside_base: int = 2, node: int = 0, nnodes: int = 1):
    """hp_shell_alloc (common.c:517-540): NEST ids and unit vectors of the pixels owned by ``node``."""
    nbases = 12 * nside_base * nside_base
    per = (nside // nside_base) ** 2
    bases = np.arange(node, nbases, nnodes, dtype=np.int64)
    listpix = (bases[:, None] * per + np.arange(per, dtype=np.int64)[None, :]).ravel()
    return listpix, pix2vec_nest(nside, listpix)
