"""Host-side logic of the slab decomposition (one process per GPU).

* ``slab_bounds``: the z-slab of a rank (init_fftw, fourier.c:172-181, for n divisible by P).
* ``init_comm``: creates the NCCL communicator of a ``ParamCoLoRe`` from a unique id that rank 0
  generates and ``torch.distributed`` broadcasts (replaces ``mpi_init``, common.c:216-274).
* ``c2r_dist_numpy`` / ``r2c_dist_numpy``: a numpy restatement of the DATA MOVEMENT of the distributed
  transform in colore_b200/csrc/clr_fft.cu (k space in y slabs ``[kz][ky_local][kx]``, one all-to-all
  whose block for rank h is the contiguous range of z planes of h, y pass reading the staging buffer
  ``[source][z_local][ky_in_source][kx]``). It exists so that the exchange bookkeeping can be tested
  on CPU with gloo (tests/test_dist_host.py); it is not a compute path.
"""
from __future__ import annotations

import ctypes as C

import numpy as np


def slab_bounds(n: int, nranks: int, rank: int):
    """(nz_here, iz0_here) of ``rank``; n must be divisible by nranks."""
    if n % nranks:
        raise ValueError(f"n_grid={n} is not divisible by {nranks} GPUs")
    nz = n // nranks
    return nz, rank * nz


def init_comm(par, rank: int, nranks: int):
    """NCCL communicator for ``par`` (a pipeline.ParamCoLoRe created with this rank's slab bounds)."""
    from ._lib import check
    if nranks == 1:
        check(par.lib.clr_comm_init(par.ctx, C.c_int(0), C.c_int(1), None))
        return
    import torch.distributed as dist
    ident = (C.c_ubyte * 128)()
    if rank == 0:
        check(par.lib.clr_comm_unique_id(ident))
    box = [bytes(ident)]
    dist.broadcast_object_list(box, src=0)
    buf = (C.c_ubyte * 128).from_buffer_copy(box[0])
    import os
    if os.environ.get("COLORE_B200_OVERLAP", "0") == "1":  # second transform pipeline for the potential (experiment, default off)
        par.set_option("fft_overlap", 1)
    check(par.lib.clr_comm_init(par.ctx, C.c_int(rank), C.c_int(nranks), buf))
    if os.environ.get("COLORE_B200_P2P", "1") == "0":      # force the NCCL all-to-all (comparison runs)
        par.set_option("p2p_fused", 0)
    if os.environ.get("COLORE_B200_P2P_TILED"):            # 0 / 1: force the staging layout (comparison runs)
        par.set_option("p2p_tiled", int(os.environ["COLORE_B200_P2P_TILED"]))


def transpose_mode(par) -> str:
    """How the slab transpose of the distributed FFT runs: peer-memory stores fused into the producing
    pass ("p2p-fused") or an NCCL all-to-all ("nccl")."""
    return "p2p-fused" if par.lib.clr_comm_p2p(par.ctx) == 1 else "nccl"


# ---- numpy restatement of the exchange (test support) ---------------------------------------------

def kspace_slab(ck_full: np.ndarray, rank: int, nranks: int) -> np.ndarray:
    """[kz][ky][kx] (reference layout) -> this rank's y slab [kz][ky_local][kx]."""
    n = ck_full.shape[0]
    nyl = n // nranks
    return np.ascontiguousarray(ck_full[:, rank * nyl:(rank + 1) * nyl, :])


def tiled_stage_index(inner, z_local, nzl: int, tile: int):
    """Position of z-pass line ``inner`` (= ky_local*nc + kx) at plane ``z_local`` inside one source block of the
    TILE-MAJOR staging layout of the fused c2r transpose (clr_fft.cu: store_peer / prefetch, ``tiled``):
    [z-pass tile][z_local][T], so that the T lines of a tile at consecutive planes are contiguous."""
    return ((inner // tile) * nzl + z_local) * tile + inner % tile


def c2r_dist_numpy(kslab: np.ndarray, rank: int, nranks: int, alltoall, tile: int = 0) -> np.ndarray:
    """Distributed unnormalised c2r. ``alltoall(list_of_P_blocks) -> list_of_P_blocks``.
    ``tile`` > 0: the blocks travel in the tile-major staging layout (``tiled_stage_index``)."""
    n, nyl, nc = kslab.shape
    nzl = n // nranks
    a = np.fft.ifft(kslab, axis=0) * n                       # z pass on [kz][ky_local][kx]
    send = [np.ascontiguousarray(a[h * nzl:(h + 1) * nzl]) for h in range(nranks)]   # contiguous z ranges
    if tile:
        n_inner = nyl * nc
        blk = nzl * tile * ((n_inner + tile - 1) // tile)
        zz, ii = np.meshgrid(np.arange(nzl), np.arange(n_inner), indexing="ij")
        pos = tiled_stage_index(ii, zz, nzl, tile)
        packed = []
        for b in send:                                        # what the z pass stores on the destination
            t = np.zeros(blk, b.dtype)
            t[pos.ravel()] = b.reshape(nzl, n_inner).ravel()
            packed.append(t)
        got = alltoall(packed)
        # what the y pass gathers: element (z_local, ky = src*nyl + ky_l, kx) from block src
        recv = [g[pos.ravel()].reshape(nzl, nyl, nc) for g in got]
    else:
        recv = alltoall(send)                                 # staging [source][z_local][ky_in_source][kx]
    stage = np.stack(recv, axis=0)
    # y pass: element e of the line = stage[e // nyl, z_local, e % nyl, kx]  (two-level stride)
    lines = stage.transpose(1, 0, 2, 3).reshape(nzl, n, nc)
    b = np.fft.ifft(lines, axis=1) * n
    # x pass: half-complex -> real, imaginary parts of DC / Nyquist dropped (numpy irfft does the same)
    return np.fft.irfft(b, n=n, axis=2) * n


def r2c_dist_numpy(rslab: np.ndarray, rank: int, nranks: int, alltoall) -> np.ndarray:
    """Distributed r2c: real z slab [z_local][y][x] -> y slab of the spectrum [kz][ky_local][kx]."""
    nzl, n, _ = rslab.shape
    nyl = n // nranks
    a = np.fft.rfft(rslab, axis=2)
    b = np.fft.fft(a, axis=1)                                 # [z_local][ky][kx]
    send = [np.ascontiguousarray(b[:, s * nyl:(s + 1) * nyl, :]) for s in range(nranks)]
    recv = alltoall(send)                                     # block h = z planes of rank h, my ky
    full_z = np.concatenate(recv, axis=0)                     # [kz][ky_local][kx]
    return np.fft.fft(full_z, axis=0)


def fill_tile_modes(tile: int, t_lines: int, n: int, nranks: int, rank: int):
    """Modes a CTA of the fused fill + transpose pass (clr_fft.cu: fill_peer_kernel) generates for z-pass tile ``tile``:
    lines = ``t_lines`` consecutive values of the flattened index ky_local * ncp + kx (ncp = row pitch: n/2+1 rounded up
    to a multiple of 8), a thread owns a PAIR of lines (same ky, kx even / odd) and draws ONE Philox block per pair and
    kz. Returns (ky_global[lines], kx[lines], live[lines], block_index[pairs, n]) with block_index the counter the
    oracle uses for the pair: kx // 2 + ceil((n/2+1)/2) * (ky + n * kz). Padding columns
    (kx >= n/2+1) and lines past the end of the slab are not live (the kernel writes zeros there)."""
    nc = n // 2 + 1
    ncp = (nc + 7) // 8 * 8
    nyl = n // nranks
    inner = tile * t_lines + np.arange(t_lines)
    kyl, kx = inner // ncp, inner % ncp
    ky = rank * nyl + kyl
    live = (inner < nyl * ncp) & (kx < nc)
    npair_row = (nc + 1) // 2
    pair = np.arange(0, t_lines, 2)
    kz = np.arange(n)
    block = (kx[pair] // 2)[:, None] + npair_row * (ky[pair][:, None] + n * kz[None, :])
    return ky, kx, live, block


# ---- numpy restatement of the LPT particle routing (test support) ----------------------------------

def lpt_planes_numpy(z: np.ndarray, interp: int, n: int, l_box: float) -> np.ndarray:
    """Global z planes touched by the NGP / CIC / TSC deposit of particles at height ``z`` (float32), as
    clr_lpt.cu:lpt_planes computes them (density.c:37-188): array [len(z), 1|2|3], wrapped into [0, n)."""
    z = np.asarray(z, np.float32)
    i_agrid = np.float32(n) / np.float32(l_box)
    if interp == 0:
        i0 = (z * i_agrid).astype(np.float64) + 0.5
        pl = np.floor(i0).astype(np.int64)[:, None]
    elif interp == 1:
        i0 = np.floor(z * i_agrid).astype(np.int64)
        pl = np.stack([i0, i0 + 1], axis=1)
    else:
        c0 = np.floor(((z * i_agrid).astype(np.float64) + 0.5).astype(np.float32)).astype(np.int64)
        pl = np.stack([c0 - 1, c0, c0 + 1], axis=1)
    return np.mod(pl, n)


def lpt_destinations_numpy(z: np.ndarray, interp: int, n: int, l_box: float, nranks: int, me: int) -> np.ndarray:
    """Boolean [len(z), nranks]: rank h needs the particle (share_particles, density.c:191-374, restated as in
    clr_lpt.cu:lpt_route_kernel): h owns one of the touched planes and h != me (own particles are deposited in place)."""
    nzl = n // nranks
    owner = lpt_planes_numpy(z, interp, n, l_box) // nzl
    need = np.zeros((len(z), nranks), bool)
    for k in range(owner.shape[1]):
        need[np.arange(len(z)), owner[:, k]] = True
    need[:, me] = False
    return need
