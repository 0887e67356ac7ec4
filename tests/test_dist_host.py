"""CPU suite, world_size 2 over gloo: the host-side logic of the slab decomposition -- slab bounds,
the all-to-all block bookkeeping of the distributed FFT (y-slab k space, z-slab real space, staging
buffer read with a two-level stride) and the reductions that replace MPI_Allreduce."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from colore_b200.dist import c2r_dist_numpy, fill_tile_modes, kspace_slab, r2c_dist_numpy, slab_bounds


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _alltoall(blocks):
    """all-to-all of numpy blocks over the default (gloo) group via point-to-point messages."""
    rank, world = dist.get_rank(), dist.get_world_size()
    out = [None] * world
    out[rank] = blocks[rank]
    reqs, bufs = [], {}
    for k in range(1, world):
        to, frm = (rank + k) % world, (rank - k) % world
        ts = torch.view_as_real(torch.from_numpy(np.ascontiguousarray(blocks[to]))).contiguous()
        tr = torch.empty_like(ts)
        bufs[frm] = tr
        reqs.append(dist.isend(ts, to))
        reqs.append(dist.irecv(tr, frm))
    for r in reqs:
        r.wait()
    for frm, tr in bufs.items():
        out[frm] = torch.view_as_complex(tr).numpy()
    return out


def _worker(rank, world, port, n, errs):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(123)                      # same full array on every rank
        nc = n // 2 + 1
        ck = rng.standard_normal((n, n, nc)) + 1j * rng.standard_normal((n, n, nc))
        nzl, iz0 = slab_bounds(n, world, rank)
        # c2r: distributed data movement == single-process transform restricted to the slab
        got = c2r_dist_numpy(kspace_slab(ck, rank, world), rank, world, _alltoall)
        ref = np.fft.irfftn(ck, s=(n, n, n), axes=(0, 1, 2)) * n ** 3
        e1 = np.abs(got - ref[iz0:iz0 + nzl]).max() / np.abs(ref).max()
        # the same exchange through the tile-major staging layout of the fused transpose (ragged last tile: nc odd)
        got_t = c2r_dist_numpy(kspace_slab(ck, rank, world), rank, world, _alltoall, tile=8)
        e1 = max(e1, np.abs(got_t - got).max() / np.abs(ref).max())
        # r2c: real z slab -> y slab of the spectrum
        x = rng.standard_normal((n, n, n))
        gotk = r2c_dist_numpy(x[iz0:iz0 + nzl], rank, world, _alltoall)
        refk = np.fft.rfftn(x, axes=(0, 1, 2))
        nyl = n // world
        e2 = np.abs(gotk - refk[:, rank * nyl:(rank + 1) * nyl, :]).max() / np.abs(refk).max()
        # the reductions that replace MPI_Allreduce (fourier.c:69-70, density.c:1262-1269)
        mom = torch.tensor([ref[iz0:iz0 + nzl].sum(), (ref[iz0:iz0 + nzl] ** 2).sum()], dtype=torch.float64)
        dist.all_reduce(mom)
        e3 = abs(mom[1].item() / (ref ** 2).sum() - 1)
        errs[rank] = max(e1, e2, e3)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [16, 32])
def test_distributed_fft_bookkeeping_gloo(n):
    world = 2
    with mp.Manager() as m:
        errs = m.dict()
        mp.spawn(_worker, args=(world, _free_port(), n, errs), nprocs=world, join=True)
        assert len(errs) == world
        assert max(errs.values()) < 1e-12


def test_tiled_stage_index_is_a_bijection():
    """tile-major staging layout (clr_fft.cu store_peer / prefetch): every (line, plane) gets its own slot, the T
    lines of a tile at consecutive planes are contiguous, and a whole (tile, destination) block is one run."""
    from colore_b200.dist import tiled_stage_index
    nzl, nyl, nc, tile = 4, 8, 17, 8
    n_inner = nyl * nc
    zz, ii = np.meshgrid(np.arange(nzl), np.arange(n_inner), indexing="ij")
    pos = tiled_stage_index(ii, zz, nzl, tile)
    assert len(np.unique(pos)) == pos.size and pos.max() < nzl * tile * ((n_inner + tile - 1) // tile)
    assert np.all(np.diff(pos[0, :tile]) == 1)                       # lines of one tile are adjacent
    assert tiled_stage_index(0, 1, nzl, tile) - tiled_stage_index(0, 0, nzl, tile) == tile   # next plane follows
    assert tiled_stage_index(tile, 0, nzl, tile) == nzl * tile       # next tile starts after all planes


@pytest.mark.parametrize("interp", [0, 1, 2])
def test_lpt_routing_covers_every_plane_once(interp):
    """Particle exchange bookkeeping of the multi-GPU LPT: with every rank depositing its own particles plus the
    ones routed to it (each only into planes it owns), every (particle, plane) pair is deposited exactly once."""
    from colore_b200.dist import lpt_destinations_numpy, lpt_planes_numpy
    n, l_box, P = 32, 100.0, 4
    rng = np.random.default_rng(5)
    nzl = n // P
    total = np.zeros(n)
    ref = np.zeros(n)
    for me in range(P):
        # particles born in slab `me`, displaced by up to ~1.5 slabs (periodic wrap like density.c:875-876)
        z = (rng.uniform(me * nzl, (me + 1) * nzl, 500) + rng.normal(0, 0.6 * nzl, 500)) * l_box / n
        z = np.mod(z, l_box).astype(np.float32)
        z[z >= np.float32(l_box)] = 0
        planes = lpt_planes_numpy(z, interp, n, l_box)
        np.add.at(ref, planes.ravel(), 1)
        need = lpt_destinations_numpy(z, interp, n, l_box, P, me)
        for h in range(P):
            sel = np.ones(len(z), bool) if h == me else need[:, h]      # own particles stay, routed ones arrive
            pl = planes[sel]
            mine = (pl // nzl) == h                                     # the deposit's slab check
            np.add.at(total, pl[mine], 1)
    assert np.array_equal(total, ref)


@pytest.mark.parametrize("n,nranks,t_lines", [(64, 2, 32), (64, 4, 32), (128, 2, 32), (1024, 8, 16), (2048, 8, 8)])
def test_fused_fill_tiles_cover_every_mode_once(n, nranks, t_lines):
    """fill_peer_kernel (several GPUs, mode fill inside the transpose pass): over all ranks and tiles every mode
    (kz, ky, kx) of the half-spectrum is generated exactly once, a pair of lines never straddles a row (so the two modes
    of a Philox block are the neighbours kx = 2p, 2p + 1 of one row, as in the stand-alone fill and the oracle), and the
    block counters of all pairs are distinct."""
    nc = n // 2 + 1
    ncp = (nc + 7) // 8 * 8
    nyl = n // nranks
    seen = np.zeros((n, nc), np.int32)                       # (ky, kx) lines; every line carries all n kz
    blocks = []
    for rank in range(nranks):
        n_tiles = (nyl * ncp + t_lines - 1) // t_lines
        tiles = range(n_tiles) if n <= 128 else list(range(3)) + [n_tiles // 2, n_tiles - 1]
        for tile in tiles:
            ky, kx, live, block = fill_tile_modes(tile, t_lines, n, nranks, rank)
            assert np.all(ky[0::2] == ky[1::2]) and np.all(kx[0::2] % 2 == 0) and np.all(kx[1::2] == kx[0::2] + 1)
            np.add.at(seen, (ky[live], kx[live]), 1)
            keep = live[0::2]                                # pairs whose even line is a real mode
            blocks.append(block[keep][:, [0, 1, n - 1]].ravel())
            # the oracle's counter of the pair, evaluated independently for kz = 1
            want = kx[0::2][keep] // 2 + ((nc + 1) // 2) * (ky[0::2][keep] + n * 1)
            assert np.array_equal(block[keep][:, 1], want)
    if n <= 128:
        assert np.all(seen == 1)                             # every (ky, kx) line exactly once over ranks and tiles
    else:
        assert seen.max() == 1
    allb = np.concatenate(blocks)
    assert len(np.unique(allb)) == len(allb)


def test_slab_bounds():
    assert slab_bounds(1024, 8, 3) == (128, 384)
    assert slab_bounds(64, 1, 0) == (64, 0)
    with pytest.raises(ValueError):
        slab_bounds(100, 8, 0)
    # slabs tile the grid exactly
    for p in (1, 2, 4, 8):
        b = [slab_bounds(256, p, r) for r in range(p)]
        assert b[0][1] == 0 and all(b[i][1] + b[i][0] == b[i + 1][1] for i in range(p - 1))
        assert b[-1][1] + b[-1][0] == 256
