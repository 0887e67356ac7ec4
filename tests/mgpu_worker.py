"""Worker of tests/test_gpu_multi.py: run under torchrun with >= 2 GPUs (one process per GPU).

Checks the slab-decomposed path against the CPU oracle on the same seeded inputs:
Gaussian fields (distributed FFT through the NCCL all-to-all, halo exchange, all-reduced sigma^2),
lognormal transform, all-reduced normalisation, per-slab Poisson counts (bit-exact) and a
distributed r2c/c2r round trip.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import colore_b200 as cb  # noqa: E402
from oracle.oracle import RNG_PHILOX, Oracle, tables_from_dump  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = int(os.environ.get("CLR_TEST_N", "64"))
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", "ref_n32_lognormal.npz")))
    t = tables_from_dump(g)
    t["l_box"] = float(np.float32(2 * t["r_max"] * (1 + 2. / n)))
    t["pos_obs"] = 0.5 * t["l_box"]
    seed = 99
    nzl, iz0 = cb.dist.slab_bounds(n, world, rank)
    par = cb.ParamCoLoRe(t, n, seed=seed, nz_here=nzl, iz0_here=iz0, device=local)
    cb.dist.init_comm(par, rank, world)
    # CLR_TEST_EXACT=0: the fp32 field kernels, i.e. the mode fill fused into the peer-store z pass (fill_peer_kernel)
    par.set_option("exact_math", int(os.environ.get("CLR_TEST_EXACT", "1")))

    # oracle, full box (every rank computes it; small n)
    o = Oracle(t, n)
    dk, pk = o.fill_modes(RNG_PHILOX, seed)
    d0, p0 = o.c2r(dk), o.c2r(pk)
    o.normalize_fields(d0, p0)
    _, s2_ref = o.sigma_dens(d0)
    d_gauss = d0.copy()

    mean, s2 = cb.create_cartesian_fields(par)
    dens = par.grid_get(cb.GRID_DENS)
    npot = par.grid_get(cb.GRID_NPOT)
    sl = slice(iz0, iz0 + nzl)
    e_d = np.abs(dens[:, :, :n] - d0[sl, :, :n]).max() / np.sqrt(s2_ref)
    e_p = np.abs(npot[:, :, :n] - p0[sl, :, :n]).max() / p0[:, :, :n].std()
    assert e_d < 2e-5 and e_p < 2e-5, (rank, e_d, e_p)
    assert abs(s2 / s2_ref - 1) < 1e-5, (s2, s2_ref)

    # distributed r2c -> c2r round trip on the potential
    cb.fftw_wrap_r2c(par, cb.GRID_NPOT)
    cb.fftw_wrap_c2r(par, cb.GRID_NPOT)
    back = par.grid_get(cb.GRID_NPOT)
    e_rt = np.abs(back[:, :, :n] / float(n) ** 3 - npot[:, :, :n]).max() / npot[:, :, :n].std()
    assert e_rt < 3e-5, (rank, e_rt)
    par.grid_put(cb.GRID_NPOT, npot)
    par.update_halo()

    # lognormal + normalisation (histograms all-reduced) + sources on the slab
    cb.compute_physical_density_field(par)
    par.set_srcs(0, t["srcs_nz_0"], t["srcs_bz_0"])
    cb.compute_density_normalization(par)
    norm, ends, _ = cb.get_norm(par, 0, 0)
    o.lognormalize(d0, s2_ref)
    nm = o.density_normalization(d0, [t["srcs_bz_0"]])
    np.testing.assert_allclose(norm, nm["norm"][0], rtol=2e-5)     # fields differ at the fp32 level
    ln = par.grid_get(cb.GRID_DENS)
    nsrc = cb.srcs_set_cartesian(par)[0]
    counts = cb.srcs_get_counts(par, 0)
    # oracle on this rank's slab with the GPU's own field / normalisation -> counts must be bit-exact
    os_ = Oracle(t, n, nz_here=nzl, iz0_here=iz0)
    plane = npot.shape[1] * npot.shape[2]
    left = p0[(iz0 - 1) % n] if False else None
    ns, tot = os_.srcs_poisson(ln, t["srcs_nz_0"], t["srcs_bz_0"], norm, ends[0], ends[1], RNG_PHILOX, seed, 0)
    assert tot == nsrc and np.array_equal(ns, counts), (rank, tot, nsrc)
    # halo planes: positions/RSD use the neighbours' potential planes -> compare placement with the oracle
    halo_l = np.empty((n, npot.shape[2]), np.float32)
    halo_r = np.empty_like(halo_l)
    allp = [torch.empty(npot.shape, dtype=torch.float32, device="cuda") for _ in range(world)]
    dist.all_gather(allp, torch.from_numpy(npot).cuda())
    full = torch.cat(allp, 0).cpu().numpy()
    halo_l[:], halo_r[:] = full[(iz0 - 1) % n], full[(iz0 + nzl) % n]
    os_.set_halo(npot, halo_l, halo_r)
    pos_ref, ipix_ref = os_.srcs_place(npot, ns, RNG_PHILOX, seed, 0)
    pos, ipix = cb.srcs_get_cartesian(par, 0)
    assert np.array_equal(ipix, ipix_ref) and np.array_equal(pos[:, :3], pos_ref[:, :3])
    np.testing.assert_allclose(pos[:, 3], pos_ref[:, 3], rtol=2e-6, atol=1e-12)
    # RSD under beaming (srcs.c:425-443, 486-504, 656-662): CIC corners of sources in the first / last plane of a slab
    # reach into the neighbour slabs (two halo rings here, the slab rotation in the reference) -> full-box oracle
    before = cb.srcs_get_local_properties(par, 0)
    cb.srcs_beams(par)
    after = cb.srcs_get_local_properties(par, 0)
    o.set_halo(full)
    rsd_ref = o.srcs_beam_rsd(full, pos, before.copy())
    np.testing.assert_allclose(after[:, 3], rsd_ref[:, 3], rtol=2e-5, atol=1e-9)
    edge = np.abs((pos[:, 2] + 0.5 * par.l_box) * (n / par.l_box) - iz0) < 0.5      # sources of the slab's first plane
    assert edge.sum() > 0 and np.abs(after[edge, 3]).max() > 0
    # srcs_distribute (srcs.c:296-373): rank ipix % NNodes gets the source; blocks arrive from rank-1, rank-2, ..., own
    # sources last, order inside a block = the sender's catalogue order. Checked against that rule applied to the
    # gathered per-slab catalogues (bit for bit). (The maps below only read the potential.)
    allc = [None] * world
    dist.all_gather_object(allc, (pos.copy(), ipix.copy()))
    want_p, want_i = [], []
    for ii in range(world):
        frm = (rank - 1 - ii) % world
        sel = allc[frm][1] % world == rank
        want_p.append(allc[frm][0][sel]); want_i.append(allc[frm][1][sel])
    want_p, want_i = np.concatenate(want_p), np.concatenate(want_i)
    n_before = par.nsources[0]
    cb.srcs_distribute(par, by_pixel=True)
    pos_d, ipix_d = cb.srcs_get_cartesian(par, 0)
    assert np.array_equal(ipix_d, want_i) and np.array_equal(pos_d, want_p), (rank, len(ipix_d), len(want_i))
    assert np.all(ipix_d % world == rank)
    tot_d = torch.tensor([len(ipix_d), n_before], device="cuda")
    dist.all_reduce(tot_d)
    assert tot_d[0].item() == tot_d[1].item()
    srcs_d = cb.srcs_get_local_properties(par, 0)
    np.testing.assert_allclose(srcs_d[:, :3], os_.srcs_local_properties(pos_d)[:, :3], rtol=3e-7, atol=3e-5)
    assert np.array_equal(srcs_d[:, 3], pos_d[:, 3])
    # maps: every GPU integrates the ray segments inside its slab, the partial maps are all-reduced
    _, pix = cb.healpix.hp_shell_pixels(8, 2)
    rf = np.sort(g["s6_kappa_rf"])
    o.set_halo(full)
    kap = cb.kappa_get_beam_properties(par, pix, rf)
    kap_ref = o.kappa(full, pix, rf)
    np.testing.assert_allclose(kap, kap_ref, rtol=2e-5, atol=2e-6 * np.abs(kap_ref).max())
    isw = cb.isw_get_beam_properties(par, pix, rf)
    isw_ref = o.isw(full, pix, rf)
    np.testing.assert_allclose(isw, isw_ref, rtol=2e-5, atol=2e-6 * np.abs(isw_ref).max())
    # LPT densities on slabs: distributed r2c/c2r + particle exchange between slabs (density.c:191-374)
    # from the oracle's Gaussian field; the deposits must agree with the single-box oracle
    lpt_sent = 0
    for order, interp in ((1, 1), (2, 2), (2, 0)):
        ref = d_gauss.copy()
        o.lpt(ref, order, interp)
        pl = cb.ParamCoLoRe(t, n, dens_type=order, seed=seed, nz_here=nzl, iz0_here=iz0, device=local)
        cb.dist.init_comm(pl, rank, world)
        pl.set_option("lpt_interp_type", interp)
        pl.grid_put(cb.GRID_DENS, np.ascontiguousarray(d_gauss[sl]))
        cb.compute_physical_density_field(pl)
        got = pl.grid_get(cb.GRID_DENS)[:, :, :n].astype(np.float64)
        want = ref[sl, :, :n].astype(np.float64)
        sent, recv = cb.lpt_exchange_counts(pl)
        lpt_sent += sent
        tot = torch.tensor([got.sum(), float(sent), float(recv)], dtype=torch.float64, device="cuda")
        dist.all_reduce(tot)
        assert abs(tot[0].item()) < 0.5, tot                      # mass conservation over all slabs
        assert tot[1].item() == tot[2].item() and (interp == 0 or tot[1].item() > 0), tot   # NGP: coarse cells may keep every particle home
        if interp == 0:
            assert np.mean(got != want) < 1e-3, (rank, order, np.mean(got != want))
        else:
            assert np.abs(got - want).max() < 3e-4, (rank, order, interp, np.abs(got - want).max())
        pl.free()
    tot_all = torch.tensor([nsrc], device="cuda")
    dist.all_reduce(tot_all)
    if rank == 0:
        print(f"MGPU OK world={world} n={n} field_err={e_d:.2e} roundtrip={e_rt:.2e} nsrc_total={int(tot_all.item())} "
              f"lpt_particles_shipped_rank0={lpt_sent}")
    par.free()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
