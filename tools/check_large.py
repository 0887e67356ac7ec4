#!/usr/bin/env python
"""Full-size run of the lognormal path with size-independent property checks (no oracle at this size):

  * r2c(c2r) round trip on the real Gaussian field, compared ON THE DEVICE: |back/N^3 - field| / sigma
  * <delta_G> ~ 0, sigma^2 from the fused moments > 0 and equal on every rank
  * lognormal: min(delta) >= -1, <delta> ~ 0 ("Total density", density.c:1101)
  * sources: sum of the per-cell counts == catalogue length; all base pixels in [0, 12 nside_base^2)

Runs on 1 GPU or under torchrun (one z slab per rank). Prints one JSON line with the stage times.

    python tools/check_large.py --n-grid 2048
    python -m torch.distributed.run --nproc-per-node 8 ... tools/check_large.py --n-grid 2048 --kappa-nside 1024
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class _DevArray:
    """Exposes a raw device pointer to torch through __cuda_array_interface__ (no copy)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 3}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-grid", type=int, default=2048)
    ap.add_argument("--dens-type", type=int, default=0)
    ap.add_argument("--kappa-nside", type=int, default=0)
    ap.add_argument("--imap-nside", type=int, default=0)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--skip-roundtrip", action="store_true")
    ap.add_argument("--opt", action="append", default=[], metavar="NAME=VALUE", help="clr_set_option before the run")
    args = ap.parse_args()
    import ctypes as C

    import torch

    import colore_b200 as cb
    from bench import build_tables, make_config

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def allsum(x):
        if dist is None:
            return float(x)
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        return float(t.item())

    def allmax(x):
        if dist is None:
            return float(x)
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    n = args.n_grid
    cfg = make_config(n)
    cfg.dens_type = args.dens_type
    t = build_tables(cfg)
    nzl, iz0 = cb.dist.slab_bounds(n, world, rank)
    par = cb.ParamCoLoRe(t, n, dens_type=args.dens_type, seed=cfg.seed, device=local, nz_here=nzl, iz0_here=iz0)
    cb.dist.init_comm(par, rank, world)
    par.set_srcs(0, t["srcs_nz_0"], t["srcs_bz_0"])
    for o in args.opt:
        par.set_option(o.split("=")[0], int(o.split("=")[1]))
    out = {"n_grid": n, "n_gpus": world, "dens_type": args.dens_type,
           "transpose": cb.dist.transpose_mode(par) if world > 1 else "none"}
    pitch = par.grid_pitch()
    nfl = nzl * n * pitch

    def dev_view(which):
        return torch.as_tensor(_DevArray(par.grid_device_ptr(which), nfl), device=f"cuda:{local}").view(nzl, n, pitch)

    # ---- property checks -----------------------------------------------------------------
    mean, s2 = cb.create_cartesian_fields(par)
    out["mean_gauss"], out["sigma2_gauss"] = mean, s2
    assert s2 > 0 and abs(mean) < 1e-3 * np.sqrt(s2), (mean, s2)
    if not args.skip_roundtrip:
        g = dev_view(cb.GRID_NPOT)
        ref = g[:, :, :n].clone()
        torch.cuda.synchronize()               # the library runs on its own non-blocking stream
        cb.fftw_wrap_r2c(par, cb.GRID_NPOT)
        cb.fftw_wrap_c2r(par, cb.GRID_NPOT)
        par.synchronize()
        # plane chunks keep the temporaries small (one field is 34 GB at 2048^3)
        s1 = s2r = 0.0
        err = 0.0
        inv = 1.0 / float(n) ** 3
        for z0 in range(0, nzl, 32):
            a = ref[z0:z0 + 32].double()
            s1 += float(a.sum().item()); s2r += float((a * a).sum().item())
            err = max(err, float((g[z0:z0 + 32, :, :n].double() * inv - a).abs().max().item()))
        cnt = allsum(float(nzl) * n * n)
        s1, s2r = allsum(s1), allsum(s2r)
        sig = np.sqrt(s2r / cnt - (s1 / cnt) ** 2)
        out["fft_roundtrip_err_over_sigma"] = allmax(err) / sig
        assert out["fft_roundtrip_err_over_sigma"] < 5e-5, out
        g[:, :, :n].copy_(ref)                     # restore the potential (and its halo)
        del ref
        torch.cuda.synchronize()
        par.update_halo()
        torch.cuda.empty_cache()
    cb.compute_physical_density_field(par)
    par.synchronize()
    d = dev_view(cb.GRID_DENS)
    dmin, dsum = 1e30, 0.0
    for z0 in range(0, nzl, 32):                   # plane chunks: no full-size temporaries
        a = d[z0:z0 + 32, :, :n].double()
        dmin = min(dmin, float(a.min().item()))
        dsum += float(a.sum().item())
    del a, d
    torch.cuda.empty_cache()
    dsum = allsum(dsum)
    out["dens_min"], out["dens_mean"] = dmin, dsum / float(n) ** 3
    assert dmin >= -1.0 and abs(out["dens_mean"]) < 5e-3, out
    cb.compute_density_normalization(par)
    nsrc = cb.srcs_set_cartesian(par)[0]
    out["nsrc_total"] = int(allsum(nsrc))
    assert nsrc > 0

    # ---- timing --------------------------------------------------------------------------
    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def step(seed):
        par.seed = seed
        cb.create_cartesian_fields(par)
        cb.compute_physical_density_field(par)
        cb.compute_density_normalization(par)
        return cb.srcs_set_cartesian(par)[0]

    step(7)
    par.synchronize()
    barrier()
    par.set_profiling(True)
    par.timer_start()
    for s in range(args.steps):
        step(100 + s)
    ms = allmax(par.timer_stop_ms()) / args.steps
    stages = {}
    for nm in ("fill_modes", "fill_fft_z", "fft_z", "fft_a2a", "fft_y", "fft_x", "fft_yx", "halo", "lognormal", "lpt_kspace", "lpt_upsilon",
               "lpt_positions", "lpt_deposit", "lpt_route", "lpt_exchange", "lpt_finalize", "norm_hist", "srcs_poisson",
               "srcs_scan", "srcs_expand", "srcs_place", "srcs_local"):
        m, nl = par.stage_ms(nm)
        if m or nl:
            stages[nm] = {"ms_per_step": allmax(m) / args.steps, "launches_per_step": nl / args.steps}
    par.set_profiling(False)
    cells_rank = float(n) ** 3 / world
    out["ms_per_step"] = ms
    out["Mcells_per_s"] = float(n) ** 3 / ms / 1e3
    out["stages"] = stages
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    fft_ms = sum(stages[k]["ms_per_step"] for k in ("fft_z", "fft_y", "fft_x", "fft_yx") if k in stages)
    nfft = stages["fft_x" if "fft_x" in stages else "fft_yx"]["launches_per_step"] if fft_ms else 0
    if fft_ms and "fill_fft_z" not in stages:      # 3-D c2r as kernels of its own: 24 B/cell per transform
        out["fft_hbm_gbs_per_gpu"] = nfft * 24.0 * cells_rank / (fft_ms * 1e-3) / 1e9
        out["fft_frac_of_hbm_peak"] = out["fft_hbm_gbs_per_gpu"] / peak
    # mode fill + both transforms (8 + 2 x 24 B/cell), whichever kernels carried them (the fill is fused into the z pass
    # on one GPU up to n_grid = 1024)
    ff_ms = fft_ms + sum(stages[k]["ms_per_step"] for k in ("fill_modes", "fill_fft_z") if k in stages)
    out["fill_plus_fft_hbm_gbs_per_gpu"] = 56.0 * cells_rank / (ff_ms * 1e-3) / 1e9
    out["fill_plus_fft_frac_of_hbm_peak"] = out["fill_plus_fft_hbm_gbs_per_gpu"] / peak
    if "lognormal" in stages:
        out["lognormal_hbm_gbs_per_gpu"] = 8.0 * cells_rank / (stages["lognormal"]["ms_per_step"] * 1e-3) / 1e9
        out["lognormal_frac_of_hbm_peak"] = out["lognormal_hbm_gbs_per_gpu"] / peak
    if out["transpose"] == "p2p-fused":
        sent = nfft * 8.0 * n * n * par.nc / world * (world - 1) / world
        out["nvlink_gbs_per_direction_during_fused_pass"] = sent / (stages["fft_z"]["ms_per_step"] * 1e-3) / 1e9
    elif "fft_a2a" in stages and world > 1:
        sent = nfft * 8.0 * n * n * par.nc / world * (world - 1) / world
        out["a2a_gbs_per_direction"] = sent / (stages["fft_a2a"]["ms_per_step"] * 1e-3) / 1e9

    # ---- maps ----------------------------------------------------------------------------
    if args.kappa_nside:
        rf = np.interp([0.2, 0.4], t["z"], t["r"]).astype(np.float32)
        _, pos = cb.healpix.hp_shell_pixels(args.kappa_nside, 2)
        for name, fn in (("kappa_los", cb.kappa_get_beam_properties), ("isw_los", cb.isw_get_beam_properties)):
            fn(par, pos[:1024], rf)
            barrier()
            par.set_profiling(True)
            t0 = time.perf_counter()
            m = fn(par, pos, rf)
            wall = allmax(time.perf_counter() - t0)
            kms, _ = par.stage_ms(name)
            if name == "kappa_los":
                kms += par.stage_ms("kappa_tidal")[0]       # the Hessian precompute pass belongs to the kappa stage
            par.set_profiling(False)
            assert np.isfinite(m).all() and m.std() > 0
            out[name] = {"nside": args.kappa_nside, "kernel_ms_max": allmax(kms), "api_wall_ms": wall * 1e3,
                         "map_rms": float(m.astype(np.float64).std())}
    if args.imap_nside:
        nu_rest = 1420.405
        edges = np.linspace(nu_rest / 1.4, nu_rest / 1.05, 21)
        zf, z0 = nu_rest / edges[:-1] - 1, nu_rest / edges[1:] - 1
        r0 = np.interp(z0, t["z"], t["r"]).astype(np.float32)
        rfi = np.interp(zf, t["z"], t["r"]).astype(np.float32)
        par.set_imap(0, np.full(cb._lib.NA, 0.05), 1.0 + 0.5 * np.asarray(t["z"]), args.imap_nside, r0, rfi)
        cb.compute_density_normalization(par)
        barrier()
        par.set_profiling(True)
        t0 = time.perf_counter()
        data, nadd = cb.imap_set_cartesian(par, 0)
        wall = allmax(time.perf_counter() - t0)
        kms, _ = par.stage_ms("imap_paint")
        par.set_profiling(False)
        hits = float(nadd.astype(np.int64).sum())       # maps are all-reduced: every rank holds the full sky
        assert hits > 0 and np.isfinite(data).all()
        out["imap_paint"] = {"nside": args.imap_nside, "channels": 20, "subcell_hits": hits, "kernel_ms_max": allmax(kms),
                             "api_wall_ms": wall * 1e3}
    if rank == 0:
        print(json.dumps(out))
    par.free()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
