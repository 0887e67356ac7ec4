#!/usr/bin/env python
"""Map stages of BASELINE configs 3 and 4 at full size on one B200 ("maps stage
reported separately as Mpixel*samples/s"): kappa + ISW line-of-sight integrals (kappa.c:78-175,
isw.c:78-147) and intensity-map painting (imap.c:135-245) on the fields of a lognormal run.

    python tools/bench_maps.py --n-grid 1024 --nside 1024 --imap-nside 256 --steps 3

Prints one JSON line: device time of each kernel (CUDA events on the launching stream, through the
C ABI stage timers), samples per second and the wall time of the whole API call with host buffers.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-grid", type=int, default=1024)
    ap.add_argument("--nside", type=int, default=1024)
    ap.add_argument("--imap-nside", type=int, default=256)
    ap.add_argument("--imap-channels", type=int, default=20)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--skip-imap", action="store_true")
    ap.add_argument("--skip-los", action="store_true")
    args = ap.parse_args()
    import colore_b200 as cb
    from bench import build_tables, make_config

    n = args.n_grid
    cfg = make_config(n)
    t = build_tables(cfg)
    par = cb.ParamCoLoRe(t, n, dens_type=0, seed=cfg.seed)
    cb.dist.init_comm(par, 0, 1)
    cb.create_cartesian_fields(par)
    cb.compute_physical_density_field(par)
    out = {"n_grid": n, "steps": args.steps}
    z_out = np.array([0.2, 0.4])
    rf = np.interp(z_out, t["z"], t["r"]).astype(np.float32)
    nr = n // 2

    if not args.skip_los:
        t0 = time.perf_counter()
        _, pos = cb.healpix.hp_shell_pixels(args.nside, 2)
        out["pix2vec_host_s"] = time.perf_counter() - t0
        npix = pos.shape[0]
        # samples actually integrated: planes end at rf[-1] (kappa.c:99-106)
        i_r_max = min(int(rf[-1] / (par.r_max / nr) + 0.5), nr - 1)
        samples = float(npix) * (i_r_max + 1)
        for name, fn in (("kappa_los", cb.kappa_get_beam_properties), ("isw_los", cb.isw_get_beam_properties)):
            fn(par, pos[:1024], rf)                       # warm-up (module load)
            par.set_profiling(True)
            t0 = time.perf_counter()
            for _ in range(args.steps):
                m = fn(par, pos, rf)
            wall = (time.perf_counter() - t0) / args.steps
            ms, nl = par.stage_ms(name)
            tid, _ = par.stage_ms("kappa_tidal") if name == "kappa_los" else (0.0, 0)     # Hessian precompute pass
            par.set_profiling(False)
            ms = (ms + tid) / args.steps                  # per call (kappa runs one pair of launches per chunk of planes)
            out[name] = {"nside": args.nside, "npix": npix, "planes": len(rf), "samples": samples, "kernel_ms": ms,
                         "of_which_hessian_precompute_ms": tid / args.steps, "launches_per_call": nl / args.steps,
                         "Gsamples_per_s": samples / ms / 1e6, "api_wall_ms": wall * 1e3,
                         "map_rms": float(m.astype(np.float64).std())}

    if not args.skip_imap:
        nu_rest = 1420.405
        edges = np.linspace(nu_rest / 1.4, nu_rest / 1.05, args.imap_channels + 1)
        zf, z0 = nu_rest / edges[:-1] - 1, nu_rest / edges[1:] - 1
        r0 = np.interp(z0, t["z"], t["r"]).astype(np.float32)
        rfi = np.interp(zf, t["z"], t["r"]).astype(np.float32)
        tz = np.full(cb._lib.NA, 0.05)
        bz = 1.0 + 0.5 * np.asarray(t["z"])
        par.set_imap(0, tz, bz, args.imap_nside, r0, rfi)
        cb.compute_density_normalization(par)
        cb.imap_set_cartesian(par, 0)
        par.set_profiling(True)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            data, nadd = cb.imap_set_cartesian(par, 0)
        wall = (time.perf_counter() - t0) / args.steps
        ms, nl = par.stage_ms("imap_paint")
        par.set_profiling(False)
        ms /= max(nl, 1)
        hits = float(nadd.astype(np.int64).sum())
        out["imap_paint"] = {"nside": args.imap_nside, "channels": args.imap_channels, "subcell_hits": hits,
                             "kernel_ms": ms, "Ghits_per_s": hits / ms / 1e6, "Mcells_per_s": n ** 3 / ms / 1e3,
                             "api_wall_ms": wall * 1e3}
    print(json.dumps(out))
    par.free()


if __name__ == "__main__":
    main()
