"""Generate the golden fixtures under tests/golden/ from the UNMODIFIED reference.

Runs oracle/_ref/ref_driver (the reference's own translation units compiled against
third_party/shim, see oracle/Makefile) with OMP_NUM_THREADS=1 on small synthetic configurations and
packs the per-stage dumps into compressed .npz files. Only runnable where /root/reference
exists (this container); the fixtures are what travels.

    python tests/golden/make_golden.py
"""
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from colore_b200.inputs import RunConfig, write_inputs, write_param_file  # noqa: E402

CASES = {
    # lognormal, 1 population, intensity maps, kappa + ISW planes: every §8(a) stage once
    "ref_n32_lognormal": RunConfig(n_grid=32, dens_type=0, nz_amplitude=60.0, imap_nside=8, imap_nchannels=4,
                                   kappa_nside=8, isw_nside=8, seed=1003),
    # clipped density (density.c:1034-1067), 2 populations, galaxies only -> cell-gradient RSD
    "ref_n32_clip": RunConfig(n_grid=32, dens_type=3, nz_amplitude=40.0, n_srcs=2, seed=77),
    # Lagrangian perturbation theory (density.c:376-1031): 1LPT + CIC, 2LPT + TSC, 2LPT + NGP
    "ref_n32_1lpt_cic": RunConfig(n_grid=32, dens_type=1, lpt_interp_type=1, nz_amplitude=20.0, seed=11),
    "ref_n32_2lpt_tsc": RunConfig(n_grid=32, dens_type=2, lpt_interp_type=2, nz_amplitude=20.0, seed=12),
    "ref_n32_2lpt_ngp": RunConfig(n_grid=32, dens_type=2, lpt_interp_type=0, nz_amplitude=20.0, seed=13),
    # no smoothing of the potential, different grid size
    "ref_n48_nosmooth": RunConfig(n_grid=48, dens_type=0, nz_amplitude=30.0, smooth_potential=False,
                                  r_smooth=-1.0, seed=5),
    # power-of-two grid without smoothing (do_smoothing=0, smooth_potential=false: fourier.c:347-351 skipped), so
    # that the CUDA path (powers of two only) runs that branch against the reference too
    "ref_n32_nosmooth": RunConfig(n_grid=32, dens_type=0, nz_amplitude=30.0, smooth_potential=False,
                                  r_smooth=-1.0, seed=9),
    # dense population: lambda up to several hundred per cell, i.e. gsl_ran_poisson's mu > 10 branch
    # (gamma / binomial reduction, common.c:187) in most occupied cells
    "ref_n32_dense": RunConfig(n_grid=32, dens_type=0, nz_amplitude=12000.0, seed=31),
    # per-source lensing (srcs.c:531-614) + density skewers (srcs.c:507-529) + a custom projected map (cstm.c:68-145);
    # the catalogue is kept above 128 KB so that the Src records come from fresh (zero) pages: the reference never
    # initialises kappa / dra / ddec before accumulating into them
    "ref_n32_lensing": RunConfig(n_grid=32, dens_type=0, nz_amplitude=25.0, srcs_lensing=True, srcs_skewers=True,
                                 cstm_nside=8, seed=41),
    # Gaussian skewers (beaming.c:55-66), no lensing
    "ref_n32_gskw": RunConfig(n_grid=32, dens_type=0, nz_amplitude=15.0, srcs_skewers=True, gaussian_skewers=True,
                              seed=43),
    # -D_USE_FAST_LENSING build (driver ref_driver_bmfl): adaptive-resolution lensing shells (lensing.c:76-250) and the
    # interpolation of shear / convergence / deflection onto the sources (srcs.c:666-721)
    "ref_n32_fastlens": RunConfig(n_grid=32, dens_type=0, nz_amplitude=25.0, nz_zcut=0.40, srcs_lensing=True, lensing_n=6,
                                  lensing_nside=16, seed=47),
    # the other compile-time bias models of common.h:414-431 (drivers built by `make -C oracle refbm`):
    # model 1 = pow(1+d,b) (no flag), model 3 = max(1+b d, 0) (-D_BIAS_MODEL_3)
    "ref_n32_bias1": RunConfig(n_grid=32, dens_type=0, nz_amplitude=60.0, imap_nside=8, imap_nchannels=4, seed=21),
    "ref_n32_bias3": RunConfig(n_grid=32, dens_type=0, nz_amplitude=60.0, imap_nside=8, imap_nchannels=4, seed=23),
}
# fixtures whose catalogue is too large to commit: arrays above 65536 elements are replaced by
# <key>__sha256 (digest of the raw bytes), <key>__size and <key>__head (first 4096 elements); tests compare digests
COMPACT = {"ref_n32_dense"}
DRIVER = {"ref_n32_bias1": "ref_driver_bm1", "ref_n32_bias3": "ref_driver_bm3", "ref_n32_fastlens": "ref_driver_bmfl"}


def compact(arrs):
    import hashlib
    out = {}
    for k, a in arrs.items():
        if a.size > 65536 and (k.startswith("s4_srcs") or k.startswith("s5_") or k.startswith("s6_srcs")):
            a = np.ascontiguousarray(a)
            out[k + "__sha256"] = np.frombuffer(hashlib.sha256(a.tobytes()).digest(), np.uint8)
            out[k + "__size"] = np.array([a.size], np.int64)
            out[k + "__head"] = a.ravel()[:4096].copy()
        else:
            out[k] = a
    return out


def main():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref", "refbm"])
    only = set(sys.argv[1:])                                   # optional: fixture names to (re)generate
    for name, cfg in CASES.items():
        if only and name not in only:
            continue
        drv = os.path.join(ROOT, "oracle", "_ref", DRIVER.get(name, "ref_driver"))
        tmp = tempfile.mkdtemp(prefix="golden_")
        try:
            paths = write_inputs(os.path.join(tmp, "in"), cfg)
            write_param_file(os.path.join(tmp, "param.cfg"), cfg, paths, os.path.join(tmp, "out"))
            os.makedirs(os.path.join(tmp, "dump"))
            env = dict(os.environ, OMP_NUM_THREADS="1")
            subprocess.check_call([drv, os.path.join(tmp, "param.cfg"), os.path.join(tmp, "dump")], env=env,
                                  stdout=subprocess.DEVNULL)
            arrs = {f[:-4]: np.load(os.path.join(tmp, "dump", f)) for f in sorted(os.listdir(os.path.join(tmp, "dump")))}
            if name in COMPACT:
                arrs = compact(arrs)
            out = os.path.join(ROOT, "tests", "golden", name + ".npz")
            np.savez_compressed(out, **arrs)
            print(name, "->", out, f"{os.path.getsize(out) / 1e6:.2f} MB",
                  "nsrc =", arrs.get("s4_srcs_ipix_0", np.zeros(0)).size)
        finally:
            shutil.rmtree(tmp)


if __name__ == "__main__":
    main()
