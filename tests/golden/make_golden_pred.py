"""Golden theory predictions: the UNMODIFIED reference (oracle/_ref/CoLoRe_ref, predictions.c + fftlog.c on the FFT shim)
run with write_pred / just_write_pred on a small configuration; the files of ONE redshift and population plus the bias
table are packed into tests/golden/ref_predictions.npz together with the run's tables.

    python tests/golden/make_golden_pred.py
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

CFG = RunConfig(n_grid=32, dens_type=0, nz_amplitude=25.0, cstm_nside=8, imap_nside=8, imap_nchannels=4, seed=3,
                write_pred=True, pred_dz=0.2, just_write_pred=True)


def main():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"])
    tmp = tempfile.mkdtemp(prefix="golden_pred_")
    try:
        paths = write_inputs(os.path.join(tmp, "in"), CFG)
        write_param_file(os.path.join(tmp, "param.cfg"), CFG, paths, os.path.join(tmp, "out"))
        env = dict(os.environ, OMP_NUM_THREADS="1")
        subprocess.check_call([os.path.join(ROOT, "oracle", "_ref", "CoLoRe_ref"), os.path.join(tmp, "param.cfg")], env=env,
                              stdout=subprocess.DEVNULL)
        # tables of the same run (the driver dumps them before any stage)
        os.makedirs(os.path.join(tmp, "dump"))
        cfg2 = RunConfig(**{**CFG.__dict__, "write_pred": False, "just_write_pred": False})
        write_param_file(os.path.join(tmp, "param2.cfg"), cfg2, paths, os.path.join(tmp, "out2"))
        subprocess.check_call([os.path.join(ROOT, "oracle", "_ref", "ref_driver"), os.path.join(tmp, "param2.cfg"),
                               os.path.join(tmp, "dump")], env=env, stdout=subprocess.DEVNULL)
        arrs = {f[:-4]: np.load(os.path.join(tmp, "dump", f)) for f in os.listdir(os.path.join(tmp, "dump"))
                if f.startswith("tab_") or f.startswith("pk_") or f == "scalars.npy"}
        names = sorted(f for f in os.listdir(tmp) if f.startswith("out_pk_") or f.startswith("out_xi_"))
        arrs["file_names"] = np.array(names)
        for kind in ("srcs", "imap", "custom"):
            for what in ("pk", "xi"):
                arrs[f"{what}_{kind}_z0.200"] = np.loadtxt(os.path.join(tmp, f"out_{what}_{kind}_pop0_z0.200.txt"))
        arrs["gbias_text"] = np.array(open(os.path.join(tmp, "out_gbias.txt")).read())
        arrs["pk_srcs_text_head"] = np.array("".join(open(os.path.join(tmp, "out_pk_srcs_pop0_z0.200.txt")).readlines()[:400]))
        out = os.path.join(ROOT, "tests", "golden", "ref_predictions.npz")
        np.savez_compressed(out, **arrs)
        print(out, f"{os.path.getsize(out) / 1e6:.2f} MB", names)
    finally:
        shutil.rmtree(tmp)


if __name__ == "__main__":
    main()
