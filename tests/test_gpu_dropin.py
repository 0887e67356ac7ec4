"""GPU suite: the drop-in executable. integration/_build/CoLoRe_b200 = the reference's own main.c /
io.c / cosmo.c (unchanged) linked with integration/colore_gpu_glue.c + libcolore_b200.so instead of
fourier.c, density.c, srcs.c, imap.c, kappa.c, isw.c, beaming.c. It must run `CoLoRe param.cfg` end to
end and agree STATISTICALLY with the CPU reference binary (different RNG streams: MT19937 per
OpenMP thread vs. the counter-based stream), which is the north star's second correctness mode.
"""
import os
import struct
import subprocess
import tempfile

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B200 = os.path.join(ROOT, "integration", "_build", "CoLoRe_b200")
REF = os.path.join(ROOT, "oracle", "_ref", "CoLoRe_ref")


def _read_dens(fname):
    """io.c:565-595: int NNodes, int sizeof(flouble), double l_box, int n_grid, int nz_here, int iz0_here, data."""
    with open(fname, "rb") as f:
        nnodes, size_fl = struct.unpack("ii", f.read(8))
        (l_box,) = struct.unpack("d", f.read(8))
        n, nz, iz0 = struct.unpack("iii", f.read(12))
        data = np.fromfile(f, dtype=np.float32 if size_fl == 4 else np.float64, count=nz * n * n)
    return dict(nnodes=nnodes, l_box=l_box, n=n, nz=nz, iz0=iz0, data=data.reshape(nz, n, n))


def _run(exe, tmp, tag, cfg):
    from colore_b200.inputs import write_inputs, write_param_file
    paths = write_inputs(os.path.join(tmp, "in"), cfg)
    prm = os.path.join(tmp, f"param_{tag}.cfg")
    write_param_file(prm, cfg, paths, os.path.join(tmp, f"out_{tag}"))
    r = subprocess.run([exe, prm], capture_output=True, text=True, cwd=tmp, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return r.stdout


@pytest.mark.skipif(not (os.path.exists(B200) and os.path.exists(REF)), reason="drop-in binaries not built")
def test_dropin_executable_matches_reference_statistically():
    from colore_b200.inputs import RunConfig
    cfg = RunConfig(n_grid=64, dens_type=0, nz_amplitude=400.0, imap_nside=16, imap_nchannels=4, kappa_nside=16,
                    isw_nside=16, output_density=True, seed=321)
    with tempfile.TemporaryDirectory() as tmp:
        out_gpu = _run(B200, tmp, "gpu", cfg)
        out_ref = _run(REF, tmp, "ref", cfg)
        assert "(GPU)" in out_gpu
        # catalogues: same expected number of objects
        n_gpu = sum(1 for _ in open(os.path.join(tmp, "out_gpu_srcs_s1_0.txt"))) - 1
        n_ref = sum(1 for _ in open(os.path.join(tmp, "out_ref_srcs_s1_0.txt"))) - 1
        assert n_gpu > 1000 and abs(n_gpu - n_ref) < 6 * np.sqrt(n_ref) + 0.01 * n_ref, (n_gpu, n_ref)
        cg = np.loadtxt(os.path.join(tmp, "out_gpu_srcs_s1_0.txt"))
        cr = np.loadtxt(os.path.join(tmp, "out_ref_srcs_s1_0.txt"))
        # N(z): histograms of z0 agree within Poisson + cosmic scatter
        bins = np.linspace(0, 0.5, 11)
        hg, _ = np.histogram(cg[:, 3], bins)
        hr, _ = np.histogram(cr[:, 3], bins)
        ok = hr > 200
        assert np.all(np.abs(hg[ok] - hr[ok]) < 8 * np.sqrt(hr[ok]) + 0.05 * hr[ok]), (hg, hr)
        # sky coverage and RSD amplitude
        assert cg[:, 1].min() >= 0 and cg[:, 1].max() <= 360 and np.abs(cg[:, 2]).max() <= 90
        assert 0.5 < cg[:, 4].std() / cr[:, 4].std() < 2.0
        # density dumps in the reference's native format: same header, same field variance
        for kind in ("gaussian", "lightcone"):
            dg = _read_dens(os.path.join(tmp, f"out_gpu_dens_{kind}_0.dat"))
            dr = _read_dens(os.path.join(tmp, f"out_ref_dens_{kind}_0.dat"))
            assert (dg["n"], dg["nz"], dg["iz0"], dg["nnodes"]) == (dr["n"], dr["nz"], dr["iz0"], dr["nnodes"])
            assert abs(dg["l_box"] - dr["l_box"]) < 1e-6 * dr["l_box"]
            assert abs(dg["data"].std() / dr["data"].std() - 1) < 0.15
        # maps written by the unchanged io.c
        for name in ("kappa_z000", "kappa_z001", "isw_z000", "imap_s1_nu000"):
            a, b = os.path.join(tmp, f"out_gpu_{name}.fits"), os.path.join(tmp, f"out_ref_{name}.fits")
            assert os.path.getsize(a) == os.path.getsize(b) > 2880
