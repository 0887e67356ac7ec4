"""CPU suite: colore_b200.predictions (write_predictions of predictions.c:25-173 + FFTLog, fftlog.c) against the files the
UNMODIFIED reference wrote for the same tables (tests/golden/make_golden_pred.py -> ref_predictions.npz)."""
import os

import numpy as np

from colore_b200 import predictions
from oracle.oracle import tables_from_dump


def _run(golden_dir, tmp_path):
    g = dict(np.load(os.path.join(golden_dir, "ref_predictions.npz")))
    t = tables_from_dump(g)
    pops = {"srcs": [t["srcs_bz_0"]], "imap": [t["imap_bz_0"]], "custom": [t["cstm_bz_0"]]}
    files = predictions.write_predictions(t, str(tmp_path / "out"), int(t["n_grid"]), t["z_max"], 0.2, pops)
    return g, files


def test_fftlog_round_trip():
    """xi2pk(pk2xi(P)) = P away from the ends of the grid (the property the reference relies on, predictions.c:71-73)."""
    k = 1e-4 * np.power(1e6, np.arange(4096) / 4095.0)
    pk = 2e4 * k / (1 + (k / 0.02) ** 2.5) * np.exp(-(k / 5.0) ** 2)
    r, xi = predictions.pk2xi(k, pk)
    k2, pk2 = predictions.xi2pk(r, xi)
    np.testing.assert_allclose(k2, k, rtol=1e-10)
    sel = (k > 1e-3) & (k < 2.0)
    np.testing.assert_allclose(pk2[sel], pk[sel], rtol=1e-6)
    # known answer: P(k) = exp(-k^2 s^2) has xi(r) = exp(-r^2 / (4 s^2)) / (8 pi^1.5 s^3). The discrete transform on
    # this k range reproduces it to a few per cent out to r = 5 s (its periodic wrap-around, same in the reference's
    # fftlog.c): a sanity bound, the parity check is the comparison with the reference's files below
    s = 3.0
    r, xi = predictions.pk2xi(k, np.exp(-(k * s) ** 2))
    sel = (r > 0.5) & (r < 15.0)
    np.testing.assert_allclose(xi[sel], np.exp(-r[sel] ** 2 / (4 * s * s)) / (8 * np.pi ** 1.5 * s ** 3), rtol=0.05)


def test_write_predictions_matches_reference_files(golden_dir, tmp_path):
    g, files = _run(golden_dir, tmp_path)
    names = sorted(os.path.basename(f) for f in files if "gbias" not in f)
    assert names == list(g["file_names"])                    # same files: kinds, populations, redshifts 0, 0.2, 0.4
    assert open(str(tmp_path / "out_gbias.txt")).read() == str(g["gbias_text"])
    for kind in ("srcs", "imap", "custom"):
        for what in ("pk", "xi"):
            mine = np.loadtxt(str(tmp_path / f"out_{what}_{kind}_pop0_z0.200.txt"))
            ref = g[f"{what}_{kind}_z0.200"]
            assert mine.shape == ref.shape, (kind, what)
            # %g keeps 6 significant digits; the lognormal-transformed spectrum crosses zero at high k
            scale = np.abs(ref).max(axis=0)
            np.testing.assert_allclose(mine, ref, rtol=2e-5, atol=1e-9 * scale.max())
    # and the text itself: the first 400 lines of one file, character for character
    head = "".join(open(str(tmp_path / "out_pk_srcs_pop0_z0.200.txt")).readlines()[:400])
    same = sum(a == b for a, b in zip(head.splitlines(), str(g["pk_srcs_text_head"]).splitlines()))
    assert same >= 380, same          # a few last digits differ (numpy's FFT and libm against the FFT shim)
