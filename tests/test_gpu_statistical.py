"""GPU suite: the north star's SECOND correctness mode. The GPU draws its Gaussian modes from a counter-based
Philox stream, the reference from per-thread MT19937 streams (fourier.c:304), so the two can only agree
statistically: the measured P(k) of the Gaussian field and the angular power spectrum C_ell of the kappa
maps from GPU realisations must match those of reference-stream realisations (oracle, MT19937) within the
cosmic-variance bands. (N(z) of the catalogue is compared in tests/test_gpu_dropin.py.)
"""
import os

import numpy as np
import pytest

import colore_b200 as cb
from oracle.oracle import RNG_MT, Oracle, tables_from_dump

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _tables(n):
    g = dict(np.load(os.path.join(GOLD, "ref_n32_lognormal.npz")))
    t = dict(tables_from_dump(g))
    t["l_box"] = float(np.float32(2 * t["r_max"] * (1 + 2. / n)))
    t["pos_obs"] = 0.5 * t["l_box"]
    return g, t


def _pk_shells(field, nbins):
    """Band powers of a real n^3 field in |k| shells (grid units), and the number of modes per shell."""
    n = field.shape[0]
    fk = np.fft.rfftn(field.astype(np.float64))
    p = (fk.real ** 2 + fk.imag ** 2)
    k1 = np.fft.fftfreq(n, 1.0 / n)
    kz, ky, kx = np.meshgrid(k1, k1, np.arange(n // 2 + 1), indexing="ij")
    kk = np.sqrt(kx ** 2 + ky ** 2 + kz ** 2)
    w = np.where((kx == 0) | (kx == n // 2), 1.0, 2.0)          # Hermitian partners of the half spectrum
    edges = np.linspace(1.0, n // 2, nbins + 1)
    idx = np.digitize(kk.ravel(), edges) - 1
    ok = (idx >= 0) & (idx < nbins)
    num = np.bincount(idx[ok], weights=(p * w).ravel()[ok], minlength=nbins)
    cnt = np.bincount(idx[ok], weights=w.ravel()[ok], minlength=nbins)
    return num / cnt, cnt


def test_pk_of_gpu_stream_matches_reference_stream():
    n, nbins = 64, 12
    g, t = _tables(n)
    o = Oracle(t, n)
    par = cb.ParamCoLoRe(t, n, seed=1)
    pk_gpu, pk_ref = [], []
    for seed in (11, 12, 13):
        par.seed = seed
        _, s2 = cb.create_cartesian_fields(par)
        pk, cnt = _pk_shells(par.grid_get(cb.GRID_DENS)[:, :, :n], nbins)
        pk_gpu.append(pk)
        dk, pkk = o.fill_modes(RNG_MT, seed)                     # the reference's own generator
        d0, p0 = o.c2r(dk), o.c2r(pkk)
        o.normalize_fields(d0, p0)
        pk_ref.append(_pk_shells(d0[:, :, :n], nbins)[0])
    pk_gpu, pk_ref = np.mean(pk_gpu, 0), np.mean(pk_ref, 0)
    # independent Gaussian realisations: var(P)/P^2 = 2/N_modes per realisation and side
    sigma = np.sqrt(2.0 / (cnt * 3) * 2)
    ratio = pk_gpu / pk_ref
    assert np.all(np.abs(ratio - 1) < 5 * sigma + 1e-3), (ratio, sigma)
    # chi^2 of the whole band-power vector
    chi2 = np.sum(((ratio - 1) / sigma) ** 2)
    assert chi2 < 3.5 * nbins, chi2
    par.free()


def _cl_direct(m, vec, lmax):
    """C_ell of a full-sky map by direct quadrature over equal-area pixels (fine for nside <= 16)."""
    from scipy.special import sph_harm_y
    theta = np.arccos(np.clip(vec[:, 2], -1, 1))
    phi = np.arctan2(vec[:, 1], vec[:, 0])
    dom = 4 * np.pi / len(m)
    cl = np.zeros(lmax + 1)
    for ell in range(2, lmax + 1):
        mm = np.arange(0, ell + 1)
        y = sph_harm_y(ell, mm[:, None], theta[None, :], phi[None, :])
        alm = (np.conj(y) * m[None, :]).sum(1) * dom
        cl[ell] = (np.abs(alm[0]) ** 2 + 2 * np.sum(np.abs(alm[1:]) ** 2)) / (2 * ell + 1)
    return cl


def test_cl_kappa_of_gpu_stream_matches_reference_stream():
    n, nside, lmax = 64, 8, 16
    g, t = _tables(n)
    o = Oracle(t, n)
    par = cb.ParamCoLoRe(t, n, seed=1)
    _, vec = cb.healpix.hp_shell_pixels(nside, 2)
    rf = np.array([0.8 * t["r_max"]], np.float32)
    cl_gpu, cl_ref = [], []
    for seed in range(21, 27):
        par.seed = seed
        cb.create_cartesian_fields(par)
        cl_gpu.append(_cl_direct(cb.kappa_get_beam_properties(par, vec, rf)[0].astype(np.float64), vec, lmax))
        dk, pkk = o.fill_modes(RNG_MT, seed)
        d0, p0 = o.c2r(dk), o.c2r(pkk)
        o.normalize_fields(d0, p0)
        o.set_halo(p0)
        cl_ref.append(_cl_direct(o.kappa(p0, vec, rf)[0].astype(np.float64), vec, lmax))
    nreal = len(cl_gpu)
    cl_gpu, cl_ref = np.mean(cl_gpu, 0), np.mean(cl_ref, 0)
    # bins of 3 multipoles; Gaussian cosmic variance 2/((2l+1) n_real) per side
    for l0 in range(2, lmax - 1, 3):
        ls = np.arange(l0, l0 + 3)
        a, b = cl_gpu[ls].sum(), cl_ref[ls].sum()
        sig = np.sqrt(2.0 / (np.sum(2 * ls + 1) * nreal) * 2)
        assert abs(a / b - 1) < 5 * sig + 0.02, (l0, a / b, sig)
    par.free()
