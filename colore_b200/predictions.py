"""Theory predictions written next to a run: ``write_predictions`` (predictions.c:25-173) on top of FFTLog
(fftlog.c:1-157, Hamilton's algorithm in the Copter formulation the reference ships).

Host-side, like the reference's: it only needs the tables of a run (P(k), growth, b(z)) and gates nothing on the GPU.
For every redshift z = 0, pred_dz, ... <= z_max and every population it writes

  <prefix>_pk_<kind>_pop<i>_z<z>.txt   k, P_tt (lognormal-transformed), P_tl, P_ll
  <prefix>_xi_<kind>_pop<i>_z<z>.txt   r, xi_tt, xi_ll b, xi_ll          (0.5 <= r <= 300 Mpc/h)
  <prefix>_gbias.txt                    z, r(z), D(z), b_i(z)

with the reference's formats (``%g``), so that the files can be compared line by line.
"""
from __future__ import annotations

import numpy as np

__all__ = ["fht", "pk2xi", "xi2pk", "pk_linear0", "write_predictions"]

_LANCZOS = np.array([0.99999999999980993227684700473478, 676.520368121885098567009190444019,
                     -1259.13921672240287047156078755283, 771.3234287776530788486528258894,
                     -176.61502916214059906584551354, 12.507343278686904814458936853,
                     -0.13857109526572011689554707, 9.984369578019570859563e-6, 1.50563273514931155834e-7])


def _gamma(z: np.ndarray) -> np.ndarray:
    """Complex Gamma function, Lanczos approximation with g = 7 (fftlog.c:10-32)."""
    z = np.asarray(z, np.complex128)
    out = np.empty_like(z)
    refl = z.real < 0.5
    if refl.any():
        zr = z[refl]
        out[refl] = np.pi / (np.sin(np.pi * zr) * _gamma(1.0 - zr))
    zz = z[~refl] - 1
    x = np.full(zz.shape, _LANCZOS[0], np.complex128)
    for n in range(1, 9):
        x = x + _LANCZOS[n] / (zz + float(n))
    t = zz + 7.5
    out[~refl] = np.sqrt(2 * np.pi) * np.power(t, zz + 0.5) * np.exp(-t) * x
    return out


def _lngamma_arg(x: float, y: np.ndarray):
    # |Gamma(x + iy)| underflows for |y| beyond ~450 (the log turns into -inf + 0i, as in the reference's clog); only the
    # phase is used there and those modes of the log-space FFT carry no power for smooth spectra
    with np.errstate(divide="ignore", invalid="ignore"):
        w = np.log(_gamma(x + 1j * np.asarray(y, np.float64)))
    return w.real, np.nan_to_num(w.imag)


def _goodkr(n: int, mu: float, q: float, L: float, kr: float) -> float:
    """Low-ringing value of k_c r_c (fftlog.c:49-62)."""
    xp, xm = (mu + 1 + q) / 2, (mu + 1 - q) / 2
    y = np.pi * n / (2 * L)
    _, argp = _lngamma_arg(xp, np.array([y]))
    _, argm = _lngamma_arg(xm, np.array([y]))
    arg = np.log(2 / kr) * n / L + (argp[0] + argm[0]) / np.pi
    iarg = np.round(arg)
    if arg != iarg:
        kr *= np.exp((arg - iarg) * L / n)
    return float(kr)


def _u_coefficients(n: int, mu: float, q: float, L: float, kcrc: float) -> np.ndarray:
    """fftlog.c:64-95."""
    y = np.pi / L
    k0r0 = kcrc * np.exp(-L)
    t = -2 * y * np.log(k0r0 / 2)
    m = np.arange(n // 2 + 1)
    u = np.empty(n, np.complex128)
    if q == 0:
        _, phi = _lngamma_arg((mu + 1) / 2, m * y)
        ang = m * t + 2 * phi
        u[:n // 2 + 1] = np.cos(ang) + 1j * np.sin(ang)
    else:
        lnrp, phip = _lngamma_arg((mu + 1 + q) / 2, m * y)
        lnrm, phim = _lngamma_arg((mu + 1 - q) / 2, m * y)
        rad, ang = np.exp(q * np.log(2) + lnrp - lnrm), m * t + phip - phim
        u[:n // 2 + 1] = rad * np.cos(ang) + 1j * rad * np.sin(ang)
    idx = np.arange(n // 2 + 1, n)
    u[idx] = np.conj(u[n - idx])
    if n % 2 == 0:
        u[n // 2] = u[n // 2].real
    return u


def fht(r: np.ndarray, a: np.ndarray, mu: float, q: float = 0.0, kcrc: float = 1.0, noring: bool = True):
    """Discrete Hankel transform on a logarithmic grid (fftlog.c:97-135). Returns (k, b)."""
    n = len(r)
    L = np.log(r[-1] / r[0]) * n / (n - 1.0)
    if noring:
        kcrc = _goodkr(n, mu, q, L, kcrc)
    u = _u_coefficients(n, mu, q, L, kcrc)
    b = np.fft.ifft(np.fft.fft(np.asarray(a, np.complex128)) * u)       # forward, * u / N, unnormalised inverse
    b = b[::-1].copy()
    k0r0 = kcrc * np.exp(-L)
    k = (k0r0 / r[0]) * np.exp(np.arange(n) * L / n)
    return k, b


def _xi_lm(l: int, m: int, k: np.ndarray, pk: np.ndarray):   # noqa: E741
    """fftlog.c:137-148."""
    a = np.power(k, m - 0.5) * pk
    r, b = fht(k, a, l + 0.5, 0.0, 1.0, True)
    return r, (np.power(2 * np.pi * r, -1.5) * b).real


def pk2xi(k, pk):
    """xi(r) from P(k) on log-spaced k (fftlog.c:150-152). Returns (r, xi)."""
    return _xi_lm(0, 2, np.asarray(k, np.float64), np.asarray(pk, np.float64))


def xi2pk(r, xi):
    """fftlog.c:154-159. Returns (k, pk)."""
    k, pk = _xi_lm(0, 2, np.asarray(r, np.float64), np.asarray(xi, np.float64))
    return k, pk * 8 * np.pi ** 3


def pk_linear0(t: dict, lgk: np.ndarray) -> np.ndarray:
    """cosmo.c:291-308, vectorised: linear interpolation of P against log10 k, k^ns below and k^-3 above the table."""
    lgk = np.asarray(lgk, np.float64)
    logk, pk = np.asarray(t["pk_logk"], np.float64), np.asarray(t["pk_pk"], np.float64)
    numk = len(pk)
    out = np.empty_like(lgk)
    lo, hi = lgk < t["logkmin"], lgk >= t["logkmax"]
    out[lo] = pk[0] * np.power(10.0, t["n_scal"] * (lgk[lo] - t["logkmin"]))
    out[hi] = pk[-1] * np.power(10.0, -3 * (lgk[hi] - t["logkmax"]))
    mid = ~(lo | hi)
    ik = np.minimum(((lgk[mid] - t["logkmin"]) * t["idlogk"]).astype(np.int64), numk - 2)
    out[mid] = pk[ik] + (lgk[mid] - logk[ik]) * (pk[ik + 1] - pk[ik]) * t["idlogk"]
    return out


def _lerp(t: dict, r: float, f: np.ndarray, f0: float, ff: float) -> float:
    """cosmo.c:30-38."""
    if r <= 0:
        return f0
    if r >= t["r"][-1]:
        return ff
    ir = int(r * t["glob_idr"])
    return float(f[ir] + (f[ir + 1] - f[ir]) * (r - t["r"][ir]) * t["glob_idr"])


def _r_of_z(t: dict, z: float) -> float:
    """cosmo.c:101-112."""
    a = 1.0 / (1 + z)
    if a >= 1:
        return 0.0
    if a <= 0:
        return float(t["a2r_r"][0])
    na = len(t["a2r_r"])
    ia = int(a * (na - 1))
    return float(t["a2r_r"][ia] + (t["a2r_r"][ia + 1] - t["a2r_r"][ia]) * (a - t["a2r_a"][ia]) * (na - 1.0))


def write_predictions(t: dict, prefix: str, n_grid: int, z_max: float, pred_dz: float, populations: dict) -> list:
    """predictions.c:25-173. ``t``: run tables (colore_b200.cosmo.cosmo_set or the reference's);
    ``populations``: {"srcs": [bz tables], "imap": [...], "custom": [...]} (NA-point b(z) tables on the r grid).
    Returns the list of files written."""
    kinds = [k for k in ("srcs", "imap", "custom") if populations.get(k)]
    if not kinds:
        return []
    nk = 10000
    ka = 1e-4 * np.power(100 / 1e-4, np.arange(nk) * 1.0 / (nk - 1))
    dx2 = (float(np.float32(t["l_box"])) / n_grid) ** 2
    rsm2_gg = t["r2_smooth"] + 0.9 * dx2 / 12.0
    rsm2_gm = t["r2_smooth"] + 2.0 * dx2 / 6.0
    rsm2_mm = t["r2_smooth"] + 1.88 * dx2 / 6.0
    files = [f"{prefix}_gbias.txt"]
    pk0 = pk_linear0(t, np.log10(ka))
    with open(files[0], "w") as fg:
        fg.write("#1-z 2-r(z) 3-g(z) ")
        col = 4
        for kind, tag in (("srcs", "bg"), ("imap", "bi"), ("custom", "bc")):
            for ipop in range(len(populations.get(kind) or [])):
                fg.write(f"{col}-{tag}_{ipop + 1}(z) ")
                col += 1
        fg.write("\n")
        z = 0.0
        while z <= z_max:
            r = _r_of_z(t, z)
            g = _lerp(t, r, t["d1"], 1.0, float(t["d1"][-1]))
            fg.write("%g %g %g " % (z, r, g))
            pklin = pk0 * g * g
            ra, xilin = pk2xi(ka, pklin)
            for kind in kinds:
                for ipop, bz in enumerate(populations[kind]):
                    bias = _lerp(t, r, bz, float(bz[0]), 1.0)
                    fg.write("%g " % bias)
                    rsm2 = rsm2_mm if kind == "custom" else rsm2_gg          # predictions.c:139 vs 68, 104
                    pk = pklin * bias * bias * np.exp(-rsm2 * ka * ka)
                    _, xi = pk2xi(ka, pk)
                    xi = np.exp(xi) - 1
                    _, pk = xi2pk(ra, xi)
                    fpk = f"{prefix}_pk_{kind}_pop{ipop}_z{z:.3f}.txt"
                    fxi = f"{prefix}_xi_{kind}_pop{ipop}_z{z:.3f}.txt"
                    p_tl = pklin * bias * np.exp(-rsm2_gm * ka * ka)
                    p_ll = pklin * np.exp(-rsm2_mm * ka * ka)
                    selk = (ka >= 1e-4) & (ka <= 100)                      # kminout = kmin, kmaxout = kmax
                    with open(fpk, "w") as f:
                        f.write("# k[h/Mpc] P_tt P_tl P_ll\n")
                        f.writelines("%g %g %g %g\n" % v for v in zip(ka[selk], pk[selk], p_tl[selk], p_ll[selk]))
                    sel = (ra >= 0.5) & (ra <= 300.0)
                    with open(fxi, "w") as f:
                        f.write("# r[Mpc/h] xi_tt xi_ll*b^2 xi_ll\n" if kind == "srcs" else "# k[Mpc/h] xi_tt xi_ll*b^2 xi_ll\n")
                        f.writelines("%g %g %g %g\n" % v for v in zip(ra[sel], xi[sel], xilin[sel] * bias, xilin[sel]))
                    files += [fpk, fxi]
            fg.write("\n")
            z += pred_dz
    return files
