"""Host-side background cosmology and lookup tables (numpy / scipy, runs once per job).

Mirror of the reference's ``cosmo_set`` (cosmo.c:516-816), ``pk_linear_set`` (cosmo.c:444-495)
and the public functions of cosmo_mad.c (321-422) for flat / curved LambdaCDM (w = -1). The
north star keeps this stage on the host; the tables built here are the INPUTS of the GPU path
(uploaded by ``clr_create``) -- they are not part of the parity contract, only their lookup
(``get_bg`` / ``pk_linear0``, evaluated on the device) is.
"""
from __future__ import annotations

import numpy as np
from scipy.integrate import quad
from scipy.interpolate import CubicSpline

NA = 5001                # common.h:132
CSM_HMPC = 2997.92458    # cosmo_mad.h: c/H0 in Mpc/h
RTOD = 57.2957795        # common.h:124


class Background:
    """cosmo_mad.c background for w=-1 ("normalDE")."""

    def __init__(self, omega_M, omega_L, omega_B=0.05, h=0.7, w=-1.0):
        if w != -1.0:
            raise NotImplementedError("host tables: only w=-1 is implemented (cosmo_mad normalDE branch)")
        self.OM, self.OL, self.OB, self.h = omega_M, omega_L, omega_B, h
        self.OK = 1.0 - omega_M - omega_L
        if abs(self.OK) < 1e-6:
            self.OK = 0.0
        self._ph1 = self._parthor(1.0)
        self._d1 = None

    def _e2a3(self, a):  # OM + OL a^3 + OK a  (= E^2 a^3)
        return self.OM + self.OL * a ** 3 + self.OK * a

    def hubble(self, a):  # cosmo_mad.c:334-348 [h/Mpc]
        return np.sqrt(self._e2a3(a) / a ** 3) / CSM_HMPC

    def omega_m(self, a):  # cosmo_mad.c:321-332
        return self.OM / self._e2a3(a)

    def _parthor(self, a):  # cosmo_mad.c:261-290
        if a <= 0:
            return 0.0
        val, _ = quad(lambda x: 1.0 / np.sqrt(x * self._e2a3(x)), 0.0, a, epsabs=0, epsrel=1e-10, limit=200)
        return val * CSM_HMPC

    def radial_comoving_distance(self, a):  # cosmo_mad.c:358-364 [Mpc/h]
        return self._ph1 - self._parthor(a)

    def growth_factor(self, a):  # cosmo_mad.c:292-319, normalised so that D ~ a at early times
        if a <= 0:
            return 0.0
        val, _ = quad(lambda x: (x / self._e2a3(x)) ** 1.5, 0.0, a, epsabs=0, epsrel=1e-10, limit=200)
        return val * 2.5 * self.OM / (a * np.sqrt(a / self._e2a3(a)))

    def f_growth(self, a):  # cosmo_mad.c:392-414
        da = self.growth_factor(a)
        apow = a ** 3
        return 0.5 * (5 * self.OM * a / da - (3 * self.OM + 2 * self.OK * a)) / (self.OM + self.OL * apow + self.OK * a)


def _natural_spline(x, y):
    return CubicSpline(x, y, bc_type="natural", extrapolate=False)


def pk_linear_set(k, pk, sigma_8, n_scal):
    """cosmo.c:444-495: equi-log-spaced table renormalised to sigma_8. Returns dict of pk fields."""
    logk = np.log10(k)
    numk = len(k)
    logkmin, logkmax = logk[0], logk[-1]
    idlogk = (numk - 1) / (logkmax - logkmin)
    sp = _natural_spline(logk, pk)
    lk = logkmin + np.arange(numk) / idlogk
    pkarr = np.array(pk, dtype=np.float64)
    logkarr = np.array(logk, dtype=np.float64)
    pkarr[:-1] = sp(lk[:-1])
    logkarr[:-1] = lk[:-1]

    def pk0(lg):  # cosmo.c:291-308
        ik = int((lg - logkmin) * idlogk)
        if ik < 0:
            return pkarr[0] * 10 ** (n_scal * (lg - logkmin))
        if ik < numk - 1:
            return pkarr[ik] + (lg - logkarr[ik]) * (pkarr[ik + 1] - pkarr[ik]) * idlogk
        if ik == numk - 1:
            return pkarr[ik]
        return pkarr[-1] * 10 ** (-3 * (lg - logkmax))

    def wth(x):  # cosmo.c:275-282 top-hat window
        if x < 0.1:
            return 1. - 0.1 * x * x + 0.003571429 * x ** 4 - 6.61376E-5 * x ** 6 + 7.51563E-7 * x ** 8
        return 3 * (np.sin(x) - x * np.cos(x)) / x ** 3

    def integrand(lg):  # cosmo.c:350-368 with r=0
        kk = 10 ** lg
        return 0.1166503235296796 * pk0(lg) * kk ** 3 * wth(8 * kk) ** 2

    s2 = quad(integrand, logkmin - 6.0, logkmin, epsrel=1e-8, limit=400)[0]
    s2 += quad(integrand, logkmin, logkmax, epsrel=1e-8, limit=2000, points=np.linspace(logkmin, logkmax, 41)[1:-1])[0]
    pkarr *= sigma_8 ** 2 / s2
    return dict(pk_logk=logkarr, pk_pk=pkarr, numk=numk, logkmin=logkmin, logkmax=logkmax, idlogk=idlogk,
                sigma8_original=np.sqrt(s2))


def cosmo_set(cfg, k, pk, nz_tabs=(), bz_tabs=(), tz_tabs=(), bz_imap_tabs=()):
    """cosmo.c:516-732. ``cfg``: colore_b200.inputs.RunConfig. ``*_tabs``: sequences of (z, f) arrays.

    Returns the table dict consumed by colore_b200.pipeline (same keys as oracle.tables_from_dump).
    """
    c = cfg.cosmo
    bg = Background(c.omega_M, c.omega_L, c.omega_B, c.h, c.w)
    t = {}
    t["fgrowth_0"] = bg.f_growth(1.0)
    t["hubble_0"] = bg.hubble(1.0)
    t["r_min"] = bg.radial_comoving_distance(1 / (1 + cfg.z_min))
    t["r_max"] = bg.radial_comoving_distance(1 / (1 + cfg.z_max))
    t["prefac_lensing"] = 1.5 * t["hubble_0"] ** 2 * c.omega_M
    l_box = np.float32(2 * t["r_max"] * (1 + 2. / cfg.n_grid))   # flouble, common.h:274
    t["l_box"] = float(l_box)
    t["pos_obs"] = 0.5 * float(l_box)
    # a -> r table (cosmo.c:669-674)
    a_arr = np.arange(NA) / (NA - 1.0)
    r_a2r = np.array([bg.radial_comoving_distance(a) for a in a_arr])
    t["a2r_a"], t["a2r_r"] = a_arr, r_a2r
    growth0 = bg.growth_factor(1.0)
    glob_idr = (NA - 1) / r_a2r[0]
    t["glob_idr"] = glob_idr
    r_arr = np.arange(NA) / glob_idr
    # a_of_r_provisional (cosmo.c:88-99): linear interpolation on the descending r(a) table
    a_of_r = np.interp(r_arr, r_a2r[::-1], a_arr[::-1])
    a_of_r[0] = 1.0
    a_of_r[r_arr >= r_a2r[0]] = 1e-6
    a_of_r = np.maximum(a_of_r, 1e-6)
    z_arr = 1. / a_of_r - 1
    d_raw = np.array([bg.growth_factor(a) for a in a_of_r])
    gz = d_raw / growth0
    om = bg.omega_m(a_of_r)
    fz = 0.5 * (5 * bg.OM * a_of_r / d_raw - (3 * bg.OM + 2 * bg.OK * a_of_r)) / bg._e2a3(a_of_r)
    hhz = bg.hubble(a_of_r)
    t["r"], t["z"], t["d1"] = r_arr, z_arr, gz
    t["d2"] = -0.42857142857 * gz * gz * om ** (-0.00699300699)
    t["v1"] = (gz * hhz * fz) / (t["fgrowth_0"] * t["hubble_0"])
    t["pd"] = gz * hhz * (fz - 1)
    t["ih"] = 1. / hhz
    # tracer tables (cosmo.c:549-629, 693-716); out-of-range z -> NaN exactly like gsl_spline_eval
    for i, (z, nz) in enumerate(nz_tabs):
        z = np.asarray(z, float)
        a = 1. / (1 + z)
        hz = bg.hubble(a)
        rz = np.array([bg.radial_comoving_distance(x) for x in a])
        with np.errstate(divide="ignore", invalid="ignore"):
            f = np.asarray(nz, float) * RTOD * RTOD * hz / (rz * rz)
        if z[0] == 0:
            f[0] = f[1]
        t[f"srcs_nz_{i}"] = _natural_spline(z, f)(z_arr)
    for i, (z, b) in enumerate(bz_tabs):
        t[f"srcs_bz_{i}"] = _natural_spline(np.asarray(z, float), np.asarray(b, float))(z_arr)
    for i, (z, f) in enumerate(tz_tabs):
        t[f"imap_tz_{i}"] = _natural_spline(np.asarray(z, float), np.asarray(f, float))(z_arr)
    for i, (z, b) in enumerate(bz_imap_tabs):
        t[f"imap_bz_{i}"] = _natural_spline(np.asarray(z, float), np.asarray(b, float))(z_arr)
    t.update(pk_linear_set(np.asarray(k, float), np.asarray(pk, float), c.sigma_8, c.ns))
    t["OmegaM"], t["n_scal"] = c.omega_M, c.ns
    t["do_smoothing"] = 1 if cfg.r_smooth > 0 else 0
    t["r2_smooth"] = cfg.r_smooth ** 2 if cfg.r_smooth > 0 else cfg.r_smooth
    t["smooth_potential"] = int(cfg.smooth_potential)
    t["n_grid"], t["seed"], t["dens_type"] = cfg.n_grid, cfg.seed, cfg.dens_type
    t["z_min"], t["z_max"] = cfg.z_min, cfg.z_max
    t["_bg"] = bg
    return t


def shell_radius(t, z):
    """compute_tracer_cosmo (cosmo.c:818-849): comoving distance of a source plane, as flouble."""
    return np.float32(t["_bg"].radial_comoving_distance(1. / (1 + z)))


def choose_nside_base(nnodes: int) -> int:
    """io.c:224-244."""
    nside_base = 2
    while True:
        npix = 12 * nside_base * nside_base
        if npix % nnodes == 0:
            return nside_base
        pernode = npix // nnodes
        if pernode > 0 and (pernode + 1.) / pernode < 1.2:       # C: (0+1.)/0 = inf -> not balanced yet
            return nside_base
        nside_base *= 2
