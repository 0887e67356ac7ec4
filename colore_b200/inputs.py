"""Synthetic input tables for CoLoRe runs.

The reference's ``param_example.cfg`` points at ``examples/simple/*`` tables that are
not shipped in its tree (SURVEY.md, facts box), so every run -- reference, oracle and
GPU -- uses tables generated here (recipe of ``example_CoLoRe.ipynb`` cells 3/5/9):

* P(k): Eisenstein & Hu (1998) no-wiggle transfer function on log-spaced k (CAMB two-column
  format ``k [h/Mpc]  P [(Mpc/h)^3]``; the reference renormalises it to sigma_8, cosmo.c:484-488)
* N(z) = A z^2 exp(-(z/z0)^1.5)  [deg^-2 per unit z],  b(z) = 1 + z
* T(z) = const [mK] and a frequency table for intensity maps
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field

import numpy as np

__all__ = ["Cosmology", "RunConfig", "eh_nowiggle_pk", "write_inputs", "write_param_file"]


@dataclass
class Cosmology:
    """cosmo_par section of the parameter file (param_example.cfg:58-72)."""

    omega_M: float = 0.3
    omega_L: float = 0.7
    omega_B: float = 0.05
    h: float = 0.7
    w: float = -1.0
    ns: float = 0.96
    sigma_8: float = 0.803869


@dataclass
class RunConfig:
    """global / field_par sections plus tracer choices (io.c:246-426)."""

    n_grid: int = 128
    z_min: float = 0.001
    z_max: float = 0.45
    seed: int = 1003
    r_smooth: float = 5.0
    smooth_potential: bool = True
    dens_type: int = 0
    lpt_buffer_fraction: float = 0.6
    lpt_interp_type: int = 1
    output_density: bool = False
    output_format: str = "ASCII"
    n_srcs: int = 1
    nz_amplitude: float = 3.0e3     # A in N(z) [deg^-2]; sets the mean number of sources per cell
    nz_z0: float = 0.25
    nz_zcut: float = 0.0            # > 0: N(z) tapers to zero above this redshift (no sources near / beyond z_max)
    imap_nside: int = 0             # 0 -> no intensity mapping
    imap_nchannels: int = 8
    kappa_nside: int = 0            # 0 -> no kappa maps
    isw_nside: int = 0
    z_out: tuple = (0.2, 0.4)
    srcs_lensing: bool = False      # include_lensing of every srcs section (srcs.c:531-614)
    srcs_skewers: bool = False      # store_skewers (srcs.c:507-529)
    gaussian_skewers: bool = False  # srcsN.gaussian_skewers (beaming.c:55-66)
    cstm_nside: int = 0             # 0 -> no custom projected map (cstm.c)
    lensing_n: int = 0              # > 0 -> "lensing" section (io.c:398-419; only read by -D_USE_FAST_LENSING builds)
    lensing_nside: int = 16
    lensing_spacing: str = "r"      # "r" or "log(1+z)" (cosmo.c:851-868)
    lensing_write: bool = True
    write_pred: bool = False        # global.write_pred / pred_dz / just_write_pred (io.c:283-289, predictions.c)
    pred_dz: float = 0.1
    just_write_pred: bool = False
    cosmo: Cosmology = field(default_factory=Cosmology)


def eh_nowiggle_pk(k_h: np.ndarray, c: Cosmology) -> np.ndarray:
    """Eisenstein & Hu (1998) zero-baryon-wiggle P(k) shape, arbitrary amplitude."""
    omh2 = c.omega_M * c.h ** 2
    obh2 = c.omega_B * c.h ** 2
    fb = c.omega_B / c.omega_M
    theta = 2.725 / 2.7
    s = 44.5 * np.log(9.83 / omh2) / np.sqrt(1.0 + 10.0 * obh2 ** 0.75)
    alpha = 1.0 - 0.328 * np.log(431.0 * omh2) * fb + 0.38 * np.log(22.3 * omh2) * fb ** 2
    k_mpc = k_h * c.h
    gamma_eff = c.omega_M * c.h * (alpha + (1.0 - alpha) / (1.0 + (0.43 * k_mpc * s) ** 4))
    q = k_h * theta ** 2 / gamma_eff
    l0 = np.log(2.0 * np.e + 1.8 * q)
    c0 = 14.2 + 731.0 / (1.0 + 62.5 * q)
    t = l0 / (l0 + c0 * q * q)
    return 2.0e4 * k_h ** c.ns * t * t * (1.0 / 0.02) ** c.ns * 1e-2


def write_inputs(dirname: str, cfg: RunConfig) -> dict:
    """Write P(k), N(z), b(z), T(z), nu tables into ``dirname``; return their paths."""
    os.makedirs(dirname, exist_ok=True)
    paths = {}
    k = np.logspace(-4, 2, 1024)
    pk = eh_nowiggle_pk(k, cfg.cosmo)
    paths["pk"] = os.path.join(dirname, "pk.dat")
    np.savetxt(paths["pk"], np.column_stack([k, pk]), fmt="%.10e")
    z = np.linspace(0.0, max(0.6, 1.2 * cfg.z_max), 1024)
    for ipop in range(cfg.n_srcs):
        amp = cfg.nz_amplitude / (1 + ipop)
        nz = amp * z ** 2 * np.exp(-(z / cfg.nz_z0) ** 1.5)
        if cfg.nz_zcut > 0:
            nz = nz * 0.5 * (1.0 - np.tanh((z - cfg.nz_zcut) / 0.004))
        bz = 1.0 + z + 0.2 * ipop
        paths[f"nz{ipop}"] = os.path.join(dirname, f"nz{ipop}.txt")
        paths[f"bz{ipop}"] = os.path.join(dirname, f"bz{ipop}.txt")
        np.savetxt(paths[f"nz{ipop}"], np.column_stack([z, nz]), fmt="%.10e")
        np.savetxt(paths[f"bz{ipop}"], np.column_stack([z, bz]), fmt="%.10e")
    if cfg.imap_nside > 0:
        paths["tz"] = os.path.join(dirname, "tz.txt")
        np.savetxt(paths["tz"], np.column_stack([z, np.full_like(z, 0.05)]), fmt="%.10e")
        paths["bz_im"] = os.path.join(dirname, "bz_im.txt")
        np.savetxt(paths["bz_im"], np.column_stack([z, 1.0 + 0.5 * z]), fmt="%.10e")
        # channels equally spaced in frequency between z=0.4 (low nu) and z=0.05 (high nu)
        nu_rest = 1420.405
        zlo, zhi = 0.05, min(0.4, 0.9 * cfg.z_max)
        edges = np.linspace(nu_rest / (1 + zhi), nu_rest / (1 + zlo), cfg.imap_nchannels + 1)
        paths["nu"] = os.path.join(dirname, "nu.txt")
        np.savetxt(paths["nu"], np.column_stack([edges[:-1], edges[1:]]), fmt="%.10e")
    if cfg.cstm_nside > 0:
        # custom projected tracer: a smooth radial kernel K(z) and its own bias b(z) (io.c:357-362, cosmo.c:631-662)
        paths["kz_cstm"] = os.path.join(dirname, "kz_cstm.txt")
        paths["bz_cstm"] = os.path.join(dirname, "bz_cstm.txt")
        kz = np.exp(-0.5 * ((z - 0.6 * cfg.z_max) / (0.2 * cfg.z_max)) ** 2)
        np.savetxt(paths["kz_cstm"], np.column_stack([z, kz]), fmt="%.10e")
        np.savetxt(paths["bz_cstm"], np.column_stack([z, 1.2 + 0.3 * z]), fmt="%.10e")
    return paths


def write_param_file(fname: str, cfg: RunConfig, paths: dict, prefix_out: str) -> None:
    """Write a libconfig parameter file with the sections read by io.c:266-426."""
    c = cfg.cosmo
    b = lambda v: "true" if v else "false"  # noqa: E731
    lines = [
        "global:", "{",
        f'  prefix_out= "{prefix_out}";',
        f'  output_format= "{cfg.output_format}";',
        f"  output_density= {b(cfg.output_density)}",
        f'  pk_filename= "{paths["pk"]}"',
        f"  z_min= {cfg.z_min!r}", f"  z_max= {cfg.z_max!r}", f"  seed= {cfg.seed}",
        f"  write_pred={b(cfg.write_pred)}", f"  pred_dz={cfg.pred_dz!r}", f"  just_write_pred= {b(cfg.just_write_pred)}", "}",
        "field_par:", "{",
        f"  r_smooth= {float(cfg.r_smooth)!r}",
        f"  smooth_potential= {b(cfg.smooth_potential)}",
        f"  n_grid= {cfg.n_grid}", f"  dens_type= {cfg.dens_type}",
        f"  lpt_buffer_fraction= {cfg.lpt_buffer_fraction!r}",
        f"  lpt_interp_type= {cfg.lpt_interp_type}", "  output_lpt= 0", "}",
        "cosmo_par:", "{",
        f"  omega_M= {c.omega_M!r}", f"  omega_L= {c.omega_L!r}", f"  omega_B= {c.omega_B!r}",
        f"  h= {c.h!r}", f"  w= {c.w!r}", f"  ns= {c.ns!r}", f"  sigma_8= {c.sigma_8!r}", "}",
    ]
    for ipop in range(cfg.n_srcs):
        lines += [f"srcs{ipop + 1}:", "{",
                  f'  nz_filename= "{paths[f"nz{ipop}"]}"',
                  f'  bias_filename= "{paths[f"bz{ipop}"]}"',
                  f"  include_lensing= {b(cfg.srcs_lensing)}", f"  store_skewers= {b(cfg.srcs_skewers)}",
                  f"  gaussian_skewers= {b(cfg.gaussian_skewers)}", "}"]
    if cfg.cstm_nside > 0:
        lines += ["custom1:", "{",
                  f'  kz_filename= "{paths["kz_cstm"]}"', f'  bias_filename= "{paths["bz_cstm"]}"',
                  f"  nside= {cfg.cstm_nside}", "}"]
    if cfg.imap_nside > 0:
        lines += ["imap1:", "{",
                  f'  tbak_filename= "{paths["tz"]}"', f'  bias_filename= "{paths["bz_im"]}"',
                  f'  freq_list= "{paths["nu"]}"', "  freq_rest= 1420.405",
                  f"  nside= {cfg.imap_nside}", "}"]
    if cfg.lensing_n > 0:
        lines += ["lensing:", "{", f"  n_lensing= {cfg.lensing_n}", f'  spacing_type= "{cfg.lensing_spacing}"',
                  f"  nside= {cfg.lensing_nside}", f"  write= {b(cfg.lensing_write)}", "}"]
    zs = ", ".join(repr(float(v)) for v in cfg.z_out)
    if cfg.kappa_nside > 0:
        lines += ["kappa:", "{", f"  z_out= [{zs}]", f"  nside= {cfg.kappa_nside}", "}"]
    if cfg.isw_nside > 0:
        lines += ["isw:", "{", f"  z_out= [{zs}]", f"  nside= {cfg.isw_nside}", "}"]
    with open(fname, "w") as f:
        f.write("\n".join(lines) + "\n")
