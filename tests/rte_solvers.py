"""Minimal broadband flux solvers for the "downstream fluxes within 0.01 W m-2" criterion.

TEST INFRASTRUCTURE.  rte-rrtmgp (rte_lw / rte_sw, call sites ecckd_rfmip_lw.F90:130, ecckd_rfmip_sw.F90:148) is
not available offline; these numpy solvers follow its formulation from the literature -- LW: no-scattering,
one Gauss angle (secant 1.66), linear-in-tau source from the layer and level Planck sources; SW: two-stream
(quadrature coefficients) + adding with a direct beam.  The SAME solver is applied to the oracle's and to the GPU's
optical properties; the test asserts that the resulting fluxes differ by less than 0.01 W m-2.
Arrays are (ncol, nlay, ngpt), top of atmosphere first.
"""
import numpy as np


def lw_fluxes(tau, lay_source, lev_source_inc, lev_source_dec, sfc_source, sfc_emis=0.98, D=1.66):
    """Returns (flux_up, flux_dn) at the nlay+1 levels [W m-2], summed over g-points."""
    ncol, nlay, ng = tau.shape
    t = np.exp(-D * tau)
    # linear-in-tau source (rte-rrtmgp lw_source_noscat): weight of the layer source vs the level source
    with np.errstate(divide="ignore", invalid="ignore"):
        fact = np.where(tau * D > 1e-8, (1.0 - t) / (tau * D) - t, 0.5 * tau * D)
    src_dn = (1.0 - t) * lev_source_inc + 2.0 * fact * (lay_source - lev_source_inc)   # leaving the bottom of the layer
    src_up = (1.0 - t) * lev_source_dec + 2.0 * fact * (lay_source - lev_source_dec)   # leaving the top of the layer
    rad_dn = np.zeros((ncol, nlay + 1, ng))
    for j in range(nlay):
        rad_dn[:, j + 1] = t[:, j] * rad_dn[:, j] + src_dn[:, j]
    rad_up = np.zeros((ncol, nlay + 1, ng))
    rad_up[:, nlay] = sfc_emis * sfc_source + (1.0 - sfc_emis) * rad_dn[:, nlay]
    for j in range(nlay - 1, -1, -1):
        rad_up[:, j] = t[:, j] * rad_up[:, j + 1] + src_up[:, j]
    return np.pi * rad_up.sum(-1), np.pi * rad_dn.sum(-1)


def sw_fluxes(tau, ssa, g, toa_src, mu0=0.6, sfc_alb=0.2):
    """Two-stream + adding.  Returns (flux_up, flux_dn_total, flux_dn_direct) at the levels, summed over g-points."""
    ncol, nlay, ng = tau.shape
    gamma1 = (8.0 - ssa * (5.0 + 3.0 * g)) * 0.25
    gamma2 = 3.0 * (ssa * (1.0 - g)) * 0.25
    gamma3 = (2.0 - 3.0 * mu0 * g) * 0.25
    gamma4 = 1.0 - gamma3
    k = np.sqrt(np.maximum((gamma1 - gamma2) * (gamma1 + gamma2), 1e-12))
    e = np.exp(-tau * k)
    e2 = e * e
    rt = 1.0 / (k + gamma1 + (k - gamma1) * e2)
    Rdif = rt * gamma2 * (1.0 - e2)
    Tdif = rt * 2.0 * k * e
    Tnoscat = np.exp(-tau / mu0)
    k_mu = k * mu0
    denom = np.where(np.abs(1.0 - k_mu * k_mu) > 1e-8, 1.0 - k_mu * k_mu, 1e-8)
    rt2 = ssa * rt / denom
    a1 = gamma1 * gamma3 + gamma2 * gamma4  # noqa: F841  (kept for readability of the formulation)
    Rdir = rt2 * ((1.0 - k_mu) * (gamma1 * gamma3 + gamma2 * gamma4 + k * gamma3) -
                  (1.0 + k_mu) * (gamma1 * gamma3 + gamma2 * gamma4 - k * gamma3) * e2 -
                  2.0 * (k * gamma3 - (gamma1 * gamma3 + gamma2 * gamma4) * k_mu) * e * Tnoscat)
    Tdir = -rt2 * ((1.0 + k_mu) * (gamma1 * gamma4 + gamma2 * gamma3 + k * gamma4) * Tnoscat -
                   (1.0 - k_mu) * (gamma1 * gamma4 + gamma2 * gamma3 - k * gamma4) * e2 * Tnoscat -
                   2.0 * (k * gamma4 + (gamma1 * gamma4 + gamma2 * gamma3) * k_mu) * e)
    dir_flux = np.zeros((ncol, nlay + 1, ng))
    dir_flux[:, 0] = toa_src * mu0
    for j in range(nlay):
        dir_flux[:, j + 1] = dir_flux[:, j] * Tnoscat[:, j]
    src_up = Rdir * dir_flux[:, :-1]
    src_dn = Tdir * dir_flux[:, :-1]
    # adding, surface upwards
    albedo = np.zeros((ncol, nlay + 1, ng)); src = np.zeros((ncol, nlay + 1, ng))
    albedo[:, nlay] = sfc_alb
    src[:, nlay] = sfc_alb * dir_flux[:, nlay]
    for j in range(nlay - 1, -1, -1):
        d = 1.0 / (1.0 - Rdif[:, j] * albedo[:, j + 1])
        albedo[:, j] = Rdif[:, j] + Tdif[:, j] ** 2 * albedo[:, j + 1] * d
        src[:, j] = src_up[:, j] + Tdif[:, j] * d * (src[:, j + 1] + albedo[:, j + 1] * src_dn[:, j])
    up = np.zeros((ncol, nlay + 1, ng)); dn = np.zeros((ncol, nlay + 1, ng))
    up[:, 0] = src[:, 0]
    for j in range(nlay):
        d = 1.0 / (1.0 - Rdif[:, j] * albedo[:, j + 1])
        dn[:, j + 1] = (Tdif[:, j] * dn[:, j] + Rdif[:, j] * src[:, j + 1] + src_dn[:, j]) * d
        up[:, j + 1] = dn[:, j + 1] * albedo[:, j + 1] + src[:, j + 1]
    return up.sum(-1), (dn + dir_flux).sum(-1), dir_flux.sum(-1)
