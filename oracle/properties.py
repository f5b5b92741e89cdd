"""ORACLE (test infrastructure only -- never imported by the product path).

numpy fp64 restatement of the closed-form material laws of the reference
(`/root/reference/src/2d/material_properties.py`).  Every function cites the
reference line it restates.  Pinned against the reference's own module (imported
with a stub `fipy`) by `tests/golden/make_property_golden.py` ->
`tests/golden/properties_golden.npz` and `tests/test_oracle_properties.py`.

Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference`
legs of `bench.py` may import this module.
"""
from __future__ import annotations

import numpy as np

T_CRIT_NA = 2509.46      # material_properties.py:5
DH_MELT_NA = 113e3       # material_properties.py:6


def _asf(T):
    return np.asarray(T, dtype=np.float64)


# --- sodium, single phase -------------------------------------------------
def na_rho_solid(T):          # :9
    T = _asf(T)
    return 972.70 - 0.2154 * (T - 273.15)


def na_rho_liquid(T):         # :10
    T = _asf(T)
    s = 1 - T / T_CRIT_NA
    return 219 + 275.32 * s + 511.58 * s ** 0.5


def na_h_vap(T):              # :12-13
    T = _asf(T)
    s = 1 - T / T_CRIT_NA
    return (393.37 * s + 4398.6 * s ** 0.29302) * 1e3


def na_psat(T):               # :56
    T = _asf(T)
    return np.exp(11.9463 - 12633.73 / T - 0.4672 * np.log(T)) * 1e6


def na_dpdt(T):               # :15-16
    T = _asf(T)
    return (12633.73 / T ** 2 - 0.4672 / T) * np.exp(11.9463 - 12633.73 / T - 0.4672 * np.log(T)) * 1e6


def na_rho_vapor(T):          # :18-19 (Clausius-Clapeyron)
    T = _asf(T)
    return (na_h_vap(T) / (T * na_dpdt(T)) + 1 / na_rho_liquid(T)) ** -1


def na_k_solid(T):            # :21
    T = _asf(T)
    return 135.6 - 0.167 * (T - 273.15)


def na_k_liquid(T):           # :22
    T = _asf(T)
    return 124.67 - 0.11381 * T + 5.5226e-5 * T ** 2 - 1.1842e-8 * T ** 3


def na_cp_solid(T):           # :24
    T = _asf(T)
    return 1199 + 0.649 * (T - 273.15) + 1052.9e-5 * (T - 273.15) ** 2


def na_cp_liquid(T):          # :25
    T = _asf(T)
    return 1436.72 - 0.58 * (T - 273.15) + 4.672e-4 * (T - 273.15) ** 2


CPV = (16105.174483770606, -111.33277032117233, 0.2994911138258808, -0.00037555960577985903,
       2.418799137943085e-07, -7.773570963102343e-11, 9.892389871569863e-15)   # :27-29
CVV = (2944.6696506593726, -27.379453000849395, 0.08934864074702555, -0.00011764446335597157,
       7.521494832758196e-08, -2.338468836342328e-11, 2.8403280414122426e-15)  # :51-53


def na_cp_vapor(T):           # :27-29  (sum of monomials, same order as the reference)
    T = _asf(T)
    return (CPV[0] + CPV[1] * T + CPV[2] * T ** 2 + CPV[3] * T ** 3
            + CPV[4] * T ** 4 + CPV[5] * T ** 5 + CPV[6] * T ** 6)


def na_cv_vapor(T):           # :51-53
    T = _asf(T)
    return (CVV[0] + CVV[1] * T + CVV[2] * T ** 2 + CVV[3] * T ** 3
            + CVV[4] * T ** 4 + CVV[5] * T ** 5 + CVV[6] * T ** 6)


def na_mu_vapor(T):           # :55
    return 6.083e-9 * _asf(T) + 1.2606e-5


def na_h_liquid(T):           # :58-59
    T = _asf(T)
    return (-365.77 + 1.6582 * T - 4.2395e-4 * T ** 2 + 1.4847e-7 * T ** 3 + 2992.6 * T ** -1) * 1e3


def na_h_vapor(T):            # :61
    return na_h_liquid(T) + na_h_vap(T)


# --- steel ------------------------------------------------------------------
def steel_rho(T):             # :89
    T = _asf(T)
    return (7.9841 - 2.656e-4 * T - 1.158e-7 * T ** 2) * 1e3


def steel_k(T):               # :90
    return (8.116e-2 + 1.618e-4 * _asf(T)) * 100


def steel_cp(T):              # :91
    return 0.1348 * _asf(T) + 469.47


def steel_rho_eff(T):         # :93
    return steel_rho(T) + 3.75e6 / steel_cp(T)


# --- sodium across the melting point (T_m +- dT smear) ----------------------
def _lin(T, T_lo, T_hi, v_lo, v_hi):       # :104-105
    return v_lo + (v_hi - v_lo) * ((T - T_lo) / (T_hi - T_lo))


def na_k_interp(T, T_m, dT):               # :107-109
    T = _asf(T)
    ks, kl = na_k_solid(T), na_k_liquid(T)
    mid = (T >= T_m - dT) & (T <= T_m + dT)
    return ks * (T < T_m - dT) + kl * (T > T_m + dT) + _lin(T, T_m - dT, T_m + dT, ks, kl) * mid


def na_cp_interp(T, T_m, dT):              # :111-115 (latent-heat spike)
    T = _asf(T)
    cs, cl = na_cp_solid(T), na_cp_liquid(T)
    mid_val = (cs + cl) / 2 + DH_MELT_NA / (2 * dT)
    mid = (T >= T_m - dT) & (T <= T_m + dT)
    return cs * (T < T_m - dT) + cl * (T > T_m + dT) + mid_val * mid


def na_rho_interp(T, T_m, dT):             # :117-120
    T = _asf(T)
    rs, rl = na_rho_solid(T), na_rho_liquid(T)
    mid = (T >= T_m - dT) & (T <= T_m + dT)
    return rs * (T < T_m - dT) + rl * (T > T_m + dT) + _lin(T, T_m - dT, T_m + dT, rs, rl) * mid


# --- wick (porous composite) ------------------------------------------------
def wick_k(T, eps, T_m, dT):               # :181-193
    T = _asf(T)
    k_na = na_k_interp(T, T_m, dT)
    k_st = steel_k(T)
    num = (k_na + k_st) - (1 - eps) * (k_na - k_st)
    den = (k_na + k_st) + (1 - eps) * (k_na - k_st)
    return k_na * num / den


def wick_cp(T, eps, T_m, dT):              # :195-200
    return eps * na_cp_interp(T, T_m, dT) + (1 - eps) * steel_cp(T)


def wick_rho(T, eps, T_m, dT):             # :202-207
    return eps * na_rho_interp(T, T_m, dT) + (1 - eps) * steel_rho(T)


# --- vapor core -------------------------------------------------------------
def vc_k(T, region, R_v, M, Rgas):         # :140-154 (live part; :155-169 is dead code)
    T = _asf(T)
    mu = na_mu_vapor(T)
    P = na_psat(T)
    h_l = na_h_liquid(T)
    h_v = na_h_vapor(T)
    h_vap = na_h_vap(T)
    factor = (2 * R_v / 3) * np.sqrt((8 * Rgas * T) / (np.pi * M))
    base = (R_v ** 2 * P) / (4 if region == 'evap_cond' else 8) / mu + factor
    enthalpy = h_l if region == 'evap_cond' else h_v
    return base * (enthalpy * h_vap * M ** 2 * P) / (Rgas ** 2 * T ** 3)


vc_cp = na_cp_vapor      # :173-178
vc_rho = na_rho_vapor
