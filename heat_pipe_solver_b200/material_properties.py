"""Material property surface of the reference (`/root/reference/src/2d/material_properties.py`).

Same getters, same dictionary keys, same call signatures:

* ``get_sodium_properties()`` -> 14 callables  (reference :70-86)
* ``get_steel_properties()``  -> 4 callables   (:95-101)
* ``get_wick_properties()``   -> ``f(T, Na_props, steel_props, params)``  (:181-207)
* ``get_vc_properties()``     -> ``thermal_conductivity(T, mesh, region, sodium_props, dims, params, consts)``,
  ``specific_heat(T)``, ``density(T)``  (:173-178)

Every callable accepts what the reference's accept -- python floats and numpy arrays
(`utils/plot_material_properties.py:38-44`) -- and additionally torch tensors.  For torch tensors
that live on a CUDA device the laws are evaluated by the library's own kernel
(``hp_eval_property``, the same device functions the solver uses, ``csrc/hp_props.cuh``); host
inputs are evaluated with the array library they came from.  The hot path never calls this module:
inside the solver the laws are fused into the assembly kernels.
"""
from __future__ import annotations

import math

import numpy as np

T_Na_crit = 2509.46      # K      (reference :5)
delta_H_melt = 113e3     # J/kg   (reference :6)


# --------------------------------------------------------------------------- array-library shim
def _is_torch(T):
    return type(T).__module__.startswith('torch')


def _xp(T):
    if _is_torch(T):
        import torch
        return torch
    return np


def _device_eval(name, T, consts=None):
    """Route CUDA tensors through the library kernel; returns None for host inputs."""
    if not (_is_torch(T) and T.is_cuda):
        return None
    from . import _device_props
    return _device_props.evaluate(name, T, consts or {})


# --------------------------------------------------------------------------- sodium (reference :9-61)
def rho_sodium_s(T):
    r = _device_eval('na_rho_s', T)
    return r if r is not None else 972.70 - 0.2154 * (T - 273.15)


def rho_sodium_l(T):
    r = _device_eval('na_rho_l', T)
    if r is not None:
        return r
    s = 1 - T / T_Na_crit
    return 219 + 275.32 * s + 511.58 * s ** 0.5


def h_vap_sodium(T):
    r = _device_eval('na_h_vap', T)
    if r is not None:
        return r
    s = 1 - T / T_Na_crit
    return (393.37 * s + 4398.6 * s ** 0.29302) * 1e3


def P_sat_sodium(T):
    r = _device_eval('na_psat', T)
    if r is not None:
        return r
    xp = _xp(T)
    return xp.exp(11.9463 - 12633.73 / T - 0.4672 * xp.log(T)) * 1e6


def dPdT_sodium(T):
    r = _device_eval('na_dpdt', T)
    if r is not None:
        return r
    xp = _xp(T)
    return (12633.73 / T ** 2 - 0.4672 / T) * xp.exp(11.9463 - 12633.73 / T - 0.4672 * xp.log(T)) * 1e6


def rho_sodium_v(T):
    r = _device_eval('na_rho_v', T)
    if r is not None:
        return r
    return (h_vap_sodium(T) / (T * dPdT_sodium(T)) + 1 / rho_sodium_l(T)) ** -1


def k_sodium_s(T):
    r = _device_eval('na_k_s', T)
    return r if r is not None else 135.6 - 0.167 * (T - 273.15)


def k_sodium_l(T):
    r = _device_eval('na_k_l', T)
    return r if r is not None else 124.67 - 0.11381 * T + 5.5226e-5 * T ** 2 - 1.1842e-8 * T ** 3


def c_p_sodium_s(T):
    r = _device_eval('na_cp_s', T)
    if r is not None:
        return r
    c = T - 273.15
    return 1199 + 0.649 * c + 1052.9e-5 * c ** 2


def c_p_sodium_l(T):
    r = _device_eval('na_cp_l', T)
    if r is not None:
        return r
    c = T - 273.15
    return 1436.72 - 0.58 * c + 4.672e-4 * c ** 2


_CP_V = (16105.174483770606, -111.33277032117233, 0.2994911138258808, -0.00037555960577985903,
         2.418799137943085e-07, -7.773570963102343e-11, 9.892389871569863e-15)
_CV_V = (2944.6696506593726, -27.379453000849395, 0.08934864074702555, -0.00011764446335597157,
         7.521494832758196e-08, -2.338468836342328e-11, 2.8403280414122426e-15)


def _monomials(c, T):
    total = c[0] + c[1] * T
    for n in range(2, len(c)):
        total = total + c[n] * T ** n
    return total


def c_p_sodium_v(T):
    r = _device_eval('na_cp_v', T)
    return r if r is not None else _monomials(_CP_V, T)


def c_v_sodium_v(T):
    r = _device_eval('na_cv_v', T)
    return r if r is not None else _monomials(_CV_V, T)


def mu_sodium_v(T):
    r = _device_eval('na_mu_v', T)
    return r if r is not None else 6.083e-9 * T + 1.2606e-5


def h_l_sodium(T):
    r = _device_eval('na_h_l', T)
    if r is not None:
        return r
    return (-365.77 + 1.6582 * T - 4.2395e-4 * T ** 2 + 1.4847e-7 * T ** 3 + 2992.6 * T ** -1) * 1e3


def h_v_sodium(T):
    r = _device_eval('na_h_v', T)
    return r if r is not None else h_l_sodium(T) + h_vap_sodium(T)


def get_sodium_properties():
    return {
        'density_solid': rho_sodium_s,
        'density_liquid': rho_sodium_l,
        'density_vapor': rho_sodium_v,
        'heat_of_vaporization': h_vap_sodium,
        'thermal_conductivity_solid': k_sodium_s,
        'thermal_conductivity_liquid': k_sodium_l,
        'specific_heat_solid': c_p_sodium_s,
        'specific_heat_liquid': c_p_sodium_l,
        'specific_heat_vapor_pressure': c_p_sodium_v,
        'specific_heat_vapor_volume': c_v_sodium_v,
        'enthalpy_liquid': h_l_sodium,
        'enthalpy_vapor': h_v_sodium,
        'viscosity': mu_sodium_v,
        'vapor_pressure': P_sat_sodium,
    }


# --------------------------------------------------------------------------- steel (reference :89-101)
def rho_steel(T):
    r = _device_eval('steel_rho', T)
    return r if r is not None else (7.9841 - 2.656e-4 * T - 1.158e-7 * T ** 2) * 1e3


def k_steel(T):
    r = _device_eval('steel_k', T)
    return r if r is not None else (8.116e-2 + 1.618e-4 * T) * 100


def c_p_steel(T):
    r = _device_eval('steel_cp', T)
    return r if r is not None else 0.1348 * T + 469.47


def rho_eff_steel(T):
    r = _device_eval('steel_rho_eff', T)
    return r if r is not None else rho_steel(T) + 3.75e6 / c_p_steel(T)


def get_steel_properties():
    return {
        'density_evap_adia': rho_eff_steel,
        'density_cond': rho_steel,
        'thermal_conductivity': k_steel,
        'specific_heat': c_p_steel,
    }


# --------------------------------------------------------------------------- melting smear (reference :104-120)
def _interp_linear(T, T_low, T_high, val_low, val_high):
    return val_low + (val_high - val_low) * ((T - T_low) / (T_high - T_low))


def _bands(T, params):
    T_m, dT = params['T_melting_sodium'], params['delta_T_sodium']
    below, above = T < T_m - dT, T > T_m + dT
    inside = (T >= T_m - dT) & (T <= T_m + dT)
    if _is_torch(T):
        below, above, inside = (m.to(T.dtype) for m in (below, above, inside))
    return T_m, dT, below, above, inside


def get_k_Na_i(T, params, k_l, k_s):
    T_m, dT, below, above, inside = _bands(T, params)
    return k_s * below + k_l * above + _interp_linear(T, T_m - dT, T_m + dT, k_s, k_l) * inside


def get_c_p_Na_i(T, sodium_props, params, c_p_l, c_p_s):
    T_m, dT, below, above, inside = _bands(T, params)
    plateau = (c_p_s + c_p_l) / 2 + delta_H_melt / (2 * dT)
    return c_p_s * below + c_p_l * above + plateau * inside


def get_rho_Na_i(T, params, rho_l, rho_s):
    T_m, dT, below, above, inside = _bands(T, params)
    return rho_s * below + rho_l * above + _interp_linear(T, T_m - dT, T_m + dT, rho_s, rho_l) * inside


# --------------------------------------------------------------------------- wick (reference :181-207)
def _wick_consts(params):
    return dict(porosity=params['porosity_wick'], T_melt=params['T_melting_sodium'], dT_melt=params['delta_T_sodium'])


def k_eff_wick(T, Na_props, steel_props, params):
    r = _device_eval('wick_k', T, _wick_consts(params))
    if r is not None:
        return r
    k_na = get_k_Na_i(T, params, Na_props['thermal_conductivity_liquid'](T), Na_props['thermal_conductivity_solid'](T))
    k_st = steel_props['thermal_conductivity'](T)
    eps = params['porosity_wick']
    return k_na * ((k_na + k_st) - (1 - eps) * (k_na - k_st)) / ((k_na + k_st) + (1 - eps) * (k_na - k_st))


def c_p_eff_wick(T, Na_props, steel_props, params):
    r = _device_eval('wick_cp', T, _wick_consts(params))
    if r is not None:
        return r
    cp_na = get_c_p_Na_i(T, Na_props, params, Na_props['specific_heat_liquid'](T), Na_props['specific_heat_solid'](T))
    eps = params['porosity_wick']
    return eps * cp_na + (1 - eps) * steel_props['specific_heat'](T)


def rho_eff_wick(T, Na_props, steel_props, params):
    r = _device_eval('wick_rho', T, _wick_consts(params))
    if r is not None:
        return r
    rho_na = get_rho_Na_i(T, params, Na_props['density_liquid'](T), Na_props['density_solid'](T))
    eps = params['porosity_wick']
    return eps * rho_na + (1 - eps) * steel_props['density_cond'](T)


def get_wick_properties():
    return {
        'thermal_conductivity': k_eff_wick,
        'specific_heat': c_p_eff_wick,
        'density': rho_eff_wick,
    }


# --------------------------------------------------------------------------- vapor core (reference :140-154,173-178)
def k_eff_vc(T, mesh, region, sodium_props, dims, params, consts):
    """Kinetic-theory effective conductivity of the vapor core; ``region`` is 'evap_cond' or anything else
    (= adiabatic variant).  ``mesh`` is accepted and ignored like in the reference's live code path
    (the sonic-limit block after the ``return`` at reference :154 is dead code)."""
    name = 'vc_k_evap_cond' if region == 'evap_cond' else 'vc_k_adiabatic'
    r = _device_eval(name, T, dict(R_vc=dims['R_vc'], M_g=params['M_g_sodium'], R_gas=consts['R']))
    if r is not None:
        return r
    xp = _xp(T)
    R_v, R, M = dims['R_vc'], consts['R'], params['M_g_sodium']
    P = sodium_props['vapor_pressure'](T)
    h_vap = sodium_props['heat_of_vaporization'](T)
    evap = region == 'evap_cond'
    enthalpy = sodium_props['enthalpy_liquid'](T) if evap else sodium_props['enthalpy_vapor'](T)
    slip = (2 * R_v / 3) * xp.sqrt((8 * R * T) / (math.pi * M))
    base = (R_v ** 2 * P) / (4 if evap else 8) / sodium_props['viscosity'](T) + slip
    return base * (enthalpy * h_vap * M ** 2 * P) / (R ** 2 * T ** 3)


def get_vc_properties():
    return {
        'thermal_conductivity': k_eff_vc,
        'specific_heat': c_p_sodium_v,
        'density': rho_sodium_v,
    }
