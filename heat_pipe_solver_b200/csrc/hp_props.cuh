// Material laws of the heat-pipe conduction path, fp64, usable from device AND host code.
//
// Restates /root/reference/src/2d/material_properties.py (line numbers cited per function)
// and the region composites of /root/reference/src/2d/main.py:103-117.  The header is
// `__host__ __device__` so that tests/hostcheck can compile the very same code with g++
// and compare it with the golden vectors produced by the reference's own module
// (tests/golden/reference_golden.npz) without a GPU.  The product only ever calls these
// from CUDA kernels.
#pragma once
#include <math.h>
#include <stdint.h>
#include "hp_solver.h"

#if defined(__CUDACC__)
#define HP_HD __host__ __device__ __forceinline__
#else
#define HP_HD inline
#endif

// region codes (host classifies with the reference's float recipe, main.py:64-72,86-90)
enum { HP_BAND_VC = 0, HP_BAND_WICK = 1, HP_BAND_WALL = 2, HP_BAND_NONE = 3 };
// z flags at a cell centre / face centre
enum {
    HP_ZF_VC_EVAPCOND = 1,   // z < L_e or z > L_e+L_a          main.py:65,87
    HP_ZF_VC_ADIAB = 2,      // L_e < z < L_e+L_a               main.py:66,88
    HP_ZF_BEFORE_COND = 4,   // z < L_e+L_a                     main.py:68,71
    HP_ZF_IN_COND = 8,       // z > L_e+L_a                     main.py:69,72
    HP_ZF_BC_EVAP = 16,      // L_input_left < z < L_input_right main.py:123
    HP_ZF_BC_COND = 32       // L_e+L_a < z < L_total            main.py:124
};

typedef hp_phys HpPhys;   // constants + mode switches: see include/hp_solver.h

namespace hp {

constexpr double T_CRIT_NA = 2509.46;   // material_properties.py:5
constexpr double DH_MELT_NA = 113e3;    // material_properties.py:6

// ---- sodium ----------------------------------------------------------------
HP_HD double na_rho_s(double T) { return 972.70 - 0.2154 * (T - 273.15); }                       // :9
HP_HD double na_rho_l(double T) {                                                                  // :10
    double s = 1.0 - T / T_CRIT_NA;
    return 219.0 + 275.32 * s + 511.58 * sqrt(s);
}
HP_HD double na_h_vap(double T) {                                                                  // :12-13
    double s = 1.0 - T / T_CRIT_NA;
    return (393.37 * s + 4398.6 * pow(s, 0.29302)) * 1e3;
}
HP_HD double na_psat(double T) { return exp(11.9463 - 12633.73 / T - 0.4672 * log(T)) * 1e6; }    // :56
HP_HD double na_dpdt(double T) {                                                                   // :15-16
    return (12633.73 / (T * T) - 0.4672 / T) * exp(11.9463 - 12633.73 / T - 0.4672 * log(T)) * 1e6;
}
HP_HD double na_rho_v(double T) {                                                                  // :18-19
    return 1.0 / (na_h_vap(T) / (T * na_dpdt(T)) + 1.0 / na_rho_l(T));
}
HP_HD double na_k_s(double T) { return 135.6 - 0.167 * (T - 273.15); }                            // :21
HP_HD double na_k_l(double T) {                                                                    // :22
    return 124.67 - 0.11381 * T + 5.5226e-5 * (T * T) - 1.1842e-8 * (T * T * T);
}
HP_HD double na_cp_s(double T) {                                                                   // :24
    double c = T - 273.15;
    return 1199.0 + 0.649 * c + 1052.9e-5 * (c * c);
}
HP_HD double na_cp_l(double T) {                                                                   // :25
    double c = T - 273.15;
    return 1436.72 - 0.58 * c + 4.672e-4 * (c * c);
}
HP_HD double na_cp_v(double T) {                                                                   // :27-29
    double T2 = T * T, T3 = T2 * T, T4 = T2 * T2, T5 = T4 * T, T6 = T3 * T3;
    return 16105.174483770606 - 111.33277032117233 * T + 0.2994911138258808 * T2
           - 0.00037555960577985903 * T3 + 2.418799137943085e-07 * T4
           - 7.773570963102343e-11 * T5 + 9.892389871569863e-15 * T6;
}
HP_HD double na_cv_v(double T) {                                                                   // :51-53
    double T2 = T * T, T3 = T2 * T, T4 = T2 * T2, T5 = T4 * T, T6 = T3 * T3;
    return 2944.6696506593726 - 27.379453000849395 * T + 0.08934864074702555 * T2
           - 0.00011764446335597157 * T3 + 7.521494832758196e-08 * T4
           - 2.338468836342328e-11 * T5 + 2.8403280414122426e-15 * T6;
}
HP_HD double na_mu_v(double T) { return 6.083e-9 * T + 1.2606e-5; }                               // :55
HP_HD double na_h_l(double T) {                                                                    // :58-59
    return (-365.77 + 1.6582 * T - 4.2395e-4 * (T * T) + 1.4847e-7 * (T * T * T) + 2992.6 / T) * 1e3;
}
HP_HD double na_h_v(double T) { return na_h_l(T) + na_h_vap(T); }                                 // :61

// ---- steel -----------------------------------------------------------------
HP_HD double steel_rho(double T) { return (7.9841 - 2.656e-4 * T - 1.158e-7 * (T * T)) * 1e3; }   // :89
HP_HD double steel_k(double T) { return (8.116e-2 + 1.618e-4 * T) * 100.0; }                      // :90
HP_HD double steel_cp(double T) { return 0.1348 * T + 469.47; }                                   // :91
HP_HD double steel_rho_eff(double T) { return steel_rho(T) + 3.75e6 / steel_cp(T); }              // :93

// ---- sodium across the melting smear T_m +- dT  (:104-120) --------------------
HP_HD double lin_interp(double T, double Tlo, double Thi, double vlo, double vhi) {
    return vlo + (vhi - vlo) * ((T - Tlo) / (Thi - Tlo));
}
HP_HD double na_k_i(double T, double Tm, double dT) {                                              // :107-109
    double lo = Tm - dT, hi = Tm + dT;
    double ks = na_k_s(T), kl = na_k_l(T);
    if (T < lo) return ks;
    if (T > hi) return kl;
    if (T >= lo && T <= hi) return lin_interp(T, lo, hi, ks, kl);
    return 0.0;   // NaN input: every mask false
}
HP_HD double na_cp_i(double T, double Tm, double dT) {                                             // :111-115
    double lo = Tm - dT, hi = Tm + dT;
    double cs = na_cp_s(T), cl = na_cp_l(T);
    if (T < lo) return cs;
    if (T > hi) return cl;
    if (T >= lo && T <= hi) return (cs + cl) / 2.0 + DH_MELT_NA / (2.0 * dT);
    return 0.0;
}
HP_HD double na_rho_i(double T, double Tm, double dT) {                                            // :117-120
    double lo = Tm - dT, hi = Tm + dT;
    double rs = na_rho_s(T), rl = na_rho_l(T);
    if (T < lo) return rs;
    if (T > hi) return rl;
    if (T >= lo && T <= hi) return lin_interp(T, lo, hi, rs, rl);
    return 0.0;
}

// ---- wick ------------------------------------------------------------------
HP_HD double wick_k(double T, double eps, double Tm, double dT) {                                  // :181-193
    double kna = na_k_i(T, Tm, dT);
    double kst = steel_k(T);
    double num = (kna + kst) - (1.0 - eps) * (kna - kst);
    double den = (kna + kst) + (1.0 - eps) * (kna - kst);
    return kna * num / den;
}
HP_HD double wick_cp(double T, double eps, double Tm, double dT) {                                 // :195-200
    return eps * na_cp_i(T, Tm, dT) + (1.0 - eps) * steel_cp(T);
}
HP_HD double wick_rho(double T, double eps, double Tm, double dT) {                                // :202-207
    return eps * na_rho_i(T, Tm, dT) + (1.0 - eps) * steel_rho(T);
}

// ---- vapor core ------------------------------------------------------------
// :140-154 (live part).  adiabatic != 0 selects the `/8` + h_v variant (:150-151).
HP_HD double vc_k(double T, int adiabatic, double R_v, double M, double Rg) {
    const double PI = 3.141592653589793;
    double mu = na_mu_v(T);
    double P = na_psat(T);
    double hvap = na_h_vap(T);
    double hl = na_h_l(T);
    double enth = adiabatic ? (hl + hvap) : hl;
    double factor = (2.0 * R_v / 3.0) * sqrt((8.0 * Rg * T) / (PI * M));
    double base = (R_v * R_v * P) / (adiabatic ? 8.0 : 4.0) / mu + factor;
    return base * (enth * hvap * (M * M) * P) / ((Rg * Rg) * (T * T * T));
}

// ---- region composites (main.py:103-117) -------------------------------------
HP_HD double k_region(const HpPhys& ph, int band, unsigned zf, double T) {
    if (ph.properties == 1) return ph.simple_k;
    if (band == HP_BAND_WALL) return steel_k(T);
    if (band == HP_BAND_WICK) return wick_k(T, ph.porosity, ph.T_melt, ph.dT_melt);
    if (band == HP_BAND_VC) {
        double k = 0.0;
        if (zf & HP_ZF_VC_EVAPCOND) k += vc_k(T, 0, ph.R_vc, ph.M_g, ph.R_gas);
        if (zf & HP_ZF_VC_ADIAB) k += vc_k(T, 1, ph.R_vc, ph.M_g, ph.R_gas);
        return k;
    }
    return 0.0;
}

HP_HD void rho_cp_region(const HpPhys& ph, int band, unsigned zf, double T, double& rho, double& cp) {
    if (ph.properties == 1) { rho = ph.simple_rho; cp = ph.simple_cp; return; }
    rho = 0.0; cp = 0.0;
    if (band == HP_BAND_WALL) {
        cp = steel_cp(T);
        if (zf & HP_ZF_BEFORE_COND) rho += steel_rho_eff(T);      // main.py:116
        if (zf & HP_ZF_IN_COND) rho += steel_rho(T);              // main.py:117
    } else if (band == HP_BAND_WICK) {
        cp = wick_cp(T, ph.porosity, ph.T_melt, ph.dT_melt);
        rho = wick_rho(T, ph.porosity, ph.T_melt, ph.dT_melt);
    } else if (band == HP_BAND_VC) {
        double w = ((zf & HP_ZF_VC_EVAPCOND) ? 1.0 : 0.0) + ((zf & HP_ZF_VC_ADIAB) ? 1.0 : 0.0);
        cp = na_cp_v(T) * w;                                     // main.py:108-109
        rho = na_rho_v(T) * w;                                   // main.py:113-114
    }
}

// Property selector for the elementwise evaluation entry point (hp_eval_property).
enum PropId {
    PROP_NA_RHO_S = 0, PROP_NA_RHO_L, PROP_NA_RHO_V, PROP_NA_HVAP, PROP_NA_K_S, PROP_NA_K_L,
    PROP_NA_CP_S, PROP_NA_CP_L, PROP_NA_CP_V, PROP_NA_CV_V, PROP_NA_H_L, PROP_NA_H_V, PROP_NA_MU_V,
    PROP_NA_PSAT, PROP_STEEL_RHO_EFF, PROP_STEEL_RHO, PROP_STEEL_K, PROP_STEEL_CP,
    PROP_WICK_K, PROP_WICK_CP, PROP_WICK_RHO, PROP_VC_K_EVAPCOND, PROP_VC_K_ADIAB,
    PROP_NA_K_I, PROP_NA_CP_I, PROP_NA_RHO_I, PROP_NA_DPDT, PROP_COUNT
};

HP_HD double eval_property(const HpPhys& ph, int id, double T) {
    switch (id) {
        case PROP_NA_RHO_S: return na_rho_s(T);
        case PROP_NA_RHO_L: return na_rho_l(T);
        case PROP_NA_RHO_V: return na_rho_v(T);
        case PROP_NA_HVAP: return na_h_vap(T);
        case PROP_NA_K_S: return na_k_s(T);
        case PROP_NA_K_L: return na_k_l(T);
        case PROP_NA_CP_S: return na_cp_s(T);
        case PROP_NA_CP_L: return na_cp_l(T);
        case PROP_NA_CP_V: return na_cp_v(T);
        case PROP_NA_CV_V: return na_cv_v(T);
        case PROP_NA_H_L: return na_h_l(T);
        case PROP_NA_H_V: return na_h_v(T);
        case PROP_NA_MU_V: return na_mu_v(T);
        case PROP_NA_PSAT: return na_psat(T);
        case PROP_STEEL_RHO_EFF: return steel_rho_eff(T);
        case PROP_STEEL_RHO: return steel_rho(T);
        case PROP_STEEL_K: return steel_k(T);
        case PROP_STEEL_CP: return steel_cp(T);
        case PROP_WICK_K: return wick_k(T, ph.porosity, ph.T_melt, ph.dT_melt);
        case PROP_WICK_CP: return wick_cp(T, ph.porosity, ph.T_melt, ph.dT_melt);
        case PROP_WICK_RHO: return wick_rho(T, ph.porosity, ph.T_melt, ph.dT_melt);
        case PROP_VC_K_EVAPCOND: return vc_k(T, 0, ph.R_vc, ph.M_g, ph.R_gas);
        case PROP_VC_K_ADIAB: return vc_k(T, 1, ph.R_vc, ph.M_g, ph.R_gas);
        case PROP_NA_K_I: return na_k_i(T, ph.T_melt, ph.dT_melt);
        case PROP_NA_CP_I: return na_cp_i(T, ph.T_melt, ph.dT_melt);
        case PROP_NA_RHO_I: return na_rho_i(T, ph.T_melt, ph.dT_melt);
        case PROP_NA_DPDT: return na_dpdt(T);
        default: return 0.0;
    }
}

}  // namespace hp
