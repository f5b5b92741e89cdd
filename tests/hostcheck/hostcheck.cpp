// TEST INFRASTRUCTURE ONLY.  Compiles the product's device header (csrc/hp_props.cuh) as plain host
// C++ so that the CPU test-suite can check the very functions the CUDA kernels inline against the golden
// vectors produced by the reference's own material_properties.py.  Never linked into libhpsolver.so.
#include "hp_props.cuh"

extern "C" {
void hc_eval_property(const hp_phys* ph, int id, const double* T, double* out, long long n) {
    for (long long i = 0; i < n; ++i) out[i] = hp::eval_property(*ph, id, T[i]);
}
void hc_k_region(const hp_phys* ph, int band, unsigned zf, const double* T, double* out, long long n) {
    for (long long i = 0; i < n; ++i) out[i] = hp::k_region(*ph, band, zf, T[i]);
}
void hc_rho_cp_region(const hp_phys* ph, int band, unsigned zf, const double* T, double* rho, double* cp, long long n) {
    for (long long i = 0; i < n; ++i) hp::rho_cp_region(*ph, band, zf, T[i], rho[i], cp[i]);
}
int hc_prop_count(void) { return hp::PROP_COUNT; }
}
