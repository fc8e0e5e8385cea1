// Host build of csrc/coflux_fastmath.cuh for tests/test_fastmath.py (test infrastructure).
#include "../climaocean.jl_b200/csrc/coflux_fastmath.cuh"
using namespace coflux;
extern "C" {
void fm_rcp(const double* x, double* y, long n) { for (long i = 0; i < n; ++i) y[i] = fm::rcp(x[i]); }
void fm_div(const double* a, const double* b, double* y, long n) { for (long i = 0; i < n; ++i) y[i] = fm::div(a[i], b[i]); }
void fm_sqrt(const double* x, double* y, long n) { for (long i = 0; i < n; ++i) y[i] = fm::sqrt(x[i]); }
void fm_cbrt(const double* x, double* y, long n) { for (long i = 0; i < n; ++i) y[i] = fm::cbrt(x[i]); }
void fm_log(const double* x, double* y, long n) { for (long i = 0; i < n; ++i) y[i] = fm::log(x[i], &COFLUX_LOG_TABLE[0][0]); }
void fm_exp(const double* x, double* y, long n) { for (long i = 0; i < n; ++i) y[i] = fm::exp(x[i], COFLUX_EXP_TABLE); }
}
