// Host build of photic_b200/csrc/exact_math.cuh, compared against the live libm (tests only).
// g++ -O2 -mfma -ffp-contract=off -shared -fPIC
#include <math.h>
#include <stdint.h>
#include "../../photic_b200/csrc/exact_math.cuh"

static const uint64_t g_exp_tab[2 * PHM_N] = PHM_EXP_TAB;
static const double g_log_tab[2 * PHM_N] = PHM_LOG_TAB;
static const double g_pow_tab[3 * PHM_N] = PHM_POWLOG_TAB;

static inline bool same(double a, double b) {
  uint64_t x = phm::bits(a), y = phm::bits(b);
  if (x == y) return true;
  return (a != a) && (b != b);  // any NaN == any NaN
}

extern "C" {
// each returns the number of mismatching results; first mismatch index in *first (or -1)
long long phm_check_exp(const double *x, long long n, long long *first) {
  long long bad = 0; *first = -1;
  for (long long i = 0; i < n; i++) {
    volatile double xi = x[i];
    if (!same(phm::exp(xi, g_exp_tab), ::exp(xi))) { if (!bad) *first = i; bad++; }
  }
  return bad;
}
long long phm_check_log(const double *x, long long n, long long *first) {
  long long bad = 0; *first = -1;
  for (long long i = 0; i < n; i++) {
    volatile double xi = x[i];
    if (!same(phm::log(xi, g_log_tab), ::log(xi))) { if (!bad) *first = i; bad++; }
  }
  return bad;
}
long long phm_check_pow(const double *x, const double *y, long long n, long long *first) {
  phm::Tables tb = {g_exp_tab, g_log_tab, g_pow_tab};
  long long bad = 0; *first = -1;
  for (long long i = 0; i < n; i++) {
    volatile double xi = x[i], yi = y[i];
    if (!same(phm::pow(xi, yi, tb), ::pow(xi, yi))) { if (!bad) *first = i; bad++; }
  }
  return bad;
}
long long phm_check_log10(const double *x, long long n, long long *first) {
  long long bad = 0; *first = -1;
  for (long long i = 0; i < n; i++) {
    volatile double xi = x[i];
    if (!same(phm::log10(xi, g_log_tab), ::log10(xi))) { if (!bad) *first = i; bad++; }
  }
  return bad;
}
double phm_host_exp(double x) { return phm::exp(x, g_exp_tab); }
double phm_host_log(double x) { return phm::log(x, g_log_tab); }
double phm_host_pow(double x, double y) { phm::Tables tb = {g_exp_tab, g_log_tab, g_pow_tab}; return phm::pow(x, y, tb); }
}
