/* TEST INFRASTRUCTURE: compares the restated exp / log / pow of csrc/rvn_glibc_math.h (host build, -ffp-contract=off)
   with the running libm on pseudo-random and boundary arguments.  Built and driven by tests/test_glibc_math.py. */
#include <math.h>
#include <stdint.h>
#include <string.h>
#include "rvn_glibc_math.h"

static uint64_t s_[2];
static uint64_t next_(void) {   /* xorshift128+ */
  uint64_t a = s_[0], b = s_[1];
  s_[0] = b;
  a ^= a << 23;
  s_[1] = a ^ b ^ (a >> 18) ^ (b >> 5);
  return s_[1] + b;
}
static double u01_(void) { return (double)(next_() >> 11) * 0x1p-53; }
static int same_(double a, double b) {
  uint64_t x, y;
  memcpy(&x, &a, 8);
  memcpy(&y, &b, 8);
  if (a != a && b != b) return 1;   /* any NaN equals any NaN */
  return x == y;
}

/* mode: 0 exp, 1 log, 2 pow, 3 expm1, 4 tanh.  dist: 0 = raw bit patterns, 1 = hydrological ranges, 2 = near the special-case boundaries.
   Returns the number of mismatches; the first `cap` offending arguments go to bad[2*i], bad[2*i+1]. */
long rg_check(int mode, int dist, long n, uint64_t seed, double *bad, int cap) {
  long nbad = 0;
  s_[0] = seed * 0x9E3779B97F4A7C15ULL + 1;
  s_[1] = seed ^ 0xD1B54A32D192ED03ULL;
  for (int w = 0; w < 16; w++) next_();
  for (long i = 0; i < n; i++) {
    double x, y = 0, got, ref;
    if (dist == 0) {
      uint64_t a = next_(), b = next_();
      memcpy(&x, &a, 8);
      memcpy(&y, &b, 8);
      if (mode == 2 && (i & 1)) {   /* keep half of the pow samples in the range where the result is finite */
        x = fabs(x);
        if (x != x || x == 0 || x > 1e300) x = 1.5;
        double lx = log(x);
        if (lx == 0) lx = 1;
        y = (u01_() * 1400 - 700) / lx;
      }
    } else if (dist == 1) {
      if (mode == 0) x = (u01_() - 0.5) * (i % 3 == 0 ? 4.0 : i % 3 == 1 ? 60.0 : 1400.0);
      else if (mode == 1) x = (i % 3 == 0) ? u01_() * 2.0 : (i % 3 == 1) ? exp((u01_() - 0.5) * 80.0) : 1.0 + (u01_() - 0.5) * 0.25;
      else {
        x = (i % 4 == 0) ? u01_() : (i % 4 == 1) ? u01_() * 50.0 : (i % 4 == 2) ? exp((u01_() - 0.5) * 40.0) : 1.0 + (u01_() - 0.5) * 0.1;
        y = (i % 5 == 0) ? u01_() * 6.0 : (i % 5 == 1) ? -u01_() * 4.0 : (i % 5 == 2) ? (double)(int)(u01_() * 9 - 4) * 0.25 : (i % 5 == 3) ? (u01_() - 0.5) * 40 : u01_();
      }
    } else {
      static const double edge[] = {0.0, -0.0, 1.0, -1.0, 2.0, -2.0, 0.5, 3.0, -3.0, 1e-320, -1e-320, 0x1p-1022, 0x1p-1074, 1e308, -1e308, 1.0 / 0.0, -1.0 / 0.0, 0.0 / 0.0,
                                    709.78, 709.79, -745.1, -745.2, -708.3, -708.5, 1024.0, -1075.0, 0x1p-54, 0x1p-55, 0x1p-65, 0x1p-66, 0x1p63, 0x1p64,
                                    0.9375, 0.93749999999999989, 1.064697265625, 1.0646972656249998, 1.0000000000000002, 0.99999999999999989, 4503599627370496.0, 4503599627370497.0, 9007199254740993.0};
      const int ne = (int)(sizeof edge / sizeof edge[0]);
      x = edge[next_() % ne];
      y = edge[next_() % ne];
      if (next_() & 1) { uint64_t a; memcpy(&a, &x, 8); a += (next_() % 5) - 2; memcpy(&x, &a, 8); }
      if (next_() & 1) { uint64_t a; memcpy(&a, &y, 8); a += (next_() % 5) - 2; memcpy(&y, &a, 8); }
      if (mode == 2 && (next_() & 3) == 0) { x = 1.0 + (u01_() - 0.5) * 1e-3; y = (u01_() - 0.5) * 3e6; }   /* results near over/underflow */
      if (mode == 2 && (next_() & 3) == 0) { x = u01_() * 4; y = -1074.0 / log2(x) * (1 + (u01_() - 0.5) * 0.1); }
      if (mode == 0 && (next_() & 1)) x = (next_() & 1) ? -708.0 - u01_() * 40.0 : 709.0 + u01_() * 1.0;
    }
    if (mode >= 3) {   /* expm1 / tanh: arguments of every branch */
      if (dist == 1) x = (u01_() - 0.5) * (i % 4 == 0 ? 0.7 : i % 4 == 1 ? 2.2 : i % 4 == 2 ? 50.0 : 1500.0);
      if (dist == 2 && (next_() & 1)) x = (u01_() - 0.5) * 4.2;
      if (mode == 3) { got = rvn_expm1(x); ref = expm1(x); } else { got = rvn_tanh(x); ref = tanh(x); }
    } else
    if (mode == 0) { got = rvn_exp(x); ref = exp(x); }
    else if (mode == 1) { got = rvn_log(x); ref = log(x); }
    else { got = rvn_pow(x, y); ref = pow(x, y); }
    if (!same_(got, ref)) {
      if (nbad < cap) { bad[2 * nbad] = x; bad[2 * nbad + 1] = y; }
      nbad++;
    }
  }
  return nbad;
}

/* element-wise evaluation through the HOST libm (the thing the device functions must reproduce); used by the GPU test */
void rg_libm_eval(int mode, long n, const double *x, const double *y, double *out) {
  for (long i = 0; i < n; i++) out[i] = mode == 0 ? exp(x[i]) : mode == 1 ? log(x[i]) : mode == 2 ? pow(x[i], y[i]) : mode == 3 ? expm1(x[i]) : tanh(x[i]);
}
void rg_restated_eval(int mode, long n, const double *x, const double *y, double *out) {
  for (long i = 0; i < n; i++) out[i] = mode == 0 ? rvn_exp(x[i]) : mode == 1 ? rvn_log(x[i]) : mode == 2 ? rvn_pow(x[i], y[i]) : mode == 3 ? rvn_expm1(x[i]) : rvn_tanh(x[i]);
}

/* x / y against the reciprocal-multiply + one Markstein correction the sweep uses for divisions by the model time step
   (csrc/rvn_kernels.cuh div_by): number of mismatches over n pseudo-random x spanning `decades` decades around 1 */
long rg_check_div(double y, long n, uint64_t seed, int decades, double *bad, int cap) {
  long nbad = 0;
  s_[0] = seed * 0x9E3779B97F4A7C15ULL + 7; s_[1] = seed ^ 0xA24BAED4963EE407ULL;
  for (int w = 0; w < 16; w++) next_();
  const double yr = 1.0 / y;
  for (long i = 0; i < n; i++) {
    double x = (u01_() - 0.3) * pow(10.0, (u01_() - 0.5) * 2.0 * decades);
    if (i % 97 == 0) x = 0.0;
    const double q = x * yr, r = __builtin_fma(-y, q, x), got = __builtin_fma(r, yr, q), ref = x / y;
    if (!same_(got, ref)) { if (nbad < cap) { bad[2 * nbad] = x; bad[2 * nbad + 1] = y; } nbad++; }
  }
  return nbad;
}
