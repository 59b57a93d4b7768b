/* Host restatement of the device exp (abx_core.cuh: exp_ni), instruction for instruction (every fp64 step an explicit fma / add / mul), checked
 * against long-double expl and glibc's exp on 2e8 points:  gcc -O2 -mfma -I marl_optimal_execution_b200/csrc -o /tmp/check_exp tools/check_exp.c -lm && /tmp/check_exp
 * Prints the maximum error in ulp and how often the result differs from glibc's.  With -DDUMP n it writes n (x, y) pairs to stdout as raw doubles
 * (tests/test_gpu_philox.py compares the device results with such a dump bit for bit). */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "abx_exp_table.h"
static const double T[256] = { ABX_EXP_TABLE };
static inline uint64_t bits(double d) { uint64_t u; memcpy(&u, &d, 8); return u; }
static inline double dbl(uint64_t u) { double d; memcpy(&d, &u, 8); return d; }
double exp_dev(double x) {
  int far = !(fabs(x) < 700.0);
  if (far) { if (x != x) return x; x = x < -800.0 ? -800.0 : (x > 800.0 ? 800.0 : x); }
  double t = __builtin_fma(x, ABX_EXP_INV_LN2N, 0x1.8p52);
  int32_t k = (int32_t)(uint32_t)bits(t);
  double kd = t - 0x1.8p52;
  double r = __builtin_fma(kd, -ABX_EXP_LN2N_HI, x); r = __builtin_fma(kd, -ABX_EXP_LN2N_LO, r);
  double h = T[2 * (k & 127)], tl = T[2 * (k & 127) + 1];
  double r2 = r * r;
  double p = __builtin_fma(__builtin_fma(1.0 / 120, r, 1.0 / 24), r2, __builtin_fma(1.0 / 6, r, 0.5));
  double q = __builtin_fma(p, r2, tl + r);
  double y = __builtin_fma(h, q, h);
  int e = k >> 7;
  if (far) { int e1 = e / 2; volatile double y1 = y * dbl((uint64_t)(1023 + e1) << 52); return y1 * dbl((uint64_t)(1023 + e - e1) << 52); }
  return dbl(bits(y) + ((uint64_t)(int64_t)e << 52));
}
int main(int argc, char **argv) {
  srand48(12345); double maxulp = 0; long bad = 0, n = argc > 1 ? atol(argv[1]) : 200000000;
  for (long i = 0; i < n; i++) {
    double x = (i & 1) ? -60.0 * drand48() : (drand48() - 0.8) * 40.0; if ((i & 15) == 3) x = -1e-7 * drand48(); if ((i & 1023) == 5) x = (drand48() - 0.5) * 1390.0; if ((i & 1023) == 7) x = (drand48() - 0.5) * 1500.0;
    double a = exp_dev(x); long double ref = expl((long double)x); double ra = (double)ref;
    if (isfinite(ra) && ra >= 0x1p-1022) { long double ulp = (long double)(dbl(bits(ra) + 1) - ra), err = fabsl((long double)a - ref) / ulp; if (err > maxulp) maxulp = err; }
    else if (fabs(a - exp(x)) > 0x1p-1074) { printf("far mismatch x=%a got %a libm %a\n", x, a, exp(x)); return 1; }
    if (a != exp(x) && !(a != a && x != x)) bad++;
  }
  printf("max error %.4f ulp; differs from glibc exp on %ld of %ld points (%.3g)\n", maxulp, bad, n, (double)bad / n);
  return 0;
}
