/* Groundwork for a division-free ZeroToOne entry (cfg 2 note).  NOT wired into the
 * product: the device still uses __ddiv_rn.  Exhaustive check, on the CPU, that over the domain the
 * scoring kernel can meet
 *     zs = 1 - |explen - qlen| / explen      1 <= qlen <= explen <= 255   (zero_to_one.rs:72)
 *     v  = (zs / tf) * tf                    1 <= tf <= 64                (zero_to_one.rs:117-119; zs/tf <= 1)
 *     e  = v / m                             1 <= m = max(field_length, query_terms_len) <= 1024
 * both divisions can be replaced by  y = RN(1/d) from a table,  q = RN(x*y),  r = fma(-q, d, x) (exact),
 * q' = fma(r, y, q)  with q' == RN(x/d) BIT FOR BIT (Markstein's correction step; its hypotheses are not
 * assumed here, every case is compared with the real division).
 *   gcc -O2 -fopenmp -ffp-contract=off scripts/prove_z2o_rcp.c -lm -o /tmp/prove && /tmp/prove          */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

static inline double div_rcp(double x, double d, double y) {
  double q = x * y;
  double r = fma(-q, d, x);
  return fma(r, y, q);
}
static inline int same(double a, double b) { return memcmp(&a, &b, 8) == 0; }

int main(void) {
  static double rcp[1025];
  for (int d = 1; d <= 1024; ++d) rcp[d] = 1.0 / (double)d;
  unsigned long long checked = 0, bad_tf = 0, bad_m = 0;
#pragma omp parallel for schedule(dynamic) reduction(+ : checked, bad_tf, bad_m)
  for (int explen = 1; explen <= 255; ++explen) {
    for (int qlen = 1; qlen <= explen; ++qlen) {
      const double e = (double)explen, q = (double)qlen;
      const double zs = 1.0 - fabs(e - q) / e;
      for (int tf = 1; tf <= 64; ++tf) {
        const double t = (double)tf;
        const double a = zs / t;
        if (!same(a, div_rcp(zs, t, rcp[tf]))) ++bad_tf;
        const double v = fmin(a, 1.0) * t;
        for (int m = 1; m <= 1024; ++m) {
          ++checked;
          if (!same(v / (double)m, div_rcp(v, (double)m, rcp[m]))) ++bad_m;
        }
      }
    }
  }
  printf("checked %llu (zs, tf, m) triples: %llu mismatches in zs/tf, %llu mismatches in v/m\n", checked, bad_tf, bad_m);
  return (bad_tf || bad_m) ? 1 : 0;
}
