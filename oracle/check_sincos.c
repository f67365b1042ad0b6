/*
 * TEST INFRASTRUCTURE (oracle side) — not part of the product path.
 *
 * Exhaustive comparison of include/gga_detmath.h with the host libm over every
 * fp32 bit pattern:  (float)gga_sin((double)x) == (float)sin((double)x)  and the
 * same for cos.  That is the per-box quantity of the membership contract
 * (SURVEY.md Appendix A.1).  Usage:
 *     check_sincos [first_hi_byte last_hi_byte]      (default 0 255 = all 2^32)
 * Prints every mismatching input (bit pattern) and a summary line.
 * Build: gcc -O2 -ffp-contract=off -fopenmp check_sincos.c -lm
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/gga_detmath.h"

static inline uint32_t fbits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

int main(int argc, char** argv) {
  unsigned lo = 0, hi = 255;
  if (argc >= 3) { lo = (unsigned)atoi(argv[1]); hi = (unsigned)atoi(argv[2]); }
  unsigned long long bad_s = 0, bad_c = 0, total = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : bad_s, bad_c, total)
  for (unsigned top = lo; top <= hi; ++top) {
    for (uint32_t low = 0; low < (1u << 24); ++low) {
      const uint32_t b = (top << 24) | low;
      float x; memcpy(&x, &b, 4);
      double s, c;
      gga_sincos_f32(x, &s, &c);
      const float fs = (float)s, fc = (float)c;
      const float gs = (float)sin((double)x), gc = (float)cos((double)x);
      ++total;
      const int nan_in = (x != x) || isinf(x);
      if (nan_in) {
        if (!(fs != fs) || !(fc != fc)) {
#pragma omp critical
          printf("NAN-HANDLING 0x%08x\n", b);
          ++bad_s;
        }
        continue;
      }
      if (fbits(fs) != fbits(gs)) {
        ++bad_s;
#pragma omp critical
        printf("SIN 0x%08x x=%.9g ours=%.17g (0x%08x) libm=%.17g (0x%08x)\n", b, x, s, fbits(fs),
               sin((double)x), fbits(gs));
      }
      if (fbits(fc) != fbits(gc)) {
        ++bad_c;
#pragma omp critical
        printf("COS 0x%08x x=%.9g ours=%.17g (0x%08x) libm=%.17g (0x%08x)\n", b, x, c, fbits(fc),
               cos((double)x), fbits(gc));
      }
    }
  }
  printf("checked=%llu sin_mismatch=%llu cos_mismatch=%llu\n", total, bad_s, bad_c);
  return (bad_s || bad_c) ? 1 : 0;
}
