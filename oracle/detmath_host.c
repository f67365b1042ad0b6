/*
 * TEST INFRASTRUCTURE — host build of include/gga_detmath.h so that tests can
 * compare (a) the deterministic routine with libm and (b) the device build of the
 * same header with this host build, bit for bit.
 */
#include "../include/gga_detmath.h"

void gga_oracle_det_sincos(const float* x, long n, float* sn, float* cs) {
  for (long i = 0; i < n; ++i) {
    double s, c;
    gga_sincos_f32(x[i], &s, &c);
    sn[i] = (float)s;
    cs[i] = (float)c;
  }
}
