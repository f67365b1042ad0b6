/*
 * gga_detmath.h — deterministic sin/cos of an fp32 angle, bit-identical on
 * host (gcc) and device (nvcc, sm_100a).
 *
 * Why this exists: the membership contract (restating
 * the un-vendored mmcv `points_in_boxes_cpu`) needs, once per box,
 *     cosa = (float)cos((double)(-rz)),  sina = (float)sin((double)(-rz)).
 * libm on the host and the CUDA math library on the device are both "<= 1-2 ulp"
 * in double but not identical, so the float they round to can differ once in
 * ~2^28 yaws.  This routine uses only IEEE-754 +,-,* (round-to-nearest) and
 * 64-bit integer arithmetic, so the same source gives the same bits on both
 * sides; `oracle/check_sincos.c` compares it with glibc over ALL 2^32 fp32
 * inputs (result recorded in DESIGN.md).
 *
 * Algorithm (published, fdlibm/Payne-Hanek lineage; restated, not copied):
 *   |x| <= pi/4           : r = |x|, quadrant 0
 *   otherwise             : x = m * 2^e with a 24-bit integer m; multiply m by a
 *                           192-bit window of 2/pi chosen so everything above
 *                           the window is a multiple of 4; the top two bits of
 *                           the product mod 4 are the quadrant, the remaining
 *                           190 bits the fraction f; round to nearest
 *                           (f in [-1/2,1/2]); r = f * pi/2 in double-double.
 *   sin/cos on [-pi/4,pi/4]: degree-13 / degree-14 minimax polynomials with a
 *                           low-order correction term for r_lo.
 */
#ifndef GGA_DETMATH_H_
#define GGA_DETMATH_H_

#include <stdint.h>

#if defined(__CUDACC__)
#define GGA_HD __host__ __device__ __forceinline__
#else
#define GGA_HD static inline
#endif

#if defined(__CUDA_ARCH__)
#define GGA_DMUL(a, b) __dmul_rn((a), (b))
#define GGA_DADD(a, b) __dadd_rn((a), (b))
#define GGA_DSUB(a, b) __dsub_rn((a), (b))
#define GGA_MULHI64(a, b) __umul64hi((a), (b))
#define GGA_CLZ64(a) __clzll((long long)(a))
#else
/* Host: the translation unit must be built with -ffp-contract=off (the oracle
 * Makefile and the nvcc host flags both do). */
#define GGA_DMUL(a, b) ((double)(a) * (double)(b))
#define GGA_DADD(a, b) ((double)(a) + (double)(b))
#define GGA_DSUB(a, b) ((double)(a) - (double)(b))
#define GGA_MULHI64(a, b) ((uint64_t)(((unsigned __int128)(a) * (unsigned __int128)(b)) >> 64))
#define GGA_CLZ64(a) __builtin_clzll((unsigned long long)(a))
#endif

GGA_HD double gga_bits2d(uint64_t u) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)u);
#else
  union { uint64_t u; double d; } c; c.u = u; return c.d;
#endif
}

GGA_HD uint32_t gga_f2bits(float f) {
#if defined(__CUDA_ARCH__)
  return __float_as_uint(f);
#else
  union { uint32_t u; float f; } c; c.f = f; return c.u;
#endif
}

/* sin on [-pi/4, pi/4] for r = x + y (|y| << |x|). */
GGA_HD double gga_ksin(double x, double y) {
  const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03,
               S3 = -1.98412698298579493134e-04, S4 = 2.75573137070700676789e-06,
               S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
  double z = GGA_DMUL(x, x);
  double v = GGA_DMUL(z, x);
  double r = GGA_DADD(S5, GGA_DMUL(z, S6));
  r = GGA_DADD(S4, GGA_DMUL(z, r));
  r = GGA_DADD(S3, GGA_DMUL(z, r));
  r = GGA_DADD(S2, GGA_DMUL(z, r));
  /* x - ((z*(y/2 - v*r) - y) - v*S1) */
  double t = GGA_DSUB(GGA_DMUL(0.5, y), GGA_DMUL(v, r));
  t = GGA_DSUB(GGA_DMUL(z, t), y);
  t = GGA_DSUB(t, GGA_DMUL(v, S1));
  return GGA_DSUB(x, t);
}

/* cos on [-pi/4, pi/4] for r = x + y. */
GGA_HD double gga_kcos(double x, double y) {
  const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03,
               C3 = 2.48015872894767294178e-05, C4 = -2.75573143513906633035e-07,
               C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
  double z = GGA_DMUL(x, x);
  double r = GGA_DADD(C5, GGA_DMUL(z, C6));
  r = GGA_DADD(C4, GGA_DMUL(z, r));
  r = GGA_DADD(C3, GGA_DMUL(z, r));
  r = GGA_DADD(C2, GGA_DMUL(z, r));
  r = GGA_DADD(C1, GGA_DMUL(z, r));
  r = GGA_DMUL(z, r);
  double hz = GGA_DMUL(0.5, z);
  double w = GGA_DSUB(1.0, hz);
  /* w + (((1-w)-hz) + (z*r - x*y)) */
  double t = GGA_DSUB(GGA_DSUB(1.0, w), hz);
  t = GGA_DADD(t, GGA_DSUB(GGA_DMUL(z, r), GGA_DMUL(x, y)));
  return GGA_DADD(w, t);
}

GGA_HD uint64_t gga_d2bits(double d) {
#if defined(__CUDA_ARCH__)
  return (uint64_t)__double_as_longlong(d);
#else
  union { uint64_t u; double d; } c; c.d = d; return c.u;
#endif
}

/* 1 when rounding the double `d` to fp32 cannot be changed by an error of 256 double ulps,
 * i.e. the 29 dropped mantissa bits are not within 256 of the round-to-nearest midpoint. */
GGA_HD int gga_f32_rounding_is_safe(double d) {
  const uint32_t low = (uint32_t)(gga_d2bits(d) & 0x1fffffffull);
  const uint32_t dist = low > 0x10000000u ? low - 0x10000000u : 0x10000000u - low;
  return dist > 256u;
}

/* Fast path for pi/4 < |x| < 64 (every yaw a detector produces): two-constant Cody-Waite
 * reduction of |x| in double-double instead of the 192-bit Payne-Hanek product.  Both paths
 * feed the same kernels and are accurate to ~1e-15 relative, so they round to the same fp32
 * unless the result sits within 256 double ulps of an fp32 rounding midpoint (or the reduced
 * argument is tiny); those inputs (about 1e-6 of them) return 0 and take the full path.  The
 * fp32 results are therefore identical by construction; `oracle/check_sincos.c` re-checks
 * every fp32 input against libm.
 *   k = rint(|x| * 2/pi) <= 41;  r = |x| - k*P1 (exact: P1 has 33 bits);  w = k*P1T;
 *   (rh, rl) = TwoSum(r, -w). */
GGA_HD int gga_sincos_small(uint32_t abits, double* sn, double* cs) {
#if defined(__CUDA_ARCH__)
  const double ax = (double)__uint_as_float(abits);
#else
  union { uint32_t u; float f; } cv; cv.u = abits;
  const double ax = (double)cv.f;
#endif
  const double INVPIO2 = 6.36619772367581382433e-01, P1 = 1.57079632673412561417e+00,
               P1T = 6.07710050650619224932e-11, MAGIC = 6755399441055744.0; /* 1.5 * 2^52 */
  const double fn = GGA_DSUB(GGA_DADD(GGA_DMUL(ax, INVPIO2), MAGIC), MAGIC);
  const double r = GGA_DSUB(ax, GGA_DMUL(fn, P1));
  const double nw = -GGA_DMUL(fn, P1T);
  const double rh = GGA_DADD(r, nw);
  const double bb = GGA_DSUB(rh, r);
  const double rl = GGA_DADD(GGA_DSUB(r, GGA_DSUB(rh, bb)), GGA_DSUB(nw, bb));
  const double arh = rh < 0.0 ? -rh : rh;
  if (!(arh > 9.5367431640625e-07)) return 0; /* |r| <= 2^-20: cancellation, use the exact reduction */
  const double ks = gga_ksin(rh, rl), kc = gga_kcos(rh, rl);
  if (!gga_f32_rounding_is_safe(ks) || !gga_f32_rounding_is_safe(kc)) return 0;
  const uint32_t q = (uint32_t)(int32_t)fn;
  double s, c;
  switch (q & 3u) {
    case 0: s = ks; c = kc; break;
    case 1: s = kc; c = -ks; break;
    case 2: s = -ks; c = -kc; break;
    default: s = -kc; c = ks; break;
  }
  *sn = s;
  *cs = c;
  return 1;
}

/* sin and cos of an fp32 angle, evaluated in double.  Results are doubles with
 * < 1 ulp error; the membership contract rounds them to fp32. */
GGA_HD void gga_sincos_f32(float xf, double* sn, double* cs) {
  const uint32_t bits = gga_f2bits(xf);
  const uint32_t abits = bits & 0x7fffffffu;
  const int negx = (int)(bits >> 31);
  const uint32_t bexp = abits >> 23;
  if (bexp == 255u) { /* inf / nan */
    *sn = gga_bits2d(0x7ff8000000000000ull);
    *cs = gga_bits2d(0x7ff8000000000000ull);
    return;
  }
  double rh, rl;
  uint32_t q = 0;
  if (abits <= 0x3f490fd9u) { /* |x| < pi/4 */
    rh = (double)xf; /* keeps the sign; kernels are odd/even */
    rl = 0.0;
    *sn = gga_ksin(rh, rl);
    *cs = gga_kcos(rh, rl);
    return;
  }
  if (abits < 0x42800000u) { /* |x| < 64: Cody-Waite reduction, falls through when the result is too close to call */
    double fs, fc;
    if (gga_sincos_small(abits, &fs, &fc)) {
      *sn = negx ? -fs : fs;
      *cs = fc;
      return;
    }
  }
  /* 32 zero bits followed by the first 352 bits of 2/pi, most significant word first. */
  const uint64_t T0 = 0x00000000a2f9836eull, T1 = 0x4e441529fc2757d1ull,
                 T2 = 0xf534ddc0db629599ull, T3 = 0x3c439041fe5163abull,
                 T4 = 0xdebbc561b7246e3aull, T5 = 0x424dd2e006492eeaull;
  const uint64_t m = (uint64_t)((abits & 0x007fffffu) | 0x00800000u);
  const uint32_t o = bexp - 120u; /* window start bit: 6..134 */
  const uint32_t k = o >> 6, sh = o & 63u;
  uint64_t a0, a1, a2, a3;
  if (k == 0)      { a0 = T0; a1 = T1; a2 = T2; a3 = T3; }
  else if (k == 1) { a0 = T1; a1 = T2; a2 = T3; a3 = T4; }
  else             { a0 = T2; a1 = T3; a2 = T4; a3 = T5; }
  uint64_t whi, wmid, wlo;
  if (sh == 0) { whi = a0; wmid = a1; wlo = a2; }
  else {
    whi  = (a0 << sh) | (a1 >> (64u - sh));
    wmid = (a1 << sh) | (a2 >> (64u - sh));
    wlo  = (a2 << sh) | (a3 >> (64u - sh));
  }
  /* P = m * (whi:wmid:wlo), 216 bits in R3:R2:R1:R0; value mod 4 = P * 2^-190. */
  const uint64_t lo0 = m * wlo, hi0 = GGA_MULHI64(m, wlo);
  const uint64_t lo1 = m * wmid, hi1 = GGA_MULHI64(m, wmid);
  const uint64_t lo2 = m * whi;
  uint64_t r0 = lo0;
  uint64_t r1 = lo1 + hi0;
  uint64_t c1 = (r1 < lo1) ? 1ull : 0ull;
  uint64_t r2 = lo2 + hi1 + c1; /* bits >= 192 are multiples of 4: dropped */
  q = (uint32_t)(r2 >> 62);
  uint64_t f2 = r2 & 0x3fffffffffffffffull, f1 = r1, f0 = r0; /* 190-bit fraction */
  int negf = 0;
  if (f2 >> 61) { /* fraction >= 1/2: round the quadrant up, f <- f - 1 */
    negf = 1;
    q += 1u;
    f0 = ~f0 + 1ull;
    uint64_t cc = (f0 == 0ull) ? 1ull : 0ull;
    f1 = ~f1 + cc;
    cc = (cc && f1 == 0ull) ? 1ull : 0ull;
    f2 = (~f2 + cc) & 0x3fffffffffffffffull;
  }
  /* Normalise: shift the 190-bit magnitude so its top set bit sits at bit 189,
   * keep the leading 128 bits in u1:u0. */
  uint64_t n2 = (f2 << 2) | (f1 >> 62), n1 = (f1 << 2) | (f0 >> 62), n0 = f0 << 2; /* 192-bit, top-aligned */
  int lz = 0;
  if (n2 == 0ull) { n2 = n1; n1 = n0; n0 = 0ull; lz = 64; }
  if (n2 == 0ull) { n2 = n1; n1 = 0ull; lz = 128; }
  if (n2 == 0ull) { /* cannot happen (pi is irrational); keep it defined */
    rh = 0.0; rl = 0.0;
  } else {
    const int s2 = GGA_CLZ64(n2);
    uint64_t u1, u0;
    if (s2 == 0) { u1 = n2; u0 = n1; }
    else { u1 = (n2 << s2) | (n1 >> (64 - s2)); u0 = (n1 << s2) | (n0 >> (64 - s2)); }
    lz += s2;
    /* |f| = (u1*2^64 + u0) * 2^(-128-lz) */
    const double hi53 = (double)(int64_t)(u1 >> 11);
    const double mid53 = (double)(int64_t)(((u1 & 0x7ffull) << 42) | (u0 >> 22));
    const double sc_hi = gga_bits2d((uint64_t)(1023 - 53 - lz) << 52);
    const double sc_lo = gga_bits2d((uint64_t)(1023 - 106 - lz) << 52);
    const double fh = GGA_DMUL(hi53, sc_hi);  /* exact */
    const double fl = GGA_DMUL(mid53, sc_lo); /* exact */
    /* r = f * pi/2 in double-double; Dekker split instead of fma so that the
     * host needs no libm/hardware fma. */
    const double PH = gga_bits2d(0x3ff921fb54442d18ull), PL = gga_bits2d(0x3c91a62633145c07ull);
    const double SPLIT = 134217729.0; /* 2^27 + 1 */
    double t = GGA_DMUL(SPLIT, fh);
    const double fh_h = GGA_DSUB(t, GGA_DSUB(t, fh)), fh_l = GGA_DSUB(fh, fh_h);
    t = GGA_DMUL(SPLIT, PH);
    const double ph_h = GGA_DSUB(t, GGA_DSUB(t, PH)), ph_l = GGA_DSUB(PH, ph_h);
    const double p = GGA_DMUL(fh, PH);
    double e = GGA_DSUB(GGA_DMUL(fh_h, ph_h), p);
    e = GGA_DADD(e, GGA_DMUL(fh_h, ph_l));
    e = GGA_DADD(e, GGA_DMUL(fh_l, ph_h));
    e = GGA_DADD(e, GGA_DMUL(fh_l, ph_l)); /* p + e == fh*PH exactly */
    e = GGA_DADD(e, GGA_DADD(GGA_DMUL(fh, PL), GGA_DMUL(fl, PH)));
    rh = GGA_DADD(p, e);
    rl = GGA_DADD(GGA_DSUB(p, rh), e);
    if (negf) { rh = -rh; rl = -rl; }
  }
  const double ks = gga_ksin(rh, rl), kc = gga_kcos(rh, rl);
  double s, c;
  switch (q & 3u) {
    case 0: s = ks; c = kc; break;
    case 1: s = kc; c = -ks; break;
    case 2: s = -ks; c = -kc; break;
    default: s = -kc; c = ks; break;
  }
  *sn = negx ? -s : s;
  *cs = c;
}

#endif /* GGA_DETMATH_H_ */
