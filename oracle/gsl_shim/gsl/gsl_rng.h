/* Minimal stand-in for <gsl/gsl_rng.h>: TEST INFRASTRUCTURE ONLY.
 *
 * GSL is an un-vendored system dependency of the reference (Makefile:7,10;
 * shared/gen_func.hpp:12) and is not installed in this image.  The reference
 * uses exactly five symbols, all for the `gsl_rng_taus` generator:
 *   gsl_rng_alloc / gsl_rng_set   (ngsDist.cpp:179-180)
 *   gsl_rng_uniform               (shared/gen_func.cpp:118)
 *   gsl_rng_free                  (ngsDist.cpp:312)
 *   type gsl_rng                  (ngsDist.hpp:36)
 * This header restates the published algorithm (L'Ecuyer 1996, three-component
 * combined Tausworthe, as shipped in GSL rng/taus.c: 69069 LCG seeding, six
 * warm-up draws, output / 2^32) so that oracle/_ref/ngsDist can be linked
 * without GSL.  Known answer: seed 1 -> 10 000th draw = 2733957125.
 */
#ifndef NGSD_ORACLE_GSL_RNG_SHIM_H
#define NGSD_ORACLE_GSL_RNG_SHIM_H

#include <stdlib.h>

typedef struct { int id; } gsl_rng_type;
typedef struct { unsigned long s1, s2, s3; } gsl_rng;

static const gsl_rng_type ngsd_shim_taus_type = { 1 };
static const gsl_rng_type *gsl_rng_taus __attribute__((unused)) = &ngsd_shim_taus_type;

static inline unsigned long gsl_rng_get(gsl_rng *r) {
  const unsigned long M = 0xffffffffUL;
#define NGSD_TAUS(s, a, b, c, d) ((((s) & (c)) << (d)) & M) ^ (((((s) << (a)) & M) ^ (s)) >> (b))
  r->s1 = NGSD_TAUS(r->s1, 13, 19, 4294967294UL, 12);
  r->s2 = NGSD_TAUS(r->s2, 2, 25, 4294967288UL, 4);
  r->s3 = NGSD_TAUS(r->s3, 3, 11, 4294967280UL, 17);
#undef NGSD_TAUS
  return r->s1 ^ r->s2 ^ r->s3;
}

static inline void gsl_rng_set(gsl_rng *r, unsigned long seed) {
  const unsigned long M = 0xffffffffUL;
  int i;
  if (seed == 0) seed = 1;
  r->s1 = (69069UL * seed) & M;
  r->s2 = (69069UL * r->s1) & M;
  r->s3 = (69069UL * r->s2) & M;
  for (i = 0; i < 6; i++) gsl_rng_get(r);
}

static inline gsl_rng *gsl_rng_alloc(const gsl_rng_type *t) {
  gsl_rng *r = (gsl_rng *) calloc(1, sizeof(gsl_rng));
  (void) t;
  if (r) gsl_rng_set(r, 0);
  return r;
}

static inline double gsl_rng_uniform(gsl_rng *r) { return gsl_rng_get(r) / 4294967296.0; }

static inline void gsl_rng_free(gsl_rng *r) { free(r); }

#endif
