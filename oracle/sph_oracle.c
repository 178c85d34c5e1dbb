/*
 * sph_oracle.c -- CPU oracle (test infrastructure only; see sph_oracle.h).
 * Instantiates sph_oracle_impl.inc for the three (T, cT) combinations.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "sph_oracle.h"

#define T double
#define CT double
#define SUF f64
#include "sph_oracle_impl.inc"
#include "tlsph_oracle_impl.inc"
#undef T
#undef CT
#undef SUF

#define T float
#define CT float
#define SUF f32
#include "sph_oracle_impl.inc"
#include "tlsph_oracle_impl.inc"
#undef T
#undef CT
#undef SUF

#define T float
#define CT double
#define SUF f32c64
#include "sph_oracle_impl.inc"
#include "tlsph_oracle_impl.inc"
#undef T
#undef CT
#undef SUF

int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
